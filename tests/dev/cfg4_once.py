"""Three forward + backward calls at the cfg4 decoder shape (300 queries, batch 16, bf16): the command to wrap in
`ncu --metrics gpu__time_duration.sum` for a launch list of the decoder-side backward (zero / density / absmax / kernel / rounding)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, workloads as W
cfg = W.CONFIGS["cfg4_decoder_300q_bf16"]
v, ss, lsi, loc, attn = W.make_decoder_inputs(cfg["shapes"], cfg["batch"], cfg["dtype"], queries=cfg["queries"], device="cuda")
go = torch.randn(v.shape[0], loc.shape[1], 256, device="cuda").to(cfg["dtype"])
for _ in range(3):
    MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 128)
    MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128)
torch.cuda.synchronize()
