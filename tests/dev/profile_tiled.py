"""One forward + backward of the operator at the cfg3 geometry (batch from argv, default 4) -- the command ncu wraps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from vision_instance_seg_b200 import workloads  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
shapes = [(128, 128), (64, 64), (32, 32), (16, 16)]
value, ss, lsi, loc, attn = workloads.make_encoder_inputs(shapes, batch, torch.bfloat16, device="cuda")
go = torch.randn(batch, loc.shape[1], 256, device="cuda").to(torch.bfloat16)
for _ in range(2):
    out = MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
    gv, gl, ga = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(gv.float().abs().mean()))
