"""Bucket depth of the fp16 accumulation (expected adds per accumulator element): error against the fp32-accumulation mode
at the cfg3 geometry (2 images, zero-mean and mean-1 gradients) and time of the whole backward at batch 16, per depth."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from vision_instance_seg_b200 import workloads as W, MultiScaleDeformableAttention as MSDA, _lib

dev = "cuda"
cfg = W.CONFIGS["cfg3_swinl_1024_bf16"]
v2, ss, lsi, loc2, attn2 = W.make_encoder_inputs(cfg["shapes"], 2, torch.bfloat16, device=dev, seed=1)
v16, _, _, loc16, attn16 = W.make_encoder_inputs(cfg["shapes"], 16, torch.bfloat16, device=dev, seed=2)
go16 = torch.randn(16, loc16.shape[1], 256, device=dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for depth in (16, 32, 48, 64, 96, 128, 256):
    rec = {"depth": depth}
    for mean in (0.0, 1.0):
        go = (torch.randn(2, loc2.shape[1], 256, device=dev) + mean).to(torch.bfloat16)
        MSDA.backward_flags = _lib.MSDA_BWD_GRAD_VALUE_FP32_ACCUM
        ref = MSDA.ms_deform_attn_backward(v2, ss, lsi, loc2, attn2, go, 128)[0].float()
        MSDA.backward_flags = _lib.accum_depth_flag(depth)
        gv = MSDA.ms_deform_attn_backward(v2, ss, lsi, loc2, attn2, go, 128)[0].float()
        per = []
        for l in range(4):
            a = int(lsi[l]); b = int(lsi[l + 1]) if l < 3 else ref.shape[1]
            per.append(round(float((gv[:, a:b] - ref[:, a:b]).abs().max() / ref[:, a:b].abs().max()), 5))
        rec[f"err_mean{int(mean)}"] = {"total": round(float((gv - ref).abs().max() / ref.abs().max()), 5), "per_level": per}
    MSDA.backward_flags = _lib.accum_depth_flag(depth)
    ts = []
    for r in range(11):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        MSDA.ms_deform_attn_backward(v16, ss, lsi, loc16, attn16, go16, 128)
        e1.record(); torch.cuda.synchronize()
        if r >= 3: ts.append(e0.elapsed_time(e1))
    ts.sort(); rec["bwd_ms"] = round(ts[len(ts) // 2], 4)
    print(json.dumps(rec), flush=True)
