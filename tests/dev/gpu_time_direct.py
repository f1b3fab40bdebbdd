"""Event timings of the direct forward / backward at the cfg3 shape (batch 16, bf16): whole call and the main kernel
(library event pairs).  One JSON line; used for A/B runs of build- or environment-level switches.

    [MSDA_B200_CARVEOUT=-1] python tests/dev/gpu_time_direct.py [tag]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import vision_instance_seg_b200 as b200  # noqa: E402
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from vision_instance_seg_b200 import _lib, workloads  # noqa: E402

cfg3 = [(128, 128), (64, 64), (32, 32), (16, 16)]
dev = "cuda"
lib = b200.load_library()
res = {"tag": sys.argv[1] if len(sys.argv) > 1 else "", "carveout_env": os.environ.get("MSDA_B200_CARVEOUT")}
for dtype, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
    value, ss, lsi, loc, attn = workloads.make_encoder_inputs(cfg3, 16, dtype, device=dev)
    go = torch.randn(16, loc.shape[1], 256, device=dev).to(dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for fn_name in ("fwd", "bwd"):
        ts = []
        for r in range(13):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if fn_name == "fwd":
                MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
            else:
                MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[f"{name}_{fn_name}_ms"] = round(ts[len(ts) // 2], 4)
    lib.msda_profile_enable(1)
    for _ in range(5):
        MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
        MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    torch.cuda.synchronize()
    kinds = {}
    for ms, k in _lib.profile_collect():
        kinds.setdefault(k, []).append(ms)
    lib.msda_profile_enable(0)
    res[f"{name}_kernel_ms"] = {str(k): round(sorted(v)[len(v) // 2], 4) for k, v in kinds.items()}
    del value, loc, attn, go
print(json.dumps(res), flush=True)
