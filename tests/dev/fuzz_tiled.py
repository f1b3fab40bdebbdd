"""Many-seed soak of the tiled (mode 1) and hybrid (mode 2) kernels of the dense call site against the direct kernels
(mode 0; grad_value against the fp32-accumulation mode) on random dense shapes: 1-5 levels of random, not necessarily
nested sizes, 1-8 heads, 1-8 points, bf16 / fp16, encoder-like offsets of random spread or uniform locations.

    python tests/dev/fuzz_tiled.py [seeds=300] [first=0]
"""
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import vision_instance_seg_b200 as b200  # noqa: E402
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from vision_instance_seg_b200 import _lib, workloads  # noqa: E402


def rel(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def one(seed):
    rng = random.Random(seed)
    L = rng.randint(1, 5)
    if rng.random() < 0.5:                      # pyramid
        h0, w0 = rng.randint(8, 72), rng.randint(8, 72)
        shapes = [(max(1, h0 >> l), max(1, w0 >> l)) for l in range(L)]
    else:
        shapes = [(rng.randint(1, 48), rng.randint(1, 48)) for _ in range(L)]
    if rng.random() < 0.3:
        rng.shuffle(shapes)
    heads = rng.choice([1, 2, 3, 4, 8])
    points = rng.choice([1, 2, 3, 4, 4, 4, 8])
    if L * points > 64:
        points = 4
    batch = rng.randint(1, 3)
    dtype = rng.choice([torch.bfloat16, torch.float16])
    kind = "uniform" if rng.random() < 0.2 else "encoder"
    sigma = rng.choice([0.5, 2.0, 2.0, 4.0, 8.0])
    mode = rng.choice([1, 2, 2])
    split = rng.choice([1, 2, 8, 30, 100])
    if kind == "encoder":
        value, ss, lsi, loc, attn = workloads.make_encoder_inputs(shapes, batch, dtype, n_heads=heads, n_points=points,
                                                                  seed=seed, device="cuda", offset_sigma_px=sigma)
    else:
        value, ss, lsi, loc, attn = workloads.make_uniform_inputs(shapes, batch, dtype, n_heads=heads, n_points=points,
                                                                  seed=seed, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    go = (torch.randn(batch, loc.shape[1], heads * 32, generator=g, device="cuda") * rng.choice([1e-3, 1.0, 30.0])).to(dtype)
    b200.set_tiled_mode(0)
    MSDA.backward_flags = _lib.MSDA_BWD_GRAD_VALUE_FP32_ACCUM
    out0 = MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
    gv0, gl0, ga0 = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    MSDA.backward_flags = _lib.MSDA_BWD_DEFAULT
    b200.set_tiled_mode(mode)
    b200.set_hybrid_split(split)
    out1 = MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
    gv1, gl1, ga1 = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    torch.cuda.synchronize()
    b200.set_tiled_mode(0)
    errs = {"out": rel(out1, out0), "gv": rel(gv1, gv0), "gl": rel(gl1, gl0), "ga": rel(ga1, ga0)}
    ok = errs["out"] < 1e-2 and errs["gv"] < 2e-2 and errs["gl"] < 1e-3 and errs["ga"] < 1e-3
    ok = ok and all(torch.isfinite(t).all() for t in (out1, gv1, gl1, ga1))
    if not ok:
        print(json.dumps({"seed": seed, "shapes": shapes, "heads": heads, "points": points, "batch": batch, "dtype": str(dtype),
                          "kind": kind, "sigma": sigma, "mode": mode, "split": split, **errs}), flush=True)
    return ok


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = sum(0 if one(s) else 1 for s in range(first, first + n))
    print(json.dumps({"seeds": n, "first": first, "failed": bad}), flush=True)
