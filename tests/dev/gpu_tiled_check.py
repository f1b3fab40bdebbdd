"""Dev check of the tiled kernels (csrc/msda_tiled.cuh) on a GPU box: tiled vs direct kernels vs the CPU oracle on a set
of shape families, then event timings of both paths at the cfg3 shape.  Prints one JSON line per case.

    python tests/dev/gpu_tiled_check.py [--no-oracle] [--time]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import vision_instance_seg_b200 as b200  # noqa: E402
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA  # noqa: E402
from vision_instance_seg_b200 import workloads  # noqa: E402


def rel(got, want):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)


def run(value, ss, lsi, loc, attn, go, tiled):
    lib = b200.load_library()
    lib.msda_set_tiled_mode(1 if tiled else 0)
    out = MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
    gv, gl, ga = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    torch.cuda.synchronize()
    lib.msda_set_tiled_mode(1)
    return out, gv, gl, ga


def case(name, shapes, batch, dtype, kind, heads=8, points=4, sigma=2.0, oracle=True, seed=1):
    dev = "cuda"
    if kind == "encoder":
        value, ss, lsi, loc, attn = workloads.make_encoder_inputs(shapes, batch, dtype, n_heads=heads, n_points=points,
                                                                  seed=seed, device=dev, offset_sigma_px=sigma)
    else:
        value, ss, lsi, loc, attn = workloads.make_uniform_inputs(shapes, batch, dtype, n_heads=heads, n_points=points,
                                                                  seed=seed, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed + 7)
    go = torch.randn(value.shape[0], loc.shape[1], heads * 32, generator=g, device=dev).to(dtype)
    t = run(value, ss, lsi, loc, attn, go, True)
    d = run(value, ss, lsi, loc, attn, go, False)
    rec = {"case": name, "shapes": shapes, "batch": batch, "dtype": str(dtype), "kind": kind, "points": points,
           "tiled_vs_direct": {k: rel(a, b) for k, a, b in zip(("out", "gv", "gl", "ga"), t, d)}}
    if oracle:
        from oracle import ms_deform_attn_oracle_grads
        ref = ms_deform_attn_oracle_grads(value.float().cpu(), ss.cpu(), loc.cpu(), attn.cpu(), go.float().cpu())
        rec["tiled_vs_oracle"] = {k: rel(a, b) for k, a, b in zip(("out", "gv", "gl", "ga"), t, ref)}
        rec["direct_vs_oracle"] = {k: rel(a, b) for k, a, b in zip(("out", "gv", "gl", "ga"), d, ref)}
        worst = max(rec["tiled_vs_oracle"].values())
        rec["ok"] = bool(worst < 2e-2)
    print(json.dumps(rec), flush=True)
    return rec


def timing(shapes, batch, dtype, reps=10, kind="encoder"):
    dev = "cuda"
    lib = b200.load_library()
    if kind == "encoder":
        value, ss, lsi, loc, attn = workloads.make_encoder_inputs(shapes, batch, dtype, device=dev)
    else:
        value, ss, lsi, loc, attn = workloads.make_uniform_inputs(shapes, batch, dtype, device=dev)
    go = torch.randn(batch, loc.shape[1], 256, device=dev).to(dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for mode in (1, 0):
        lib.msda_set_tiled_mode(mode)
        for fn_name in ("fwd", "bwd"):
            ts = []
            for r in range(reps + 3):
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
            res[f"{'tiled' if mode else 'direct'}_{fn_name}_ms"] = round(ts[len(ts) // 2], 4)
    lib.msda_set_tiled_mode(1)
    # per-kernel event timings from the library's profiler
    lib.msda_profile_enable(1)
    for _ in range(3):
        MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
        MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    torch.cuda.synchronize()
    from vision_instance_seg_b200 import _lib
    recs = _lib.profile_collect()
    lib.msda_profile_enable(0)
    kinds = {}
    for ms, k in recs:
        kinds.setdefault(k, []).append(ms)
    res["kernel_ms_by_kind"] = {str(k): round(sorted(v)[len(v) // 2], 4) for k, v in kinds.items()}
    print(json.dumps({"timing": kind, "shapes": shapes, "batch": batch, **res}), flush=True)


if __name__ == "__main__":
    oracle = "--no-oracle" not in sys.argv
    bf, hf = torch.bfloat16, torch.float16
    pyr = [(32, 32), (16, 16), (8, 8), (4, 4)]
    ok = True
    cases = [
        ("pyramid_small", pyr, 2, bf, "encoder", {}),
        ("pyramid_small_f16", pyr, 2, hf, "encoder", {}),
        ("pyramid_uniform", pyr, 2, bf, "uniform", {}),
        ("pyramid_big_sigma", pyr, 1, bf, "encoder", {"sigma": 5.0}),
        ("odd_shapes", [(25, 38), (13, 19), (7, 10)], 2, bf, "encoder", {}),
        ("odd_shapes_uniform", [(25, 38), (13, 19), (7, 10)], 1, bf, "uniform", {}),
        ("single_level", [(20, 24)], 2, bf, "encoder", {}),
        ("non_nested", [(8, 8), (24, 40), (5, 3)], 1, bf, "encoder", {}),
        ("heads4_points2", pyr, 2, bf, "encoder", {"heads": 4, "points": 2}),
        ("points3", pyr, 1, bf, "encoder", {"points": 3}),
        ("levels5_points8", [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)], 1, bf, "encoder", {"points": 8}),
        ("cfg3_one_image", [(128, 128), (64, 64), (32, 32), (16, 16)], 1, bf, "encoder", {}),
        ("wide", [(6, 200), (3, 100)], 1, bf, "encoder", {}),
        ("tall", [(300, 5), (150, 3)], 1, bf, "encoder", {}),
    ]
    if "--time" in sys.argv:
        timing([(128, 128), (64, 64), (32, 32), (16, 16)], 16, bf)
        timing([(128, 128), (64, 64), (32, 32), (16, 16)], 16, bf, kind="uniform")
    for name, shapes, batch, dtype, kind, kw in cases:
        r = case(name, shapes, batch, dtype, kind, oracle=oracle, **kw)
        ok = ok and r.get("ok", True)
    print("ALL_OK" if ok else "SOME_FAILED", flush=True)
