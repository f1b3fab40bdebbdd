"""Dev check of the hybrid backward (mode 2: direct kernel + sorting kernel for the coarse levels) on a GPU box:
parity against the direct kernels / the CPU oracle on small shapes, event timings at the cfg3 shape per split.

    python tests/dev/gpu_hybrid_check.py [--time] [--no-oracle]
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


def rel(got, want):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)


def bwd(value, ss, lsi, loc, attn, go, mode, split=8):
    b200.set_tiled_mode(mode)
    b200.set_hybrid_split(split)
    r = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    torch.cuda.synchronize()
    b200.set_tiled_mode(0)
    return r


def case(name, shapes, batch, dtype, kind, split, oracle=True, seed=1, **kw):
    dev = "cuda"
    mk = workloads.make_encoder_inputs if kind == "encoder" else workloads.make_uniform_inputs
    value, ss, lsi, loc, attn = mk(shapes, batch, dtype, seed=seed, device=dev, **kw)
    g = torch.Generator(device=dev).manual_seed(seed + 7)
    go = torch.randn(value.shape[0], loc.shape[1], value.shape[2] * 32, generator=g, device=dev).to(dtype)
    h = bwd(value, ss, lsi, loc, attn, go, 2, split)
    d = bwd(value, ss, lsi, loc, attn, go, 0)
    rec = {"case": name, "split": split, "hybrid_vs_direct": {k: rel(a, b) for k, a, b in zip(("gv", "gl", "ga"), h, d)}}
    if oracle:
        from oracle import ms_deform_attn_oracle_grads
        ref = ms_deform_attn_oracle_grads(value.float().cpu(), ss.cpu(), loc.cpu(), attn.cpu(), go.float().cpu())[1:]
        rec["hybrid_vs_oracle"] = {k: rel(a, b) for k, a, b in zip(("gv", "gl", "ga"), h, ref)}
        rec["ok"] = bool(max(rec["hybrid_vs_oracle"].values()) < 2e-2)
    print(json.dumps(rec), flush=True)
    return rec.get("ok", True)


def timing(shapes, batch, dtype, reps=8, kind="encoder"):
    dev = "cuda"
    lib = b200.load_library()
    mk = workloads.make_encoder_inputs if kind == "encoder" else workloads.make_uniform_inputs
    value, ss, lsi, loc, attn = mk(shapes, batch, dtype, device=dev)
    go = torch.randn(batch, loc.shape[1], 256, device=dev).to(dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for mode, split in ((0, 0), (1, 0), (2, 1), (2, 8), (2, 30), (2, 100)):
        b200.set_tiled_mode(mode)
        b200.set_hybrid_split(max(split, 1))
        ts = []
        for r in range(reps + 3):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        lib.msda_profile_enable(1)
        for _ in range(3):
            MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
        torch.cuda.synchronize()
        kinds = {}
        for ms, k in _lib.profile_collect():
            kinds.setdefault(k, []).append(ms)
        lib.msda_profile_enable(0)
        print(json.dumps({"timing": kind, "mode": mode, "split": split, "bwd_ms": round(ts[len(ts) // 2], 4),
                          "kernel_ms_by_kind": {str(k): round(sorted(v)[len(v) // 2], 4) for k, v in kinds.items()}}), flush=True)
    b200.set_tiled_mode(0)


if __name__ == "__main__":
    oracle = "--no-oracle" not in sys.argv
    bf, hf = torch.bfloat16, torch.float16
    pyr = [(32, 32), (16, 16), (8, 8), (4, 4)]
    cfg3 = [(128, 128), (64, 64), (32, 32), (16, 16)]
    if "--time" in sys.argv:
        timing(cfg3, 16, bf)
        timing(cfg3, 16, bf, kind="uniform")
    ok = True
    for split in (1, 8, 30):
        ok &= case("pyramid_small", pyr, 2, bf, "encoder", split, oracle)
        ok &= case("pyramid_small_f16", pyr, 2, hf, "encoder", split, oracle)
        ok &= case("pyramid_uniform", pyr, 2, bf, "uniform", split, oracle)
        ok &= case("odd_shapes", [(25, 38), (13, 19), (7, 10)], 2, bf, "encoder", split, oracle)
        ok &= case("non_nested", [(8, 8), (24, 40), (5, 3)], 1, bf, "encoder", split, oracle)
        ok &= case("cfg3_one_image", cfg3, 1, bf, "encoder", split, oracle)
    print("ALL_OK" if ok else "SOME_FAILED", flush=True)
