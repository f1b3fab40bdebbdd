"""Many-seed soak of the decoder-side paths: sparse levels adding straight into grad_value (mixed with bucketed levels in
one call), the strided entry points on a random layer of a stacked (N, S, K, M, D) projection, fp32 / bf16 / fp16 values,
plain and fused operators.  Development aid (oracle-checked):  python tests/dev/fuzz_decoder.py [first_seed last_seed]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
from oracle import ms_deform_attn_fused_oracle_grads, ms_deform_attn_oracle_grads
from tests.helpers import lsi_of, rel_to_max

pkg.load_library()
dev = "cuda:0"
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 200
bad = 0
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 6e-3}


def smooth_mask(locxy, shapes):
    """(N, Lq, M, L, P, 2) bool: False where the sampling point lies within 1e-4 px of an integer pixel line -- the bilinear
    gradient with respect to the location jumps there, so fp32 and the fp64 oracle may legitimately pick different sides."""
    wh = torch.tensor([[w, h] for h, w in shapes], dtype=torch.double)[None, None, None, :, None, :]
    px = locxy.double() * wh - 0.5
    ok = ((px - px.round()).abs() > 1e-4).all(-1, keepdim=True)
    return ok.expand_as(locxy)


def report(what, seed, c, err, tol):
    global bad
    if not (err < tol):
        bad += 1
        print("FAIL", what, seed, c, f"{err:.3e} >= {tol}", flush=True)


for seed in range(lo, hi):
    rng = random.Random(seed * 13 + 5)
    L = rng.randint(1, 4)
    shapes = [(rng.randint(1, 40), rng.randint(1, 40)) for _ in range(L)]
    c = dict(N=rng.randint(1, 3), Lq=rng.choice([1, 2, 7, 30, 100, 300]), M=rng.choice([1, 2, 4, 8]), D=rng.choice([16, 32, 64]),
             L=L, P=rng.randint(1, 4), shapes=shapes, K=rng.randint(1, 4))
    g = torch.Generator().manual_seed(seed)
    ss = torch.as_tensor(shapes, dtype=torch.long)
    S = int(ss.prod(1).sum())
    N, Lq, M, D, P, K = c["N"], c["Lq"], c["M"], c["D"], c["P"], c["K"]
    sparse = [4 * Lq * P <= h * w for h, w in shapes]
    c["sparse"] = sparse
    dtype = rng.choice([torch.float32, torch.bfloat16, torch.float16])
    layer = rng.randrange(K)
    value_all = torch.randn(N, S, K, M, D, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.3 - 0.15
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    go = torch.randn(N, Lq, M * D, generator=g)
    ssd, lsid = ss.to(dev), lsi_of(ss).to(dev)
    old = MSDA.backward_flags
    try:
        MSDA.backward_flags = rng.choice([0, 0, 4, 8 << 8])
        c["flags"] = MSDA.backward_flags
        va = value_all.to(dev, dtype)
        want = ms_deform_attn_oracle_grads(value_all[:, :, layer].to(dtype).double(), ss, loc.double(), attn.double(),
                                           go.to(dtype).double())
        # (1) strided view of the stacked projection
        out = MSDA.ms_deform_attn_forward_stacked(va, layer, ssd, lsid, loc.to(dev), attn.to(dev), 64)
        gbuf = torch.full_like(va, float("nan"))
        gl, ga = MSDA.ms_deform_attn_backward_stacked(va, layer, ssd, lsid, loc.to(dev), attn.to(dev), go.to(dev, dtype), gbuf, 64)
        ok = smooth_mask(loc, shapes)
        for name, a, b in zip(("out", "gv", "gl", "ga"), (out, gbuf[:, :, layer], gl, ga), want):
            if name == "gl":
                a, b = a.cpu() * ok, b * ok
            report(f"stacked {dtype} {name}", seed, c, rel_to_max(a, b), TOL[dtype])
        others = [k for k in range(K) if k != layer]
        if others and not torch.isnan(gbuf[:, :, others]).all():
            report("stacked wrote outside its layer", seed, c, 1.0, 0.5)
        # (2) dense plain operator
        dense = va[:, :, layer].contiguous()
        out2 = MSDA.ms_deform_attn_forward(dense, ssd, lsid, loc.to(dev), attn.to(dev), 64)
        gv2, gl2, ga2 = MSDA.ms_deform_attn_backward(dense, ssd, lsid, loc.to(dev), attn.to(dev), go.to(dev, dtype), 64)
        for name, a, b in zip(("out", "gv", "gl", "ga"), (out2, gv2, gl2, ga2), want):
            if name == "gl":
                a, b = a.cpu() * ok, b * ok
            report(f"dense {dtype} {name}", seed, c, rel_to_max(a, b), TOL[dtype])
        # (3) fused operator (box or point references), dense
        R = rng.choice([2, 4])
        ref = torch.rand(N, Lq, L, R, generator=g)
        if R == 4:
            ref[..., 2:] = ref[..., 2:] * 0.5 + 0.02
        off = torch.randn(N, Lq, M, L, P, 2, generator=g) * 2
        lg = torch.randn(N, Lq, M, L * P, generator=g) * 2
        outf = MSDA.ms_deform_attn_fused_forward(dense, ssd, lsid, ref.to(dev), off.to(dev), lg.to(dev), 64)
        gvf, gof, glf = MSDA.ms_deform_attn_fused_backward(dense, ssd, lsid, ref.to(dev), off.to(dev), lg.to(dev), go.to(dev, dtype), 64)
        wantf = ms_deform_attn_fused_oracle_grads(value_all[:, :, layer].to(dtype).double(), ss, ref.double(), off.double(),
                                                  lg.double(), go.to(dtype).double())
        if R == 2:
            locf = ref.double()[:, :, None, :, None, :] + off.double() / torch.tensor([[w, h] for h, w in shapes], dtype=torch.double)[None, None, None, :, None, :]
        else:
            locf = ref.double()[:, :, None, :, None, :2] + off.double() / P * ref.double()[:, :, None, :, None, 2:] * 0.5
        okf = smooth_mask(locf, shapes)
        for name, a, b in zip(("out", "gv", "goff", "glg"), (outf, gvf, gof, glf), wantf):
            if name == "goff":
                a, b = a.cpu() * okf, b * okf
            report(f"fused {dtype} R{R} {name}", seed, c, rel_to_max(a, b), max(TOL[dtype], 2e-5))
        torch.cuda.synchronize()
    except Exception as e:
        print("ERROR", seed, c, repr(e)[:300], flush=True)
        raise
    finally:
        MSDA.backward_flags = old
print("done, failures:", bad)
