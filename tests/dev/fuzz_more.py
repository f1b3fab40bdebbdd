"""Many-seed soak of the paths tests/test_gpu_fuzz.py does not draw: 16-bit offsets / logits in the fused operator, the
fp32-accumulation backward flag, fp16 values, and the encoder glue kernels at random row counts.  Development aid."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, encoder_ops as E, _lib
from oracle import (ms_deform_attn_fused_oracle_grads, ms_deform_attn_oracle_grads, add_layernorm_oracle, colsum_oracle,
                    relu_bwd_colsum_oracle)
from tests.helpers import lsi_of, rel_to_max
from tests.test_gpu_fuzz import _draw

pkg.load_library()
dev = "cuda:0"
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 200
bad = 0

def report(what, seed, c, err, tol):
    global bad
    if not (err < tol):
        bad += 1
        print("FAIL", what, seed, c, f"{err:.3e} >= {tol}", flush=True)

for seed in range(lo, hi):
    c = _draw(seed)
    rng = random.Random(seed * 7 + 1)
    if c["D"] == 24:
        c["D"] = rng.choice([16, 32, 64, 128])
    g = torch.Generator().manual_seed(seed)
    ss = torch.as_tensor(c["shapes"], dtype=torch.long)
    S = int(ss.prod(1).sum())
    N, Lq, M, D, L, P = c["N"], c["Lq"], c["M"], c["D"], c["L"], c["P"]
    value = torch.randn(N, S, M, D, generator=g)
    go = torch.randn(N, Lq, M * D, generator=g)
    try:
        # (1) fused, 16-bit aux
        dtype = rng.choice([torch.bfloat16, torch.float16])
        R = rng.choice([2, 4])
        ref = torch.rand(N, Lq, L, R, generator=g)
        if R == 4:
            ref[..., 2:] = ref[..., 2:] * 0.5 + 0.02
        off = (torch.randn(N, Lq, M, L, P, 2, generator=g) * 2).to(dtype)
        lg = (torch.randn(N, Lq, M, L * P, generator=g) * 2).to(dtype)
        v = value.to(dev, dtype).requires_grad_(True)
        o = off.to(dev).requires_grad_(True)
        l_ = lg.to(dev).requires_grad_(True)
        out = pkg.MSDeformAttnFusedFunction.apply(v, ss.to(dev), lsi_of(ss).to(dev), ref.to(dev), o, l_, 64)
        out.backward(go.to(dev, dtype))
        want = ms_deform_attn_fused_oracle_grads(value.to(dtype).double(), ss, ref.double(), off.double(), lg.double(), go.to(dtype).double())
        tol = 2e-2 if dtype == torch.bfloat16 else 6e-3
        for name, a, b in zip(("out", "gv", "goff", "glg"), (out, v.grad, o.grad, l_.grad), want):
            report(f"fused16 {dtype} R{R} {name}", seed, c, rel_to_max(a, b), tol)
        # (2) plain, fp32-accumulate flag and fp16 values
        dtype = rng.choice([torch.bfloat16, torch.float16])
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.4 - 0.2
        attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
        old = MSDA.backward_flags
        MSDA.backward_flags = rng.choice([0, 2, (8 << 8)])
        v = value.to(dev, dtype).requires_grad_(True)
        lo_ = loc.to(dev).requires_grad_(True)
        at = attn.to(dev).requires_grad_(True)
        out = pkg.MSDeformAttnFunction.apply(v, ss.to(dev), lsi_of(ss).to(dev), lo_, at, 64)
        out.backward(go.to(dev, dtype))
        flags = MSDA.backward_flags
        MSDA.backward_flags = old
        want = ms_deform_attn_oracle_grads(value.to(dtype).double(), ss, loc.double(), attn.double(), go.to(dtype).double())
        tol = 2e-2 if dtype == torch.bfloat16 else 6e-3
        for name, a, b in zip(("out", "gv", "gl", "ga"), (out, v.grad, lo_.grad, at.grad), want):
            report(f"plain {dtype} flags{flags} {name}", seed, c, rel_to_max(a, b), tol)
        # (3) encoder glue kernels
        C = rng.choice([128, 256, 384, 512, 768, 1024])
        rows = rng.randint(1, 3000)
        x = torch.randn(rows, C, generator=g) * 2
        d16 = torch.randn(rows, C, generator=g).to(torch.bfloat16)
        gamma = torch.rand(C, generator=g) + 0.5
        beta = torch.randn(C, generator=g)
        gy = torch.randn(rows, C, generator=g)
        y, y16, mean, rstd = E.add_layernorm_forward(x.to(dev), d16.to(dev), gamma.to(dev), beta.to(dev), 1e-5)
        wy, wm, wr, wdx, wdg, wdb = add_layernorm_oracle(x, d16, gamma, beta, 1e-5, gy)
        report("ln fwd", seed, (rows, C), rel_to_max(y, wy), 1e-5)
        dx, dd, dg, db = E.add_layernorm_backward(gy.to(dev), None, x.to(dev), d16.to(dev), mean, rstd, gamma.to(dev))
        report("ln dx", seed, (rows, C), rel_to_max(dx, wdx), 1e-5)
        report("ln dgamma", seed, (rows, C), rel_to_max(dg, wdg), 3e-5)
        report("ln dbeta", seed, (rows, C), rel_to_max(db, wdb), 3e-5)
        C2 = rng.choice([8, 40, 128, 256, 2048])
        t = torch.randn(rng.randint(1, 3), rng.randint(1, 700), C2, generator=g).to(torch.bfloat16)
        a_, b_ = sorted(rng.sample(range(0, t.shape[1] + 1), 2)) if t.shape[1] > 1 else (0, 1)
        if a_ == b_:
            b_ = a_ + 1
        report("colsum", seed, tuple(t.shape), rel_to_max(E.colsum(t.to(dev), a_, b_), colsum_oracle(t, a_, b_)), 1e-5)
        h = torch.relu(torch.randn(rows, C2, generator=g)).to(torch.bfloat16)
        gr = torch.randn(rows, C2, generator=g).to(torch.bfloat16)
        masked, sums = relu_bwd_colsum_oracle(gr, h)
        gg = gr.to(dev).clone()
        cs = E.relu_bwd_colsum(gg, h.to(dev))
        report("relu colsum", seed, (rows, C2), rel_to_max(cs, sums), 1e-5)
        if not torch.equal(gg.cpu().double(), masked):
            report("relu mask", seed, (rows, C2), 1.0, 0.5)
        torch.cuda.synchronize()
    except Exception as e:
        print("ERROR", seed, c, repr(e)[:300], flush=True)
        raise
print("done, failures:", bad)
