"""GPU: seeded random-shape fuzz of the operator (plain and fused entry points, fp32 and bf16) against the CPU oracle.
Every draw picks its own batch, query count, head count / width, level count and (non-square, possibly 1-pixel) level
shapes, so ragged warps, single-pixel levels and odd L*P all get exercised together."""
import random

import pytest
import torch

from oracle import ms_deform_attn_fused_oracle_grads, ms_deform_attn_oracle_grads
from tests.helpers import lsi_of, rel_to_max

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


@pytest.fixture(scope="module")
def ops(built_library):
    assert torch.cuda.is_available()
    import vision_instance_seg_b200 as pkg
    pkg.load_library()
    return pkg


def _draw(seed):
    rng = random.Random(seed)
    L = rng.randint(1, 5)
    shapes = [(rng.randint(1, 14), rng.randint(1, 14)) for _ in range(L)]
    return dict(N=rng.randint(1, 3), Lq=rng.randint(1, 70), M=rng.choice([1, 2, 3, 5, 8]), D=rng.choice([16, 32, 64, 128, 24]),
                L=L, P=rng.randint(1, 5), shapes=shapes, seed=seed)


@pytest.mark.parametrize("seed", list(range(16)) + [1001, 1016])      # 1001 (L*P = 2), 1016 (L*P = 3, one pair per warp): found by tests/dev/fuzz_many.py
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_random_shapes_plain_operator(ops, seed, dtype):
    c = _draw(seed)
    g = torch.Generator().manual_seed(seed)
    ss = torch.as_tensor(c["shapes"], dtype=torch.long)
    S = int(ss.prod(1).sum())
    value = torch.randn(c["N"], S, c["M"], c["D"], generator=g)
    loc = torch.rand(c["N"], c["Lq"], c["M"], c["L"], c["P"], 2, generator=g) * 1.4 - 0.2
    attn = torch.softmax(torch.randn(c["N"], c["Lq"], c["M"], c["L"] * c["P"], generator=g), -1).view(c["N"], c["Lq"], c["M"], c["L"], c["P"])
    go = torch.randn(c["N"], c["Lq"], c["M"] * c["D"], generator=g)
    dev = "cuda:0"
    v = value.to(dev, dtype).requires_grad_(True)
    lo = loc.to(dev).requires_grad_(True)
    at = attn.to(dev).requires_grad_(True)
    out = ops.MSDeformAttnFunction.apply(v, ss.to(dev), lsi_of(ss).to(dev), lo, at, 64)
    out.backward(go.to(dev, dtype))
    want = ms_deform_attn_oracle_grads(value.to(dtype).double(), ss, loc.double(), attn.double(), go.to(dtype).double())
    for name, a, b in zip(("out", "grad_value", "grad_loc", "grad_attn"), (out, v.grad, lo.grad, at.grad), want):
        assert rel_to_max(a, b) < TOL[dtype], (name, c)


@pytest.mark.parametrize("seed", list(range(100, 112)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_random_shapes_fused_operator(ops, seed, dtype):
    c = _draw(seed)
    if c["D"] == 24:
        c["D"] = 32                              # fused kernels: vector head dims only
    g = torch.Generator().manual_seed(seed)
    ss = torch.as_tensor(c["shapes"], dtype=torch.long)
    S = int(ss.prod(1).sum())
    R = 2 if seed % 2 == 0 else 4
    value = torch.randn(c["N"], S, c["M"], c["D"], generator=g)
    ref = torch.rand(c["N"], c["Lq"], c["L"], R, generator=g)
    if R == 4:
        ref[..., 2:] = ref[..., 2:] * 0.5 + 0.02
    off = torch.randn(c["N"], c["Lq"], c["M"], c["L"], c["P"], 2, generator=g) * 2
    logits = torch.randn(c["N"], c["Lq"], c["M"], c["L"] * c["P"], generator=g) * 2
    go = torch.randn(c["N"], c["Lq"], c["M"] * c["D"], generator=g)
    dev = "cuda:0"
    v = value.to(dev, dtype).requires_grad_(True)
    o = off.to(dev).requires_grad_(True)
    lg = logits.to(dev).requires_grad_(True)
    out = ops.MSDeformAttnFusedFunction.apply(v, ss.to(dev), lsi_of(ss).to(dev), ref.to(dev), o, lg, 64)
    out.backward(go.to(dev, dtype))
    want = ms_deform_attn_fused_oracle_grads(value.to(dtype).double(), ss, ref.double(), off.double(), logits.double(),
                                             go.to(dtype).double())
    for name, a, b in zip(("out", "grad_value", "grad_offsets", "grad_logits"), (out, v.grad, o.grad, lg.grad), want):
        assert rel_to_max(a, b) < TOL[dtype], (name, R, c)
