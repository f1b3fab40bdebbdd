"""Generates tests/golden/*.npz — committed golden input/output vectors for the MSDeformAttn path.

The reference repository has no tests or fixtures for this path and does not vendor the operator
(SURVEY.md §0, §8c), so the pins are produced by an implementation that is independent of this repo's
oracle: ``transformers`` (5.5.x) ``MultiScaleDeformableAttention.forward``
(site-packages/transformers/models/deformable_detr/modeling_deformable_detr.py), run in float64 on the CPU,
with its gradients taken by torch autograd.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os

import numpy as np
import torch
from transformers.models.deformable_detr.modeling_deformable_detr import MultiScaleDeformableAttention

HERE = os.path.dirname(os.path.abspath(__file__))


def lsi_of(shapes):
    s = torch.as_tensor(shapes, dtype=torch.long)
    return torch.cat((s.new_zeros(1), s.prod(1).cumsum(0)[:-1]))


def run(value, shapes, loc, attn, grad_out):
    v = value.clone().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    at = attn.clone().requires_grad_(True)
    ss = torch.as_tensor(shapes, dtype=torch.long)
    out = MultiScaleDeformableAttention().forward(v, ss, [tuple(s) for s in shapes], lsi_of(shapes), lo, at, 64)
    out.backward(grad_out)
    return out.detach(), v.grad, lo.grad, at.grad


def save(name, value, shapes, loc, attn, grad_out):
    out, gv, gl, ga = run(value, shapes, loc, attn, grad_out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), value=value.numpy(), shapes=np.asarray(shapes, dtype=np.int64),
                        level_start_index=lsi_of(shapes).numpy(), loc=loc.numpy(), attn=attn.numpy(),
                        grad_out=grad_out.numpy(), out=out.numpy(), grad_value=gv.numpy(), grad_loc=gl.numpy(),
                        grad_attn=ga.numpy())
    print(name, tuple(value.shape), tuple(loc.shape), float(out.abs().max()))


def norm_attn(N, Lq, M, L, P, gen):
    a = torch.rand(N, Lq, M, L, P, generator=gen, dtype=torch.float64) + 1e-5
    return a / a.sum(-1, keepdim=True).sum(-2, keepdim=True)


def main():
    f64 = torch.float64
    # 1. the upstream ops/test.py vector: N,M,D = 1,2,2; Lq,L,P = 2,2,2; shapes [(6,4),(3,2)]; seed 3
    g = torch.Generator().manual_seed(3)
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    value = torch.rand(1, S, 2, 2, generator=g, dtype=f64) * 0.01
    loc = torch.rand(1, 2, 2, 2, 2, 2, generator=g, dtype=f64)
    attn = norm_attn(1, 2, 2, 2, 2, g)
    save("upstream_test_tiny", value, shapes, loc, attn, torch.rand(1, 2, 4, generator=g, dtype=f64))

    # 2. model-shaped small case: M=8, D=32, L=3, P=4, non-square levels, locations incl. out-of-range
    g = torch.Generator().manual_seed(11)
    shapes = [(8, 6), (4, 3), (2, 2)]
    S = sum(h * w for h, w in shapes)
    value = torch.randn(2, S, 8, 32, generator=g, dtype=f64)
    loc = torch.rand(2, 19, 8, 3, 4, 2, generator=g, dtype=f64) * 1.3 - 0.15
    attn = norm_attn(2, 19, 8, 3, 4, g)
    save("model_small_d32", value, shapes, loc, attn, torch.randn(2, 19, 256, generator=g, dtype=f64))

    # 3. edge locations: exactly 0, 1, pixel centres, integer pixel lines (dyadic, so that every implementation
    #    rounds them to the same side of the bilinear kink), just inside / outside the (-1, H) gate
    g = torch.Generator().manual_seed(5)
    shapes = [(4, 8), (2, 4)]
    S = sum(h * w for h, w in shapes)
    value = torch.randn(1, S, 2, 16, generator=g, dtype=f64)
    specials = torch.tensor([0.0, 1.0, 0.5, 0.125, 0.25, 1.0 / 16, 15.0 / 16, -1.0 / 16 + 1e-4, -1.0 / 16 - 1e-3, 1.0 + 1.0 / 16 - 1e-4,
                             1.0 + 1.0 / 8, -0.3, 1.4, 0.999999, 1e-7, 0.75], dtype=f64)
    idx = torch.randint(0, len(specials), (1, 24, 2, 2, 3, 2), generator=g)
    loc = specials[idx]
    attn = norm_attn(1, 24, 2, 2, 3, g)
    attn[0, :4] = 0.0                                       # some all-zero attention rows
    save("edge_locations", value, shapes, loc, attn, torch.randn(1, 24, 32, generator=g, dtype=f64))

    # 4. odd channel count (compatibility kernels) with a single level / single point
    g = torch.Generator().manual_seed(7)
    shapes = [(5, 7)]
    value = torch.randn(2, 35, 3, 30, generator=g, dtype=f64)
    loc = torch.rand(2, 9, 3, 1, 1, 2, generator=g, dtype=f64)
    attn = norm_attn(2, 9, 3, 1, 1, g)
    save("odd_channels_d30", value, shapes, loc, attn, torch.randn(2, 9, 90, generator=g, dtype=f64))

    # 5. decoder-shaped: few queries, 4 levels, D=32
    g = torch.Generator().manual_seed(13)
    shapes = [(8, 8), (4, 4), (2, 2), (1, 1)]
    S = sum(h * w for h, w in shapes)
    value = torch.randn(1, S, 8, 32, generator=g, dtype=f64)
    loc = torch.rand(1, 7, 8, 4, 4, 2, generator=g, dtype=f64)
    attn = norm_attn(1, 7, 8, 4, 4, g)
    save("decoder_small", value, shapes, loc, attn, torch.randn(1, 7, 256, generator=g, dtype=f64))


if __name__ == "__main__":
    main()
