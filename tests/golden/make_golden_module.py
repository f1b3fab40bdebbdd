"""Generates tests/golden/module_*.npz — module-level golden vectors (whole MSDeformAttn.forward: the four Linears, the
padding mask, softmax over L*P, sampling-location arithmetic for 2-d and 4-d reference points, the sampling core).

Produced by an implementation independent of this repo: ``transformers`` (5.5.x)
``DeformableDetrMultiscaleDeformableAttention`` (modeling_deformable_detr.py), a descendant of the same Deformable-DETR
module as upstream MaskDINO's ``MSDeformAttn`` with the same parameter names, run in float64 on the CPU; gradients by
autograd.  Run from the repo root:  python tests/golden/make_golden_module.py
"""
import os

import numpy as np
import torch
from transformers.models.deformable_detr.configuration_deformable_detr import DeformableDetrConfig
from transformers.models.deformable_detr.modeling_deformable_detr import DeformableDetrMultiscaleDeformableAttention

HERE = os.path.dirname(os.path.abspath(__file__))


def lsi_of(shapes):
    s = torch.as_tensor(shapes, dtype=torch.long)
    return torch.cat((s.new_zeros(1), s.prod(1).cumsum(0)[:-1]))


def make(name, d_model, heads, levels, points, shapes, N, Lq, ref_dim, seed, masked):
    g = torch.Generator().manual_seed(seed)
    cfg = DeformableDetrConfig(d_model=d_model, num_feature_levels=levels, disable_custom_kernels=True)
    m = DeformableDetrMultiscaleDeformableAttention(cfg, num_heads=heads, n_points=points).double()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float64) * (0.3 if p.dim() == 1 else 0.08))
        m.sampling_offsets.bias.mul_(5.0)                     # offsets of a few pixels
    S = sum(h * w for h, w in shapes)
    query = torch.randn(N, Lq, d_model, generator=g, dtype=torch.float64).requires_grad_(True)
    src = torch.randn(N, S, d_model, generator=g, dtype=torch.float64).requires_grad_(True)
    if ref_dim == 2:
        ref = torch.rand(N, Lq, levels, 2, generator=g, dtype=torch.float64)
    else:
        ref = torch.cat([torch.rand(N, Lq, levels, 2, generator=g, dtype=torch.float64),
                         torch.rand(N, Lq, levels, 2, generator=g, dtype=torch.float64) * 0.5 + 0.05], -1)
    padding = torch.zeros(N, S, dtype=torch.bool)
    if masked:
        padding[0, ::5] = True
        padding[-1, -S // 3:] = True
    ss = torch.as_tensor(shapes, dtype=torch.long)
    out, _ = m(query, attention_mask=~padding if masked else None, encoder_hidden_states=src, reference_points=ref,
               spatial_shapes=ss, spatial_shapes_list=[tuple(s) for s in shapes], level_start_index=lsi_of(shapes))
    grad_out = torch.randn(out.shape, generator=g, dtype=torch.float64)
    out.backward(grad_out)
    data = dict(shapes=np.asarray(shapes, dtype=np.int64), level_start_index=lsi_of(shapes).numpy(), heads=np.int64(heads),
                points=np.int64(points), query=query.detach().numpy(), src=src.detach().numpy(), ref=ref.numpy(),
                padding=padding.numpy(), masked=np.bool_(masked), grad_out=grad_out.numpy(), out=out.detach().numpy(),
                grad_query=query.grad.numpy(), grad_src=src.grad.numpy())
    for k, p in m.named_parameters():
        data["param." + k] = p.detach().numpy()
        data["grad." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(name, tuple(out.shape), float(out.abs().max()))


if __name__ == "__main__":
    make("module_encoder_points", 64, 4, 3, 4, [(6, 5), (3, 3), (2, 1)], 2, 6 * 5 + 9 + 2, 2, 21, masked=True)
    make("module_decoder_boxes", 64, 2, 2, 3, [(5, 4), (2, 2)], 2, 7, 4, 22, masked=False)
