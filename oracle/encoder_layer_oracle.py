"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatements of the encoder-layer glue around the operator,
following upstream ``MSDeformAttnTransformerEncoderLayer.forward`` / ``forward_ffn``
(IDEA-Research/MaskDINO maskdino/modeling/pixel_decoder/maskdino_encoder.py; reached by the reference through
``build_model(cfg)``, /root/reference/training/maskdino/train_full.py:308): ``src = norm(src + branch)``, ReLU FFN, bias
gradients as column sums.  Plain torch on the CPU in float64; gradients by autograd."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def add_layernorm_oracle(x, delta, gamma, beta, eps, grad_y=None):
    """y = LayerNorm(x + delta) (last dim).  With grad_y: also (dx, dgamma, dbeta); d(delta) equals dx."""
    x = x.detach().double().cpu().requires_grad_(True)
    d = delta.detach().double().cpu() if delta is not None else None
    g = gamma.detach().double().cpu().requires_grad_(True)
    b = beta.detach().double().cpu().requires_grad_(True)
    s = x if d is None else x + d
    y = F.layer_norm(s, (x.shape[-1],), g, b, eps)
    mean = s.mean(-1)
    rstd = (s.var(-1, unbiased=False) + eps).rsqrt()
    if grad_y is None:
        return y.detach(), mean.detach(), rstd.detach()
    y.backward(grad_y.detach().double().cpu())
    return y.detach(), mean.detach(), rstd.detach(), x.grad, g.grad, b.grad


def colsum_oracle(g, row_begin=0, row_end=None):
    g = g.detach().double().cpu()
    if g.dim() == 2:
        g = g[None]
    g = g.reshape(-1, g.shape[-2], g.shape[-1])
    return g[:, row_begin:row_end].sum((0, 1))


def relu_bwd_colsum_oracle(g, h):
    g = g.detach().double().cpu()
    masked = g * (h.detach().double().cpu() > 0)
    return masked, masked.reshape(-1, g.shape[-1]).sum(0)
