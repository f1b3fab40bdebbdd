"""Call-site shape builders and the synthetic workloads of BASELINE.json.

The builders restate what upstream's ``MSDeformAttnTransformerEncoderOnly.forward`` /
``MSDeformAttnTransformerEncoder.get_reference_points`` compute before calling ``MSDeformAttn``:
``spatial_shapes`` (L, 2) int64 rows (H, W); ``level_start_index = cat(0, prod(1).cumsum(0)[:-1])``;
reference points = pixel centres ``linspace(0.5, H-0.5, H) / (valid_ratio * H)`` per level, then multiplied
by the valid ratio of the level being sampled.  The reference reaches them only through
``build_model(cfg)`` (/root/reference/training/maskdino/train_full.py:308).
"""
from __future__ import annotations

import math
from typing import Sequence

import torch

#: BASELINE.json configs: name -> (levels (H, W), default batch, dtype of value)
CONFIGS = {
    "cfg1_512_fp32": dict(shapes=[(64, 64), (32, 32), (16, 16)], batch=2, dtype=torch.float32, kind="encoder"),
    "cfg3_swinl_1024_bf16": dict(shapes=[(128, 128), (64, 64), (32, 32), (16, 16)], batch=16,
                                 dtype=torch.bfloat16, kind="encoder", layers=6),
    "cfg4_decoder_300q_bf16": dict(shapes=[(128, 128), (64, 64), (32, 32), (16, 16)], batch=16,
                                   dtype=torch.bfloat16, kind="decoder", queries=300, layers=9),
    "cfg5_2048_bf16": dict(shapes=[(256, 256), (128, 128), (64, 64), (32, 32)], batch=16,
                           dtype=torch.bfloat16, kind="encoder", layers=6),
    # BASELINE.json configs[4]: one batch-sharded training step of the 6-layer pixel-decoder encoder at 2048^2
    # (forward + backward under bf16 autocast, per-layer NCCL gradient all-reduce, fused AdamW); `batch` is the GLOBAL
    # batch, split over the ranks (strong scaling).  BASELINE.json gives no batch size: 8 keeps the 1-GPU leg well inside
    # 180 GB (~10 GB of saved activations per image).
    "cfg5_train_step_2048": dict(shapes=[(256, 256), (128, 128), (64, 64), (32, 32)], batch=8, dtype=torch.bfloat16,
                                 kind="train_step", layers=6, d_model=256, heads=8, points=4, d_ffn=2048),
    # BASELINE.json configs[3] as a training step: the 9-layer deformable decoder (self-attention, cross-attention on the
    # encoder memory, FFN) with 300 box queries, batch-sharded like the encoder step
    "cfg4_decoder_step_300q": dict(shapes=[(128, 128), (64, 64), (32, 32), (16, 16)], batch=16, dtype=torch.bfloat16,
                                   kind="decoder_step", layers=9, queries=300, d_model=256, heads=8, points=4, d_ffn=2048),
    # the same step at the cfg3 geometry (1024^2), for quick runs
    "cfg3_train_step_1024": dict(shapes=[(128, 128), (64, 64), (32, 32), (16, 16)], batch=16, dtype=torch.bfloat16,
                                 kind="train_step", layers=6, d_model=256, heads=8, points=4, d_ffn=2048),
}


def make_spatial_shapes(shapes: Sequence[Sequence[int]], device=None) -> torch.Tensor:
    return torch.as_tensor([list(s) for s in shapes], dtype=torch.long, device=device)


def make_level_start_index(spatial_shapes: torch.Tensor) -> torch.Tensor:
    return torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))


def get_reference_points(spatial_shapes: torch.Tensor, valid_ratios: torch.Tensor, device=None) -> torch.Tensor:
    """(N, S, L, 2) reference points of every pixel of every level, as the encoder builds them.
    valid_ratios: (N, L, 2) (w, h) fraction of each level that is not padding."""
    device = device if device is not None else valid_ratios.device
    pts = []
    for lvl, (H_, W_) in enumerate(spatial_shapes.tolist()):
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                      torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device),
                                      indexing="ij")
        ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
        ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
        pts.append(torch.stack((ref_x, ref_y), -1))
    reference_points = torch.cat(pts, 1)
    return reference_points[:, :, None] * valid_ratios[:, None]


def decoder_reference_points_input(reference_points: torch.Tensor, valid_ratios: torch.Tensor) -> torch.Tensor:
    """What the deformable decoder hands to its cross-attention ``MSDeformAttn`` as ``reference_points`` (upstream
    ``TransformerDecoder.forward`` in maskdino/modeling/transformer_decoder/dino_decoder.py, from Deformable-DETR):
    per-level copies of the (sigmoid) query boxes / points, scaled by each level's valid ratio.

    reference_points: (N, nq, 4) boxes (cx, cy, w, h) or (N, nq, 2) points in [0, 1]; valid_ratios: (N, L, 2) (w, h).
    Returns (N, nq, L, 4) or (N, nq, L, 2)."""
    if reference_points.shape[-1] == 4:
        return reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None]
    if reference_points.shape[-1] == 2:
        return reference_points[:, :, None] * valid_ratios[:, None]
    raise ValueError("reference_points must end in 2 or 4 coordinates")


def init_offset_pattern(n_heads: int, n_levels: int, n_points: int) -> torch.Tensor:
    """(M, L, P, 2) the sampling-offset bias MSDeformAttn._reset_parameters installs."""
    thetas = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
    for i in range(n_points):
        grid[:, :, i, :] *= i + 1
    return grid


def _softmax_weights(N, Lq, M, L, P, gen, device):
    w = torch.randn(N, Lq, M, L * P, generator=gen, device=device, dtype=torch.float32)
    return torch.softmax(w, -1).view(N, Lq, M, L, P).contiguous()


def make_encoder_inputs(shapes, batch, dtype, n_heads=8, head_dim=32, n_points=4, seed=1234, device="cuda",
                        offset_sigma_px=2.0):
    """Encoder-like workload: every pixel of every level is a query; sampling locations are
    the pixel-centre reference points plus N(0, sigma^2) pixel offsets of the sampled level."""
    gen = torch.Generator(device=device).manual_seed(seed)
    ss = make_spatial_shapes(shapes, device)
    lsi = make_level_start_index(ss)
    S = int(ss.prod(1).sum())
    L = len(shapes)
    value = torch.randn(batch, S, n_heads, head_dim, generator=gen, device=device, dtype=torch.float32).to(dtype)
    ref = get_reference_points(ss, torch.ones(batch, L, 2, device=device), device)          # (N, S, L, 2)
    off = torch.randn(batch, S, n_heads, L, n_points, 2, generator=gen, device=device) * offset_sigma_px
    norm = torch.stack([ss[:, 1], ss[:, 0]], -1).to(torch.float32)                          # (L, 2) (W, H)
    loc = (ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]).contiguous()
    attn = _softmax_weights(batch, S, n_heads, L, n_points, gen, device)
    return value, ss, lsi, loc, attn


def make_decoder_inputs(shapes, batch, dtype, queries=300, n_heads=8, head_dim=32, n_points=4, seed=1234,
                        device="cuda"):
    """Decoder-like workload: box reference points (centre U(0,1)^2, size U(0.05, 0.5)), offsets = the
    init-bias pattern + N(0, 1), scaled as MSDeformAttn.forward does for 4-d references."""
    gen = torch.Generator(device=device).manual_seed(seed)
    ss = make_spatial_shapes(shapes, device)
    lsi = make_level_start_index(ss)
    S = int(ss.prod(1).sum())
    L = len(shapes)
    value = torch.randn(batch, S, n_heads, head_dim, generator=gen, device=device, dtype=torch.float32).to(dtype)
    ctr = torch.rand(batch, queries, 1, 2, generator=gen, device=device).expand(-1, -1, L, -1)
    wh = torch.rand(batch, queries, 1, 2, generator=gen, device=device).expand(-1, -1, L, -1) * 0.45 + 0.05
    off = init_offset_pattern(n_heads, L, n_points).to(device)[None, None] \
        + torch.randn(batch, queries, n_heads, L, n_points, 2, generator=gen, device=device)
    loc = (ctr[:, :, None, :, None, :] + off / n_points * wh[:, :, None, :, None, :] * 0.5).contiguous()
    attn = _softmax_weights(batch, queries, n_heads, L, n_points, gen, device)
    return value, ss, lsi, loc, attn


def make_uniform_inputs(shapes, batch, dtype, queries=None, n_heads=8, head_dim=32, n_points=4, seed=1234,
                        device="cuda"):
    """Worst-case locality: every sampling location ~ U(-0.05, 1.05)^2."""
    gen = torch.Generator(device=device).manual_seed(seed)
    ss = make_spatial_shapes(shapes, device)
    lsi = make_level_start_index(ss)
    S = int(ss.prod(1).sum())
    L = len(shapes)
    Lq = S if queries is None else queries
    value = torch.randn(batch, S, n_heads, head_dim, generator=gen, device=device, dtype=torch.float32).to(dtype)
    loc = torch.rand(batch, Lq, n_heads, L, n_points, 2, generator=gen, device=device) * 1.1 - 0.05
    attn = _softmax_weights(batch, Lq, n_heads, L, n_points, gen, device)
    return value, ss, lsi, loc, attn


def algorithmic_bytes(N, S, Lq, M, D, L, P, value_bytes, aux_bytes=4):
    """SURVEY.md §8d / BASELINE.md §4: compulsory HBM bytes of one forward and one backward call."""
    pts = N * Lq * M * L * P
    C = M * D
    V = min(N * S * C, 4 * D * pts) * value_bytes
    O = N * Lq * C * value_bytes
    return dict(points=pts, fwd=V + O + 3 * pts * aux_bytes, bwd=2 * V + O + 6 * pts * aux_bytes)


def make_feature_pyramid(shapes, batch, channels=256, seed=1234, device="cuda", pin=False):
    """Synthetic multi-scale backbone features + positional embeddings, per level (N, C, H_l, W_l) fp32 — what the
    pixel decoder's input projections hand to MSDeformAttnTransformerEncoderOnly.forward."""
    gen = torch.Generator(device=device).manual_seed(seed)
    srcs = [torch.randn(batch, channels, h, w, generator=gen, device=device) for h, w in shapes]
    pos = [torch.randn(batch, channels, h, w, generator=gen, device=device) * 0.1 for h, w in shapes]
    if pin:
        srcs = [t.pin_memory() for t in srcs]
        pos = [t.pin_memory() for t in pos]
    return srcs, pos
