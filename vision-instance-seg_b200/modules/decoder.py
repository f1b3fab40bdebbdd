"""detectron2-free harness of the deformable decoder that drives the operator's second call site (SURVEY.md §8 row a8,
BASELINE configs[3]: "MaskDINO decoder deformable cross-attention, 300 queries x 4 levels x 4 points, 9 decoder layers").

Restates the interface of upstream ``maskdino/modeling/transformer_decoder/dino_decoder.py`` (``TransformerDecoder`` /
``DeformableTransformerDecoderLayer``, themselves descendants of DINO / DAB-DETR) -- the classes the reference reaches
through ``build_model(cfg)`` (/root/reference/training/maskdino/train_full.py:308).  That checkout is not vendored by
the reference and cannot be fetched here, so the layer layout, the sequence-first ``(Lq, N, C)`` convention, the
sub-module names (``layers.<i>.{cross_attn, self_attn, norm1, norm2, norm3, linear1, linear2}``, ``ref_point_head``,
``norm``) and the call into ``MSDeformAttn`` --

    cross_attn(with_pos_embed(tgt, query_pos).transpose(0, 1), reference_points.transpose(0, 1).contiguous(),
               memory.transpose(0, 1), spatial_shapes, level_start_index, memory_key_padding_mask).transpose(0, 1)

with ``reference_points_input = reference_points[:, :, None] * cat([valid_ratios, valid_ratios], -1)[None, :]`` (boxes,
last dim 4) -- follow the published code from memory and are *not* checked against a checkpoint (no network).  What the
harness is for: exercising the operator in the decoder's calling pattern (box reference points, ~300 queries, the same
encoder memory handed to every layer) and giving ``share_value_proj`` its real call site.

Everything except ``MSDeformAttn``'s sampling core is stock torch.  Mask / class / box heads, denoising queries and
the query selection of MaskDINO are outside the hot path and are not built.
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import torch
from torch import nn

from .encoder import _get_activation_fn
from .ms_deform_attn import MSDeformAttn
from .stacked_value_proj import share_value_proj, unshare_value_proj


class MLP(nn.Module):
    """Simple multi-layer perceptron (``layers.<i>`` Linears with ReLU in between), as in DETR."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = torch.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


def gen_sineembed_for_position(pos_tensor: torch.Tensor, d_half: int = 128) -> torch.Tensor:
    """(Lq, N, 2|4) normalised (x, y[, w, h]) -> (Lq, N, d_half * 2|4) sine embedding, ordered (y, x[, w, h])."""
    scale = 2 * math.pi
    dim_t = torch.arange(d_half, dtype=torch.float32, device=pos_tensor.device)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / d_half)

    def embed(coord):
        e = (coord * scale)[:, :, None] / dim_t
        return torch.stack((e[:, :, 0::2].sin(), e[:, :, 1::2].cos()), dim=3).flatten(2)

    parts = [embed(pos_tensor[:, :, 1]), embed(pos_tensor[:, :, 0])]
    if pos_tensor.size(-1) == 4:
        parts += [embed(pos_tensor[:, :, 2]), embed(pos_tensor[:, :, 3])]
    elif pos_tensor.size(-1) != 2:
        raise ValueError("Unknown pos_tensor shape(-1):{}".format(pos_tensor.size(-1)))
    return torch.cat(parts, dim=2)


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        # cross attention: the hot path
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        # self attention over the queries
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        # ffn
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, tgt):
        tgt2 = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
        return self.norm3(tgt + self.dropout4(tgt2))

    def forward(self, tgt, tgt_query_pos=None, tgt_reference_points=None, memory=None, memory_key_padding_mask=None,
                memory_level_start_index=None, memory_spatial_shapes=None, self_attn_mask=None):
        """tgt (Lq, N, C); tgt_reference_points (Lq, N, L, 4); memory (S, N, C); masks as in nn.MultiheadAttention."""
        q = k = self.with_pos_embed(tgt, tgt_query_pos)
        tgt2 = self.self_attn(q, k, tgt, attn_mask=self_attn_mask)[0]
        tgt = self.norm2(tgt + self.dropout2(tgt2))
        tgt2 = self.cross_attn(self.with_pos_embed(tgt, tgt_query_pos).transpose(0, 1),
                               tgt_reference_points.transpose(0, 1).contiguous(),
                               memory.transpose(0, 1), memory_spatial_shapes, memory_level_start_index,
                               memory_key_padding_mask).transpose(0, 1)
        tgt = self.norm1(tgt + self.dropout1(tgt2))
        return self.forward_ffn(tgt)


class TransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=True, d_model=256, query_dim=4):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate
        self.query_dim = query_dim
        self.d_model = d_model
        self.ref_point_head = MLP(query_dim // 2 * d_model, d_model, d_model, 2)
        self.bbox_embed: Optional[nn.ModuleList] = None      # set by the model when boxes are refined layer by layer
        self._shared_value = None

    def forward(self, tgt, memory, tgt_mask=None, memory_key_padding_mask=None, refpoints_unsigmoid=None,
                level_start_index=None, spatial_shapes=None, valid_ratios=None):
        """tgt (Lq, N, C); memory (S, N, C); refpoints_unsigmoid (Lq, N, 4); valid_ratios (N, L, 2).
        -> ([per layer (N, Lq, C)], [per layer reference boxes (N, Lq, 4)])"""
        output = tgt
        intermediate = []
        reference_points = refpoints_unsigmoid.sigmoid()
        ref_points = [reference_points]
        for layer_id, layer in enumerate(self.layers):
            reference_points_input = reference_points[:, :, None] \
                * torch.cat([valid_ratios, valid_ratios], -1)[None, :]                     # (Lq, N, L, 4)
            query_sine_embed = gen_sineembed_for_position(reference_points_input[:, :, 0, :], self.d_model // 2)
            query_pos = self.ref_point_head(query_sine_embed)
            output = layer(output, tgt_query_pos=query_pos, tgt_reference_points=reference_points_input, memory=memory,
                           memory_key_padding_mask=memory_key_padding_mask, memory_level_start_index=level_start_index,
                           memory_spatial_shapes=spatial_shapes, self_attn_mask=tgt_mask)
            if self.bbox_embed is not None:
                eps = 1e-5
                r = reference_points.clamp(min=0, max=1)
                before = torch.log(r.clamp(min=eps) / (1 - r).clamp(min=eps))             # inverse sigmoid
                new_reference_points = (self.bbox_embed[layer_id](output) + before).sigmoid()
                reference_points = new_reference_points.detach()
                ref_points.append(new_reference_points)
            intermediate.append(self.norm(output) if self.norm is not None else output)
        return [[o.transpose(0, 1) for o in intermediate], [r.transpose(0, 1) for r in ref_points]]


def set_shared_value_proj(decoder: TransformerDecoder, enabled: bool = True):
    """Opt-in: run the ``value_proj`` of all decoder layers' cross-attention as one stacked GEMM per
    forward pass (``modules/stacked_value_proj.py``).  Parameters and ``state_dict`` are untouched."""
    mods = [layer.cross_attn for layer in decoder.layers]
    if enabled:
        decoder._shared_value = share_value_proj(mods)
    else:
        unshare_value_proj(mods)
        decoder._shared_value = None
    return decoder._shared_value


def build_decoder(d_model=256, nhead=8, num_decoder_layers=9, dim_feedforward=2048, dropout=0.0, activation="relu",
                  num_feature_levels=4, dec_n_points=4) -> TransformerDecoder:
    """MaskDINO's settings by default (9 layers, 300 queries are the caller's, d_ffn 2048, dropout 0)."""
    layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                                              dec_n_points)
    dec = TransformerDecoder(layer, num_decoder_layers, nn.LayerNorm(d_model), True, d_model, 4)
    for p in dec.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
    for m in dec.modules():
        if isinstance(m, MSDeformAttn):
            m._reset_parameters()
    return dec
