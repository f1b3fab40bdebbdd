"""detectron2-free harness of the pixel-decoder encoder that drives the hot path.

Mirrors the interface of upstream ``maskdino/modeling/pixel_decoder/maskdino_encoder.py`` — the classes the reference
reaches through ``build_model(cfg)`` (/root/reference/training/maskdino/train_full.py:308) — closely enough that

* sub-module / parameter names match (``encoder.layers.<i>.self_attn.{sampling_offsets,attention_weights,value_proj,
  output_proj}``, ``norm1``, ``linear1``, ``linear2``, ``norm2``, ``level_embed``), so the transformer part of a MaskDINO
  checkpoint (the ``model_final.pth`` the reference loads at evaluate.py:113-114, keys under
  ``sem_seg_head.pixel_decoder.transformer.``) loads with ``load_state_dict`` after stripping that prefix;
* the call-site metadata is built the same way: ``spatial_shapes`` int64 (H, W) rows, ``level_start_index =
  cat(0, prod(1).cumsum(0)[:-1])``, ``valid_ratios`` from the padding masks, reference points = pixel centres
  ``linspace(0.5, H-0.5, H) / (valid_ratio * H)`` scaled by the valid ratio of the level being sampled.

Everything except ``MSDeformAttn``'s sampling core is stock torch (Linear / LayerNorm / activation) — SURVEY.md §1: the
GEMMs stay library GEMMs.  Parameter count at the MaskDINO Swin-L settings (d_model 256, 8 heads, 4 levels, 4 points,
d_ffn 2048, 6 layers): 6 x 1 282 176 + 1 024 = 7 694 080, the gradient bucket of the batch-sharded training step.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import normal_, xavier_uniform_

from .ms_deform_attn import MSDeformAttn


def _get_activation_fn(activation: str):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "glu":
        return F.glu
    raise RuntimeError(f"activation should be relu/gelu/glu, not {activation}.")


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _get_activation_fn(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        src2 = self.linear2(self.dropout2(self.activation(self.linear1(src))))
        return self.norm2(src + self.dropout3(src2))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        # every pixel of every level is a query; the un-embedded features are the values
        src2 = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes, level_start_index,
                              padding_mask)
        src = self.norm1(src + self.dropout1(src2))
        return self.forward_ffn(src)


def set_fused_encoder_layers(module: nn.Module, enabled: bool = True) -> int:
    """Opt-in: run every encoder layer below ``module`` as one fused autograd node
    (``modules/fused_encoder_layer.py``: bf16 GEMM operands, fp32 residual stream, glue kernels of
    include/msda_encoder_b200.h).  Returns how many encoders were switched."""
    n = 0
    for m in module.modules():
        if isinstance(m, MSDeformAttnTransformerEncoder):
            m.fused_layers = bool(enabled)
            n += 1
    return n


class MSDeformAttnTransformerEncoder(nn.Module):
    #: opt-in fused layers; the constructor signature stays upstream's (flip with ``set_fused_encoder_layers``)
    fused_layers = False

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """(N, S, L, 2): centre of every pixel of every level, normalised by that level's valid extent, then
        expressed in each sampled level's padded frame.  ``spatial_shapes``: the (L, 2) tensor (read back with one
        device->host sync, as upstream does) or a Python list of (H, W) (no sync: CUDA-graph capturable)."""
        reference_points_list = []
        for lvl, (H_, W_) in enumerate(spatial_shapes.tolist() if isinstance(spatial_shapes, torch.Tensor) else spatial_shapes):
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device),
                                          indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            reference_points_list.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(reference_points_list, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                level_embed=None, spatial_shapes_list=None):
        """``level_embed`` (fused layers only): when given, ``pos`` is taken as a constant that already contains it and
        the level-embedding gradient is produced by the fused layers themselves (segmented column sums).
        ``spatial_shapes_list``: the same shapes as Python ints; when given nothing here reads ``spatial_shapes`` back from
        the device, which keeps the whole forward free of host syncs (CUDA-graph capture)."""
        output = src
        shapes_host = spatial_shapes_list if spatial_shapes_list is not None else spatial_shapes.tolist()
        reference_points = self.get_reference_points(shapes_host, valid_ratios, device=src.device)
        if self.fused_layers and pos is not None:
            from .fused_encoder_layer import fused_layer_forward, fused_layer_supported
            if all(fused_layer_supported(layer, src, reference_points) for layer in self.layers):
                sizes = [int(h) * int(w) for h, w in shapes_host]
                bounds, a = [], 0
                for n in sizes:
                    bounds.append((a, a + n))
                    a += n
                reference_points = reference_points.contiguous()
                pos_const = pos.detach().contiguous()
                if level_embed is None and pos.requires_grad:
                    raise RuntimeError("fused encoder layers treat `pos` as a constant: pass level_embed separately")
                out32 = src.contiguous()
                out16 = out32.to(torch.bfloat16)
                for i, layer in enumerate(self.layers):
                    out32, out16 = fused_layer_forward(layer, out32, out16, pos_const, level_embed, reference_points,
                                                       spatial_shapes, level_start_index, padding_mask, bounds,
                                                       want16=i + 1 < len(self.layers))
                return out32
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1, activation="relu",
                 num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        encoder_layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                            num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(encoder_layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        """mask (N, H, W) bool, True = padding -> (N, 2) fraction (w, h) of the level that holds image content."""
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def forward(self, srcs, masks, pos_embeds):
        """srcs / pos_embeds: per level (N, C, H_l, W_l); masks: per level (N, H_l, W_l) bool or None.
        Returns (memory (N, S, C), spatial_shapes (L, 2), level_start_index (L,))."""
        use_masks = masks is not None and any(s.size(2) % 32 or s.size(3) % 32 for s in srcs)
        if not use_masks:
            masks = [torch.zeros((x.size(0), x.size(2), x.size(3)), device=x.device, dtype=torch.bool) for x in srcs]
        src_flatten, mask_flatten, lvl_pos_embed_flatten, spatial_shapes = [], [], [], []
        for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
            _, _, h, w = src.shape
            spatial_shapes.append((h, w))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            mask_flatten.append(mask.flatten(1))
            lvl_pos_embed_flatten.append(pos_embed.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten = torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        lvl_pos_embed_flatten = torch.cat(lvl_pos_embed_flatten, 1)
        shapes_list = [tuple(s) for s in spatial_shapes]
        # the two int64 metadata tensors are cached per (shapes, device): no pageable host->device copy per call, which a
        # CUDA-graph capture would reject
        key = (tuple(shapes_list), str(src_flatten.device))
        cache = self.__dict__.setdefault("_shape_cache", {})
        if key not in cache:
            ss = torch.as_tensor(spatial_shapes, dtype=torch.long, device=src_flatten.device)
            cache[key] = (ss, torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1])))
        spatial_shapes, level_start_index = cache[key]
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        # level_embed is handed over separately for the fused layers, which take the positional term as a constant and
        # return level_embed's gradient themselves; the stock layers ignore it and differentiate through `pos`
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                              mask_flatten if use_masks else None, level_embed=self.level_embed,
                              spatial_shapes_list=shapes_list)
        return memory, spatial_shapes, level_start_index


def encoder_state_dict_keys(num_layers=6):
    """The parameter names a MaskDINO checkpoint holds for this sub-tree (prefix
    ``sem_seg_head.pixel_decoder.transformer.`` stripped); used by the checkpoint-compatibility test."""
    keys = ["level_embed"]
    per_layer = [f"self_attn.{lin}.{wb}" for lin in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
                 for wb in ("weight", "bias")]
    per_layer += [f"{n}.{wb}" for n in ("norm1", "linear1", "linear2", "norm2") for wb in ("weight", "bias")]
    for i in range(num_layers):
        keys += [f"encoder.layers.{i}.{k}" for k in per_layer]
    return keys
