"""``MSDeformAttn`` — the nn.Module the pixel-decoder encoder layers (self-attention over all pixels
of all levels) and the deformable decoder layers (cross-attention from ~300 queries) instantiate.

Mirrors upstream ``maskdino/modeling/pixel_decoder/ops/modules/ms_deform_attn.py``: constructor
arguments, attribute names, sub-module / parameter names (``sampling_offsets``, ``attention_weights``,
``value_proj``, ``output_proj`` — the checkpoints the reference loads at
/root/reference/training/maskdino/evaluate.py:113-114 carry these keys), the parameter initialisation and
the forward signature are unchanged.  The four Linears stay stock torch GEMMs; only the sampling core
runs in this repo's CUDA kernels.  Unlike the Mask2Former variant there is no try/except fallback to the
PyTorch core.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from .. import MultiScaleDeformableAttention as MSDA
from ..functions import MSDeformAttnFunction, MSDeformAttnFusedFunction
from .stacked_value_proj import MSDeformAttnStackedFunction


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


def set_fused_preop(module: nn.Module, enabled: bool = True) -> int:
    """Switch the opt-in fused pre-op (softmax + sampling-location arithmetic folded into the kernels, SURVEY.md §8f
    rank 1) on every ``MSDeformAttn`` below ``module``; returns how many were switched."""
    n = 0
    for m in module.modules():
        if isinstance(m, MSDeformAttn):
            m.fused_preop = bool(enabled)
            n += 1
    return n


class MSDeformAttn(nn.Module):
    #: opt-in: fold softmax(L*P) and ``reference_points (+) sampling_offsets`` into the CUDA kernels instead of
    #: materialising ``sampling_locations`` / ``attention_weights`` (same maths and rounding order; no new parameters,
    #: so checkpoints are unaffected).  Constructor signature stays upstream's; flip it with ``set_fused_preop``.
    fused_preop = False
    #: opt-in: (StackedValueProj, index) when this module shares one stacked ``value_proj`` GEMM with the other decoder
    #: layers' cross-attention modules (see modules/stacked_value_proj.py ``share_value_proj``)
    _stacked_value = None

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        """Multi-Scale Deformable Attention Module
        :param d_model      hidden dimension
        :param n_levels     number of feature levels
        :param n_heads      number of attention heads
        :param n_points     number of sampling points per attention head per feature level
        """
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        _d_per_head = d_model // n_heads
        if not _is_power_of_2(_d_per_head):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head "
                          "a power of 2 which is more efficient in our CUDA implementation.")

        self.im2col_step = 128

        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)

        self._reset_parameters()

    def _reset_parameters(self):
        # offsets: zero weight; bias = one unit direction per head (max-norm normalised), scaled by point index + 1
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]) \
            .view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """
        :param query                       (N, Length_{query}, C)
        :param reference_points            (N, Length_{query}, n_levels, 2), range in [0, 1], top-left (0,0),
                                           bottom-right (1, 1), including padding area
                                        or (N, Length_{query}, n_levels, 4), add additional (w, h) to form reference boxes
        :param input_flatten               (N, sum_{l} H_l*W_l, C)
        :param input_spatial_shapes        (n_levels, 2), [(H_0, W_0), ..., (H_{L-1}, W_{L-1})]
        :param input_level_start_index     (n_levels, ), [0, H_0*W_0, H_0*W_0+H_1*W_1, ...]
        :param input_padding_mask          (N, sum_{l} H_l*W_l), True for padding elements
        :return output                     (N, Length_{query}, C)
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        # upstream's assert reads the shapes back from the device (a host sync); the verdict is remembered for the very
        # tensor OBJECT that passed (kept alive here, so its storage cannot be recycled for other contents), its version
        # counter and Len_in: steady-state steps -- and CUDA-graph captures after a warm-up -- do not sync, and a new
        # shape tensor, an in-place edit or another input length is checked again
        chk = self.__dict__.get("_shapes_checked")
        if chk is None or chk[0] is not input_spatial_shapes or chk[1] != input_spatial_shapes._version or chk[2] != Len_in:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
            self.__dict__["_shapes_checked"] = (input_spatial_shapes, input_spatial_shapes._version, Len_in)

        stacked = None
        if self._stacked_value is not None and input_flatten.is_cuda:
            proj, index = self._stacked_value
            stacked = proj.value_for(index, input_flatten, input_padding_mask)   # (view, value_all, grad state)
            value = stacked[0]
        else:
            if input_flatten.is_cuda and not input_flatten.is_contiguous():
                # the decoder passes `memory.transpose(0, 1)`; on a strided operand F.linear runs matmul plus a separate
                # broadcast add of the bias over the whole output.  One fused transpose (+ autocast cast) to a dense
                # operand keeps the bias in the GEMM epilogue -- same stock Linear, same numbers as on a dense input.
                dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else input_flatten.dtype
                input_flatten = input_flatten.to(dtype=dtype, memory_format=torch.contiguous_format)
            value = self.value_proj(input_flatten)
            if input_padding_mask is not None:
                value = value.masked_fill(input_padding_mask[..., None], float(0))
            value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        sampling_offsets = self.sampling_offsets(query) \
            .view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = self.attention_weights(query) \
            .view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                             .format(reference_points.shape[-1]))
        if stacked is None and self.fused_preop and MSDA.fused_supported(value, reference_points) \
                and not (torch.is_grad_enabled() and reference_points.requires_grad):
            output = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                                     reference_points.contiguous(), sampling_offsets.contiguous(),
                                                     attention_weights.contiguous(), self.im2col_step)
            return self.output_proj(output)
        attention_weights = F.softmax(attention_weights, -1) \
            .view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:        # encoder: points; offsets are in pixels of each level
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:      # decoder: boxes; offsets are fractions of the box size
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                             .format(reference_points.shape[-1]))
        if stacked is not None:
            output = MSDeformAttnStackedFunction.apply(value, stacked[1], stacked[2], self._stacked_value[1],
                                                       input_spatial_shapes, input_level_start_index,
                                                       sampling_locations.contiguous(), attention_weights.contiguous(),
                                                       self.im2col_step)
            return self.output_proj(output)
        output = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index,
                                            sampling_locations, attention_weights, self.im2col_step)
        output = self.output_proj(output)
        return output
