"""``MSDeformAttnFunction`` — the autograd boundary of the hot path.

Mirrors upstream ``maskdino/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py`` (the module the
reference's models import after ``sys.path.insert(0, MASKDINO_PATH)``,
/root/reference/training/maskdino/train_full.py:15-16): same six positional arguments, same return
shape ``(N, Lq, M*D)``, gradients for ``value``, ``sampling_locations`` and ``attention_weights`` only,
backward not differentiable again.  ``ms_deform_attn_core_pytorch`` (upstream's debug-only PyTorch
reference) is *not* part of the product package; it lives in ``oracle/`` as the test oracle.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, sampling_locations, attention_weights = ctx.saved_tensors
        # autograd may hand over an expanded / transposed gradient; the kernels want it dense
        grad_output = grad_output.contiguous()
        grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
            value, shapes, level_start, sampling_locations, attention_weights, grad_output, ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


class MSDeformAttnFusedFunction(Function):
    """Opt-in fused variant: consumes what ``MSDeformAttn.forward`` holds *before* it builds
    ``sampling_locations`` / softmaxed ``attention_weights`` — ``reference_points (N, Lq, L, 2|4)``, the raw
    ``sampling_offsets (N, Lq, M, L, P, 2)`` and ``attention logits (N, Lq, M, L*P)`` — and returns the same
    ``(N, Lq, M*D)`` output as ``MSDeformAttnFunction`` would on the derived tensors.  Gradients: value, offsets,
    logits.  ``reference_points`` must not require grad (the module composes the plain operator in that case)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, reference_points, sampling_offsets,
                attention_logits, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_fused_forward(value, value_spatial_shapes, value_level_start_index,
                                                   reference_points, sampling_offsets, attention_logits, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, reference_points,
                              sampling_offsets, attention_logits)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, reference_points, sampling_offsets, attention_logits = ctx.saved_tensors
        if ctx.needs_input_grad[3]:
            raise RuntimeError("MSDeformAttnFusedFunction does not differentiate reference_points; "
                               "use MSDeformAttnFunction on the materialised sampling locations instead")
        grad_output = grad_output.contiguous()
        grad_value, grad_off, grad_logits = MSDA.ms_deform_attn_fused_backward(
            value, shapes, level_start, reference_points, sampling_offsets, attention_logits, grad_output,
            ctx.im2col_step)
        return grad_value, None, None, None, grad_off, grad_logits, None
