"""One pixel-decoder encoder layer as a single autograd node.

Same parameters, same mathematics as ``MSDeformAttnTransformerEncoderLayer.forward`` under bf16 autocast — fp32
residual stream and LayerNorms, bf16 operands into every GEMM and into the sampling kernels — but with the memory-bound
glue done by the kernels of ``include/msda_encoder_b200.h`` and with a hand-written backward:

    q16   = bf16(src + pos)                                   msda_enc_add_cast
    off, logits = Linear(q16);  value = Linear(src16)         library GEMMs (bf16 in, bf16 out, fp32 accumulate)
    attn  = MSDeformAttn core on (value, ref, off, logits)    msda_fused_forward (softmax / locations folded in)
    x1    = LayerNorm1(src + Linear(attn))                    GEMM + msda_enc_add_layernorm_forward (fp32 + bf16 copies)
    h     = relu(Linear1(x1_16))                              cuBLASLt bias+ReLU epilogue (torch._addmm_activation)
    x2    = LayerNorm2(x1 + Linear2(h))                       GEMM + msda_enc_add_layernorm_forward

The backward walks the same chain in reverse: LayerNorm backward with the residual and bf16 gradients summed in the
kernel, ReLU backward fused with the bias gradient, bias gradients as deterministic column sums, weight gradients as
bf16 GEMMs with fp32 output.  The layer hands the next one both the fp32 residual stream and its bf16 copy, so no cast
kernel runs between layers; the gradient of the bf16 copy carries the (query + value) paths back to the previous
LayerNorm backward, which adds it to the fp32 residual gradient on the fly.

Dropout must be inactive (p = 0 or eval mode — MaskDINO's encoder uses dropout 0.0) and the activation ReLU; otherwise the
stock layer code runs.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import MultiScaleDeformableAttention as MSDA
from .. import encoder_ops as ops

_BF16 = torch.bfloat16


def _linear_relu(bias16: torch.Tensor, x16: torch.Tensor, w16_t: torch.Tensor) -> torch.Tensor:
    """relu(x @ w^T + b) with the bias + ReLU in the GEMM epilogue (cuBLASLt) where this torch build exposes it."""
    fused = getattr(torch, "_addmm_activation", None)
    if fused is not None:
        try:
            return fused(bias16, x16, w16_t, use_gelu=False)
        except (TypeError, RuntimeError):
            pass
    return torch.relu_(torch.addmm(bias16, x16, w16_t))


def _mm_f32(a16: torch.Tensor, b16: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 -> fp32 GEMM (weight gradients): fp32 output straight from the accumulator where torch offers it."""
    try:
        return torch.mm(a16, b16, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.mm(a16, b16).float()


class FusedEncoderLayerFunction(Function):
    """forward(src, src16, pos, level_embed, reference_points, spatial_shapes, level_start_index, padding_mask, cfg,
    *16 parameters) -> (x2 float32, x2_16 bfloat16 | None).  ``cfg`` = dict(heads, levels, points, level_bounds, eps1, eps2,
    im2col_step, want16)."""

    @staticmethod
    def forward(ctx, src, src16, pos, level_embed, reference_points, spatial_shapes, level_start_index, padding_mask, cfg,
                so_w, so_b, aw_w, aw_b, v_w, v_b, o_w, o_b, n1_w, n1_b, l1_w, l1_b, l2_w, l2_b, n2_w, n2_b):
        N, S, C = src.shape
        T = N * S
        M, L, P = cfg["heads"], cfg["levels"], cfg["points"]
        D = C // M
        step = cfg["im2col_step"]
        w16 = [w.to(_BF16) for w in (so_w, so_b, aw_w, aw_b, v_w, v_b, o_w, o_b, l1_w, l1_b, l2_w, l2_b)]
        so_w16, so_b16, aw_w16, aw_b16, v_w16, v_b16, o_w16, o_b16, l1_w16, l1_b16, l2_w16, l2_b16 = w16
        src2d = src.view(T, C)
        q16 = ops.add_cast(src, pos).view(T, C)
        off = torch.addmm(so_b16, q16, so_w16.t())
        lg = torch.addmm(aw_b16, q16, aw_w16.t())
        val = torch.addmm(v_b16, src16.view(T, C), v_w16.t())
        if padding_mask is not None:
            val.view(N, S, C).masked_fill_(padding_mask[..., None], 0.0)
        attn = MSDA.ms_deform_attn_fused_forward(val.view(N, S, M, D), spatial_shapes, level_start_index, reference_points,
                                                 off.view(N, S, M, L, P, 2), lg.view(N, S, M, L * P), step).view(T, C)
        o = torch.addmm(o_b16, attn, o_w16.t())
        x1, x1_16, mean1, rstd1 = ops.add_layernorm_forward(src2d, o, n1_w, n1_b, cfg["eps1"])
        h = _linear_relu(l1_b16, x1_16, l1_w16.t())
        f = torch.addmm(l2_b16, h, l2_w16.t())
        x2, x2_16, mean2, rstd2 = ops.add_layernorm_forward(x1, f, n2_w, n2_b, cfg["eps2"], want16=cfg["want16"])
        ctx.cfg = cfg
        ctx.shape = (N, S, C)
        ctx.has_mask = padding_mask is not None
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(src, src16, q16, off, lg, val, attn, o, x1, x1_16, mean1, rstd1, h, f, mean2, rstd2,
                              reference_points, spatial_shapes, level_start_index, padding_mask,
                              so_w16, aw_w16, v_w16, o_w16, l1_w16, l2_w16, n1_w, n2_w)
        out16 = x2_16.view(N, S, C) if x2_16 is not None else None
        if out16 is None:
            return x2.view(N, S, C), None
        return x2.view(N, S, C), out16

    @staticmethod
    @once_differentiable
    def backward(ctx, g_x2, g_x2_16):
        (src, src16, q16, off, lg, val, attn, o, x1, x1_16, mean1, rstd1, h, f, mean2, rstd2, ref, shapes, lsi, mask,
         so_w16, aw_w16, v_w16, o_w16, l1_w16, l2_w16, n1_w, n2_w) = ctx.saved_tensors
        cfg = ctx.cfg
        N, S, C = ctx.shape
        T = N * S
        M, L, P = cfg["heads"], cfg["levels"], cfg["points"]
        D = C // M
        if g_x2 is None and g_x2_16 is None:
            return (None,) * 25
        g2 = g_x2.contiguous().view(T, C) if g_x2 is not None else None
        g2_16 = g_x2_16.contiguous().view(T, C) if g_x2_16 is not None else None
        # LayerNorm2 / linear2 / ReLU / linear1
        dx1, df, dn2_w, dn2_b = ops.add_layernorm_backward(g2, g2_16, x1, f, mean2, rstd2, n2_w)
        dl2_b = ops.colsum(df)
        dl2_w = _mm_f32(df.t(), h)
        dh = torch.mm(df, l2_w16)
        dl1_b = ops.relu_bwd_colsum(dh, h)
        dl1_w = _mm_f32(dh.t(), x1_16)
        dx1_16 = torch.mm(dh, l1_w16)
        del dh
        # LayerNorm1 / output_proj
        dsrc, do, dn1_w, dn1_b = ops.add_layernorm_backward(dx1, dx1_16, src.view(T, C), o, mean1, rstd1, n1_w)
        do_b = ops.colsum(do)
        do_w = _mm_f32(do.t(), attn)
        dattn = torch.mm(do, o_w16)
        # sampling core
        dval, doff, dlg = MSDA.ms_deform_attn_fused_backward(val.view(N, S, M, D), shapes, lsi, ref,
                                                             off.view(N, S, M, L, P, 2), lg.view(N, S, M, L * P),
                                                             dattn.view(N, S, C), cfg["im2col_step"])
        dval = dval.view(T, C)
        doff = doff.view(T, M * L * P * 2)
        dlg = dlg.view(T, M * L * P)
        if ctx.has_mask:
            dval.view(N, S, C).masked_fill_(mask[..., None], 0.0)
        dv_b = ops.colsum(dval)
        dv_w = _mm_f32(dval.t(), src16.view(T, C))
        dso_b = ops.colsum(doff)
        dso_w = _mm_f32(doff.t(), q16)
        daw_b = ops.colsum(dlg)
        daw_w = _mm_f32(dlg.t(), q16)
        # gradient of q = src + pos (bf16), then the value path on top: total gradient of the bf16 view of src
        dq16 = torch.addmm(torch.mm(doff, so_w16), dlg, aw_w16)
        dlevel = None
        if ctx.needs_input_grad[3]:
            dq3 = dq16.view(N, S, C)
            dlevel = torch.stack([ops.colsum(dq3, a, b) for a, b in cfg["level_bounds"]], 0)
        dsrc16 = torch.addmm(dq16, dval, v_w16)
        return (dsrc.view(N, S, C), dsrc16.view(N, S, C), None, dlevel, None, None, None, None, None,
                dso_w, dso_b, daw_w, daw_b, dv_w, dv_b, do_w, do_b, dn1_w, dn1_b, dl1_w, dl1_b, dl2_w, dl2_b, dn2_w, dn2_b)


def fused_layer_supported(layer, src, reference_points) -> bool:
    """The fused node covers: CUDA float32 residual stream, d_model a multiple of 128 (<= 1024), d_ffn a multiple of 32,
    head dim 16/32/64/128, 2-d reference points, ReLU, inactive dropout."""
    import torch.nn.functional as F
    attn = layer.self_attn
    C = attn.d_model
    if not (src.is_cuda and src.dtype == torch.float32 and C % 128 == 0 and C <= 1024):
        return False
    if layer.linear1.out_features % 32 != 0 or layer.activation is not F.relu:
        return False
    if reference_points.shape[-1] != 2 or (torch.is_grad_enabled() and reference_points.requires_grad):
        return False
    if (C // attn.n_heads) not in (16, 32, 64, 128):
        return False
    if layer.training and any(d.p > 0 for d in (layer.dropout1, layer.dropout2, layer.dropout3)):
        return False
    return True


def fused_layer_forward(layer, src, src16, pos, level_embed, reference_points, spatial_shapes, level_start_index,
                        padding_mask, level_bounds, want16=True):
    attn = layer.self_attn
    cfg = dict(heads=attn.n_heads, levels=attn.n_levels, points=attn.n_points, level_bounds=tuple(level_bounds),
               eps1=layer.norm1.eps, eps2=layer.norm2.eps, im2col_step=attn.im2col_step, want16=want16)
    return FusedEncoderLayerFunction.apply(
        src, src16, pos, level_embed, reference_points, spatial_shapes, level_start_index, padding_mask, cfg,
        attn.sampling_offsets.weight, attn.sampling_offsets.bias, attn.attention_weights.weight, attn.attention_weights.bias,
        attn.value_proj.weight, attn.value_proj.bias, attn.output_proj.weight, attn.output_proj.bias,
        layer.norm1.weight, layer.norm1.bias, layer.linear1.weight, layer.linear1.bias,
        layer.linear2.weight, layer.linear2.bias, layer.norm2.weight, layer.norm2.bias)
