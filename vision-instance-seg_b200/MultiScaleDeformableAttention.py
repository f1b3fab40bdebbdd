"""Stand-in for the upstream pybind11 extension module ``MultiScaleDeformableAttention``.

Upstream (IDEA-Research/MaskDINO ``maskdino/modeling/pixel_decoder/ops/src/vision.cpp``) exports
``ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
im2col_step) -> Tensor`` and ``ms_deform_attn_backward(..., grad_output, im2col_step) -> [Tensor x3]``;
``functions/ms_deform_attn_func.py`` imports the module as ``MSDA``.  The same two functions, with the
same argument meaning and the same error behaviour (``RuntimeError`` for non-contiguous / non-CUDA
tensors and for a batch that ``im2col_step`` does not divide), are provided here on top of the C ABI.

Differences that are deliberate and documented in INTEGRATION.md:
* 16-bit ``value`` (bfloat16 / float16) is accepted (upstream dispatches float / double only);
  ``sampling_loc`` / ``attn_weight`` are consumed in float32 in that case.
* kernel launch failures raise instead of being printf-ed and swallowed.
"""
from __future__ import annotations

import os

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.MSDA_F32, torch.float64: _lib.MSDA_F64,
           torch.bfloat16: _lib.MSDA_BF16, torch.float16: _lib.MSDA_F16}

#: backward flags (see include/msda_b200.h); MSDA_B200_FP32_ACCUM=1 forces fp32 accumulation of grad_value for
#: 16-bit values, MSDA_B200_ACCUM_DEPTH=<n> overrides the fp16 bucket depth
backward_flags = (_lib.MSDA_BWD_GRAD_VALUE_FP32_ACCUM if os.environ.get("MSDA_B200_FP32_ACCUM") == "1"
                  else _lib.MSDA_BWD_DEFAULT) | _lib.accum_depth_flag(int(os.environ.get("MSDA_B200_ACCUM_DEPTH", "0"))) \
    | (_lib.MSDA_BWD_NO_SPARSE_DIRECT if os.environ.get("MSDA_B200_NO_SPARSE_DIRECT") == "1" else 0) \
    | (_lib.MSDA_BWD_NO_CLUSTER_GUARD if os.environ.get("MSDA_B200_NO_CLUSTER_GUARD") == "1" else 0)


def _require(t: torch.Tensor, name: str) -> None:
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _aux_dtype(value: torch.Tensor) -> torch.dtype:
    return torch.float64 if value.dtype == torch.float64 else torch.float32


def _dims(value, spatial_shapes, sampling_loc):
    if value.dim() != 4 or sampling_loc.dim() != 6:
        raise RuntimeError("value must be (N, S, M, D) and sampling_loc (N, Lq, M, L, P, 2)")
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = sampling_loc.shape
    if M2 != M or two != 2 or spatial_shapes.shape[0] != L or sampling_loc.shape[0] != N:
        raise RuntimeError("inconsistent shapes between value, spatial_shapes and sampling_loc")
    return N, S, M, D, Lq, L, P


def _meta(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.int64 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.int64).contiguous()
    return t


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    ext = _lib.torch_extension()
    if ext is not None:        # compiled torch extension over the same C ABI (same checks, same errors, less host time)
        return ext.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                          int(im2col_step))
    for t, n in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")):
        _require(t, n)
    if value.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn_forward not implemented for '{value.dtype}'")
    N, S, M, D, Lq, L, P = _dims(value, spatial_shapes, sampling_loc)
    aux = _aux_dtype(value)
    loc = sampling_loc if sampling_loc.dtype == aux else sampling_loc.to(aux)
    attn = attn_weight if attn_weight.dtype == aux else attn_weight.to(aux)
    shapes = _meta(spatial_shapes, value.device)
    lsi = _meta(level_start_index, value.device)
    lib = _lib.load_library()
    with torch.cuda.device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        if out.numel() == 0:
            return out
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.msda_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                              out.data_ptr(), N, S, M, D, Lq, L, P, _DTYPES[value.dtype], int(im2col_step), stream)
    _lib.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    ext = _lib.torch_extension()
    if ext is not None:
        return ext.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                           grad_output, int(im2col_step), int(backward_flags))
    for t, n in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight"), (grad_output, "grad_output")):
        _require(t, n)
    if value.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn_backward not implemented for '{value.dtype}'")
    N, S, M, D, Lq, L, P = _dims(value, spatial_shapes, sampling_loc)
    aux = _aux_dtype(value)
    loc = sampling_loc if sampling_loc.dtype == aux else sampling_loc.to(aux)
    attn = attn_weight if attn_weight.dtype == aux else attn_weight.to(aux)
    go = grad_output if grad_output.dtype == value.dtype else grad_output.to(value.dtype)
    shapes = _meta(spatial_shapes, value.device)
    lsi = _meta(level_start_index, value.device)
    lib = _lib.load_library()
    code = _DTYPES[value.dtype]
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)
        grad_loc = torch.empty(sampling_loc.shape, dtype=aux, device=value.device)
        grad_attn = torch.empty(attn_weight.shape, dtype=aux, device=value.device)
        if grad_loc.numel() == 0 or value.numel() == 0:
            return [grad_value.zero_(), grad_loc.zero_(), grad_attn.zero_()]
        nbytes = lib.msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, code, backward_flags)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=value.device) if nbytes else None
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.msda_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                               go.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
                               scratch.data_ptr() if scratch is not None else None, nbytes,
                               N, S, M, D, Lq, L, P, code, int(im2col_step), backward_flags, stream)
    _lib.check(rc, "ms_deform_attn_backward")
    if grad_loc.dtype != sampling_loc.dtype:
        grad_loc = grad_loc.to(sampling_loc.dtype)
    if grad_attn.dtype != attn_weight.dtype:
        grad_attn = grad_attn.to(attn_weight.dtype)
    return [grad_value, grad_loc, grad_attn]


# ---------------------------------------------------------------------------------------------------
# Strided value views (opt-in; SURVEY.md §8f rank 2: one stacked ``value_proj`` GEMM for all decoder layers).
# ``value_all`` is a contiguous (N, S, K, M, D) tensor -- the output of a Linear whose weight stacks the K layers'
# ``value_proj`` weights -- and ``layer`` picks the view [:, :, layer] that the kernels read in place; the backward
# writes grad_value into the same view of ``grad_value_all``.  Not part of the upstream extension.
# ---------------------------------------------------------------------------------------------------
def _stacked_dims(value_all, layer, spatial_shapes, sampling_loc):
    if value_all.dim() != 5 or sampling_loc.dim() != 6:
        raise RuntimeError("value_all must be (N, S, K, M, D) and sampling_loc (N, Lq, M, L, P, 2)")
    N, S, K, M, D = value_all.shape
    if not 0 <= int(layer) < K:
        raise RuntimeError(f"layer {layer} out of range for {K} stacked projections")
    _, Lq, M2, L, P, two = sampling_loc.shape
    if M2 != M or two != 2 or spatial_shapes.shape[0] != L or sampling_loc.shape[0] != N:
        raise RuntimeError("inconsistent shapes between value_all, spatial_shapes and sampling_loc")
    return N, S, K, M, D, Lq, L, P


def ms_deform_attn_forward_stacked(value_all, layer, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                   im2col_step):
    """``ms_deform_attn_forward`` on ``value_all[:, :, layer]`` without materialising the view."""
    for t, n in ((value_all, "value_all"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")):
        _require(t, n)
    if value_all.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn_forward_stacked not implemented for '{value_all.dtype}'")
    N, S, K, M, D, Lq, L, P = _stacked_dims(value_all, layer, spatial_shapes, sampling_loc)
    aux = _aux_dtype(value_all)
    loc = sampling_loc if sampling_loc.dtype == aux else sampling_loc.to(aux)
    attn = attn_weight if attn_weight.dtype == aux else attn_weight.to(aux)
    shapes = _meta(spatial_shapes, value_all.device)
    lsi = _meta(level_start_index, value_all.device)
    lib = _lib.load_library()
    with torch.cuda.device(value_all.device):
        out = torch.empty((N, Lq, M * D), dtype=value_all.dtype, device=value_all.device)
        if out.numel() == 0:
            return out
        stream = torch.cuda.current_stream().cuda_stream
        base = value_all.data_ptr() + int(layer) * M * D * value_all.element_size()
        rc = lib.msda_forward_strided(base, K * M * D, shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                      out.data_ptr(), N, S, M, D, Lq, L, P, _DTYPES[value_all.dtype], int(im2col_step),
                                      stream)
    _lib.check(rc, "ms_deform_attn_forward_stacked")
    return out


def ms_deform_attn_backward_stacked(value_all, layer, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                    grad_output, grad_value_all, im2col_step):
    """``ms_deform_attn_backward`` on ``value_all[:, :, layer]``; ``grad_value_all[:, :, layer]`` is fully overwritten in
    place (every other layer's slice is left untouched).  -> [grad_sampling_loc, grad_attn_weight]"""
    for t, n in ((value_all, "value_all"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight"), (grad_output, "grad_output"),
                 (grad_value_all, "grad_value_all")):
        _require(t, n)
    if value_all.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn_backward_stacked not implemented for '{value_all.dtype}'")
    if grad_value_all.shape != value_all.shape or grad_value_all.dtype != value_all.dtype:
        raise RuntimeError("grad_value_all must match value_all in shape and dtype")
    N, S, K, M, D, Lq, L, P = _stacked_dims(value_all, layer, spatial_shapes, sampling_loc)
    aux = _aux_dtype(value_all)
    loc = sampling_loc if sampling_loc.dtype == aux else sampling_loc.to(aux)
    attn = attn_weight if attn_weight.dtype == aux else attn_weight.to(aux)
    go = grad_output if grad_output.dtype == value_all.dtype else grad_output.to(value_all.dtype)
    shapes = _meta(spatial_shapes, value_all.device)
    lsi = _meta(level_start_index, value_all.device)
    lib = _lib.load_library()
    code = _DTYPES[value_all.dtype]
    flags = backward_flags & ~_lib.MSDA_BWD_GRAD_VALUE_FP32_ACCUM      # strided grad_value: default accumulation only
    with torch.cuda.device(value_all.device):
        grad_loc = torch.empty(sampling_loc.shape, dtype=aux, device=value_all.device)
        grad_attn = torch.empty(attn_weight.shape, dtype=aux, device=value_all.device)
        if grad_loc.numel() == 0 or value_all.numel() == 0:
            grad_value_all[:, :, int(layer)].zero_()
            return [grad_loc.zero_(), grad_attn.zero_()]
        nbytes = lib.msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, code, flags)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=value_all.device) if nbytes else None
        stream = torch.cuda.current_stream().cuda_stream
        off = int(layer) * M * D * value_all.element_size()
        rc = lib.msda_backward_strided(value_all.data_ptr() + off, K * M * D, shapes.data_ptr(), lsi.data_ptr(),
                                       loc.data_ptr(), attn.data_ptr(), go.data_ptr(),
                                       grad_value_all.data_ptr() + off, K * M * D, grad_loc.data_ptr(), grad_attn.data_ptr(),
                                       scratch.data_ptr() if scratch is not None else None, nbytes,
                                       N, S, M, D, Lq, L, P, code, int(im2col_step), flags, stream)
    _lib.check(rc, "ms_deform_attn_backward_stacked")
    if grad_loc.dtype != sampling_loc.dtype:
        grad_loc = grad_loc.to(sampling_loc.dtype)
    if grad_attn.dtype != attn_weight.dtype:
        grad_attn = grad_attn.to(attn_weight.dtype)
    return [grad_loc, grad_attn]


# ---------------------------------------------------------------------------------------------------
# Fused pre-op (opt-in).  Not part of the upstream extension: the two functions below fold
# the softmax over L*P and the sampling-location arithmetic of upstream ``MSDeformAttn.forward`` into the kernels.
# ---------------------------------------------------------------------------------------------------
def fused_supported(value: torch.Tensor, reference_points: torch.Tensor) -> bool:
    """True when the fused kernels cover this call (vector kernels: head dim 16/32/64/128, fp32/bf16/fp16
    values; reference points of width 2 or 4)."""
    if value.dtype not in _DTYPES or value.dim() != 4 or reference_points.shape[-1] not in (2, 4):
        return False
    return bool(_lib.load_library().msda_fused_supported(int(value.shape[-1]), _DTYPES[value.dtype]))


def _fused_aux(value, sampling_offsets, attn_logits):
    """Offsets / logits are consumed as they are when both are float32, or both already in the 16-bit value dtype (what the
    Linears emit inside torch.autocast); any other combination is brought to float32."""
    if value.dtype in (torch.bfloat16, torch.float16) and sampling_offsets.dtype == value.dtype \
            and attn_logits.dtype == value.dtype:
        return sampling_offsets, attn_logits, value.dtype
    off = sampling_offsets if sampling_offsets.dtype == torch.float32 else sampling_offsets.float()
    logits = attn_logits if attn_logits.dtype == torch.float32 else attn_logits.float()
    return off, logits, torch.float32


def _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits):
    if value.dim() != 4 or sampling_offsets.dim() != 6 or reference_points.dim() != 4:
        raise RuntimeError("value must be (N, S, M, D), sampling_offsets (N, Lq, M, L, P, 2), reference_points (N, Lq, L, 2|4)")
    N, S, M, D = value.shape
    N2, Lq, M2, L, P, two = sampling_offsets.shape
    if (M2 != M or two != 2 or N2 != N or spatial_shapes.shape[0] != L or attn_logits.numel() != N * Lq * M * L * P
            or tuple(reference_points.shape[:3]) != (N, Lq, L) or reference_points.shape[3] not in (2, 4)):
        raise RuntimeError("inconsistent shapes between value, spatial_shapes, reference_points, sampling_offsets and attn_logits")
    return N, S, M, D, Lq, L, P, int(reference_points.shape[3])


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                 attn_logits, im2col_step):
    """output = MSDeformAttn core applied to softmax(attn_logits) and reference_points (+) sampling_offsets."""
    for t, n in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (reference_points, "reference_points"), (sampling_offsets, "sampling_offsets"), (attn_logits, "attn_logits")):
        _require(t, n)
    if value.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn_fused_forward not implemented for '{value.dtype}'")
    N, S, M, D, Lq, L, P, R = _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits)
    ref = reference_points if reference_points.dtype == torch.float32 else reference_points.float()
    off, logits, aux = _fused_aux(value, sampling_offsets, attn_logits)
    shapes = _meta(spatial_shapes, value.device)
    lsi = _meta(level_start_index, value.device)
    lib = _lib.load_library()
    with torch.cuda.device(value.device):
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        if out.numel() == 0:
            return out
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.msda_fused_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), ref.data_ptr(), R, off.data_ptr(),
                                    logits.data_ptr(), out.data_ptr(), N, S, M, D, Lq, L, P, _DTYPES[value.dtype],
                                    _DTYPES[aux], int(im2col_step), stream)
    _lib.check(rc, "ms_deform_attn_fused_forward")
    return out


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                                  attn_logits, grad_output, im2col_step):
    """-> [grad_value, grad_sampling_offsets, grad_attn_logits] (reference_points gets no gradient here)."""
    for t, n in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (reference_points, "reference_points"), (sampling_offsets, "sampling_offsets"), (attn_logits, "attn_logits"),
                 (grad_output, "grad_output")):
        _require(t, n)
    if value.dtype not in _DTYPES:
        raise RuntimeError(f"ms_deform_attn_fused_backward not implemented for '{value.dtype}'")
    N, S, M, D, Lq, L, P, R = _fused_dims(value, spatial_shapes, reference_points, sampling_offsets, attn_logits)
    ref = reference_points if reference_points.dtype == torch.float32 else reference_points.float()
    off, logits, aux = _fused_aux(value, sampling_offsets, attn_logits)
    go = grad_output if grad_output.dtype == value.dtype else grad_output.to(value.dtype)
    shapes = _meta(spatial_shapes, value.device)
    lsi = _meta(level_start_index, value.device)
    lib = _lib.load_library()
    code = _DTYPES[value.dtype]
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)
        grad_off = torch.empty(sampling_offsets.shape, dtype=aux, device=value.device)
        grad_logits = torch.empty(attn_logits.shape, dtype=aux, device=value.device)
        if grad_off.numel() == 0 or value.numel() == 0:
            return [grad_value.zero_(), grad_off.zero_(), grad_logits.zero_()]
        nbytes = lib.msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, code, backward_flags)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=value.device) if nbytes else None
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.msda_fused_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), ref.data_ptr(), R, off.data_ptr(),
                                     logits.data_ptr(), go.data_ptr(), grad_value.data_ptr(), grad_off.data_ptr(),
                                     grad_logits.data_ptr(), scratch.data_ptr() if scratch is not None else None, nbytes,
                                     N, S, M, D, Lq, L, P, code, _DTYPES[aux], int(im2col_step), backward_flags, stream)
    _lib.check(rc, "ms_deform_attn_fused_backward")
    if grad_off.dtype != sampling_offsets.dtype:
        grad_off = grad_off.to(sampling_offsets.dtype)
    if grad_logits.dtype != attn_logits.dtype:
        grad_logits = grad_logits.to(attn_logits.dtype)
    return [grad_value, grad_off, grad_logits]
