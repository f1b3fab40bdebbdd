"""Python face of the encoder-layer glue kernels (include/msda_encoder_b200.h): thin wrappers that allocate outputs with
torch and enqueue the CUDA kernels on the current stream.  No CPU path — CPU tensors raise."""
from __future__ import annotations

import torch

from . import _lib


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise RuntimeError("encoder glue kernels need contiguous CUDA tensors")


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def add_cast(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16(a + b) for float32 a, b of the same shape (with_pos_embed + the cast the Linears need)."""
    _need_cuda(a, b)
    if a.dtype != torch.float32 or b.dtype != torch.float32 or a.shape != b.shape:
        raise RuntimeError("add_cast expects two float32 tensors of the same shape")
    out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load_library().msda_enc_add_cast(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream(a))
    _lib.check(rc, "msda_enc_add_cast")
    return out


def add_layernorm_forward(x, delta16, gamma, beta, eps: float, want16: bool = True):
    """LayerNorm(x + delta16) over the last dim -> (y float32, y16 bfloat16 | None, mean, rstd)."""
    _need_cuda(x, delta16, gamma, beta)
    C = x.shape[-1]
    rows = x.numel() // C
    if x.dtype != torch.float32 or (delta16 is not None and (delta16.dtype != torch.bfloat16 or delta16.shape != x.shape)):
        raise RuntimeError("add_layernorm expects float32 x and bfloat16 delta of the same shape")
    y = torch.empty_like(x)
    y16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want16 else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load_library().msda_enc_add_layernorm_forward(x.data_ptr(), _ptr(delta16), gamma.data_ptr(), beta.data_ptr(),
                                                              y.data_ptr(), _ptr(y16), mean.data_ptr(), rstd.data_ptr(),
                                                              rows, C, float(eps), _stream(x))
    _lib.check(rc, "msda_enc_add_layernorm_forward")
    return y, y16, mean, rstd


def add_layernorm_backward(gy, gy16, x, delta16, mean, rstd, gamma, want_ddelta: bool = True):
    """-> (dx float32, ddelta16 bfloat16 | None, dgamma, dbeta); the incoming gradient is gy + float(gy16)."""
    _need_cuda(gy, gy16, x, delta16, mean, rstd, gamma)
    C = x.shape[-1]
    rows = x.numel() // C
    lib = _lib.load_library()
    dx = torch.empty_like(x)
    dd = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_ddelta else None
    dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
    nbytes = lib.msda_enc_add_layernorm_backward_scratch_bytes(C)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.msda_enc_add_layernorm_backward(_ptr(gy), _ptr(gy16), x.data_ptr(), _ptr(delta16), mean.data_ptr(),
                                                 rstd.data_ptr(), gamma.data_ptr(), dx.data_ptr(), _ptr(dd), dgamma.data_ptr(),
                                                 dbeta.data_ptr(), scratch.data_ptr(), nbytes, rows, C, _stream(x))
    _lib.check(rc, "msda_enc_add_layernorm_backward")
    return dx, dd, dgamma, dbeta


def colsum(g16: torch.Tensor, row_begin: int = 0, row_end: int | None = None) -> torch.Tensor:
    """float32 column sums of a bfloat16 (..., rows, C) tensor over rows [row_begin, row_end) of every leading entry."""
    _need_cuda(g16)
    if g16.dtype != torch.bfloat16 or g16.dim() < 2:
        raise RuntimeError("colsum expects a bfloat16 tensor with at least 2 dims")
    C = g16.shape[-1]
    if row_begin == 0 and row_end is None:
        batch, rpb = 1, g16.numel() // C
        row_end = rpb
    else:
        rpb = g16.shape[-2]
        batch = g16.numel() // (C * rpb)
        row_end = rpb if row_end is None else row_end
    lib = _lib.load_library()
    out = torch.empty(C, dtype=torch.float32, device=g16.device)
    nbytes = lib.msda_enc_colsum_scratch_bytes(C)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=g16.device)
    with torch.cuda.device(g16.device):
        rc = lib.msda_enc_colsum(g16.data_ptr(), out.data_ptr(), scratch.data_ptr(), nbytes, batch, rpb, row_begin, row_end, C,
                                 _stream(g16))
    _lib.check(rc, "msda_enc_colsum")
    return out


def relu_bwd_colsum(g16: torch.Tensor, h16: torch.Tensor) -> torch.Tensor:
    """In place g16 *= (h16 > 0); returns the float32 column sums of the masked gradient (bias gradient)."""
    _need_cuda(g16, h16)
    if g16.dtype != torch.bfloat16 or h16.dtype != torch.bfloat16 or g16.shape != h16.shape:
        raise RuntimeError("relu_bwd_colsum expects two bfloat16 tensors of the same shape")
    C = g16.shape[-1]
    rows = g16.numel() // C
    lib = _lib.load_library()
    out = torch.empty(C, dtype=torch.float32, device=g16.device)
    nbytes = lib.msda_enc_colsum_scratch_bytes(C)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=g16.device)
    with torch.cuda.device(g16.device):
        rc = lib.msda_enc_relu_bwd_colsum(g16.data_ptr(), h16.data_ptr(), out.data_ptr(), scratch.data_ptr(), nbytes, rows, C,
                                          _stream(g16))
    _lib.check(rc, "msda_enc_relu_bwd_colsum")
    return out
