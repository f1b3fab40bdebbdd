"""One stacked ``value_proj`` GEMM for all decoder layers.

Upstream's deformable decoder (MaskDINO ``maskdino/modeling/transformer_decoder/dino_decoder.py``: nine
``DeformableTransformerDecoderLayer``s, each with ``cross_attn = MSDeformAttn(...)``) hands the *same* encoder memory
``src (N, S, 256)`` to every layer, and every layer's ``MSDeformAttn.forward`` starts with its own
``value = value_proj(src)`` + ``masked_fill``: nine 256x256 GEMMs that each re-read the 178 MB memory (cfg4).

``share_value_proj([layer.cross_attn for layer in decoder.layers])`` makes those modules share one projection:

* the first ``cross_attn`` call of a forward pass runs a single Linear whose weight stacks the K ``value_proj`` weights
  (``torch.cat`` of the K parameters -- their names, shapes and ``state_dict`` entries are untouched, gradients flow
  back through the ``cat``) -> ``value_all (N, S, K, M, D)``: the memory is read once instead of K times;
* layer ``i`` samples the view ``value_all[:, :, i]`` *in place* through ``msda_forward_strided`` (pixel stride
  ``K*M*D``), and its backward writes ``grad_value`` through ``msda_backward_strided`` into the same view of one shared
  ``(N, S, K, M, D)`` gradient buffer, which is then handed to the stacked Linear's backward as it is: one
  ``(N*S, K*256) x (K*256, 256)`` GEMM for the gradient of the memory and one for the stacked weight instead of K each,
  and no per-layer copies in either direction.

Everything else of ``MSDeformAttn.forward`` (offsets, softmax, output_proj) is unchanged.  Opt-in; modules that were
not passed to ``share_value_proj`` behave exactly as upstream.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import MultiScaleDeformableAttention as MSDA


class SharedGradBuffer:
    """The ``(N, S, K, M, D)`` gradient of ``value_all`` that the K per-layer backward nodes of one backward pass fill
    slice by slice.  Allocated on first use in a pass, handed over (and forgotten) by ``_SplitStackedValue.backward``."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.written: set = set()

    def slice_for(self, like: torch.Tensor, layer: int) -> Optional[torch.Tensor]:
        """The full buffer if ``layer``'s slice has not been written in this pass yet, else None (a layer that is
        differentiated twice in one pass gets a private gradient and autograd adds the two)."""
        if layer in self.written:
            return None
        if self.buf is None or self.buf.shape != like.shape or self.buf.dtype != like.dtype or self.buf.device != like.device:
            self.buf = torch.empty_like(like)
            self.written = set()
        self.written.add(layer)
        return self.buf

    def take(self):
        buf, written = self.buf, self.written
        self.buf, self.written = None, set()
        return buf, written


class _SplitStackedValue(Function):
    """``value_all (N, S, K, M, D)`` -> its K views ``[:, :, i]``.  The backward recognises gradients that already are
    the matching views of one shared buffer (what ``MSDeformAttnStackedFunction`` returns) and passes that buffer on
    without touching it; anything else (a layer that was never sampled, a layer differentiated twice, a foreign
    consumer of a view) is copied / zero-filled into place, so the result is always the exact gradient."""

    @staticmethod
    def forward(ctx, value_all, shared: SharedGradBuffer):
        ctx.shared = shared
        ctx.shape = value_all.shape
        ctx.meta = (value_all.dtype, value_all.device)
        return tuple(value_all[:, :, i] for i in range(value_all.shape[2]))

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        buf, written = ctx.shared.take()
        if buf is None or tuple(buf.shape) != tuple(ctx.shape):
            buf = torch.empty(ctx.shape, dtype=ctx.meta[0], device=ctx.meta[1])
            written = set()
        for i, g in enumerate(grads):
            dst = buf[:, :, i]
            if g is None:
                dst.zero_()
            elif not (i in written and g.data_ptr() == dst.data_ptr() and g.stride() == dst.stride()
                      and g.dtype == dst.dtype):
                dst.copy_(g)
        return buf, None


class MSDeformAttnStackedFunction(Function):
    """``MSDeformAttnFunction`` on the view ``value_all[:, :, layer]`` of a stacked projection.  ``value_view`` is that
    view (it carries the autograd edge to ``_SplitStackedValue``); the kernels address it through ``value_all``."""

    @staticmethod
    def forward(ctx, value_view, value_all, shared, layer, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.shared = shared
        ctx.layer = int(layer)
        output = MSDA.ms_deform_attn_forward_stacked(value_all, ctx.layer, value_spatial_shapes, value_level_start_index,
                                                     sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value_all, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value_all, shapes, level_start, sampling_locations, attention_weights = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        buf = ctx.shared.slice_for(value_all, ctx.layer)
        if buf is None:        # second differentiation of this layer in one pass: private, dense gradient
            grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
                value_all[:, :, ctx.layer].contiguous(), shapes, level_start, sampling_locations, attention_weights,
                grad_output, ctx.im2col_step)
        else:
            grad_loc, grad_attn = MSDA.ms_deform_attn_backward_stacked(
                value_all, ctx.layer, shapes, level_start, sampling_locations, attention_weights, grad_output, buf,
                ctx.im2col_step)
            grad_value = buf[:, :, ctx.layer]
        return grad_value, None, None, None, None, None, grad_loc, grad_attn, None


class StackedValueProj:
    """Shared state of K ``MSDeformAttn`` modules whose ``value_proj`` runs as one stacked GEMM (see module docstring)."""

    def __init__(self, modules: Sequence[torch.nn.Module]):
        mods = list(modules)
        if len(mods) < 2:
            raise ValueError("share_value_proj needs at least two MSDeformAttn modules")
        d, h = mods[0].d_model, mods[0].n_heads
        for m in mods:
            if m.d_model != d or m.n_heads != h:
                raise ValueError("all modules must have the same d_model and n_heads")
        self.modules = mods
        self.K, self.d_model, self.n_heads = len(mods), d, h
        self._clear()

    def _clear(self):
        self._src = self._mask = self._views = self._value_all = self._shared = self._key = None
        self._served = 0
        self._last_index = -1

    @staticmethod
    def _tensor_key(t: Optional[torch.Tensor]):
        """Identity of a tensor's *contents and autograd origin*: the decoder layers call
        ``cross_attn(..., memory.transpose(0, 1), ...)``, i.e. every layer presents a fresh view object of the same
        memory, which must hit the cache; an in-place update (version bump) or another base tensor must not."""
        if t is None:
            return None
        base = t._base if t._base is not None else t
        return (id(base), t.data_ptr(), t._version, tuple(t.shape), t.stride(), t.dtype, t.requires_grad)

    def project(self, input_flatten: torch.Tensor, input_padding_mask: Optional[torch.Tensor]):
        """-> (views, value_all, shared): the K per-layer value views ``(N, S, M, D)``, the stacked tensor they alias and
        the gradient buffer state of this forward pass."""
        N, S, _ = input_flatten.shape
        weight = torch.cat([m.value_proj.weight for m in self.modules], 0)
        bias = torch.cat([m.value_proj.bias for m in self.modules], 0)
        # The decoder hands over `memory.transpose(0, 1)`: a strided view.  F.linear on it would fall back to matmul plus a
        # separate broadcast add of the bias over the whole (N, S, K*C) output (1.8 ms at cfg4); one fused transpose + cast
        # to a dense 2-d operand keeps the bias in the GEMM epilogue (addmm).
        dtype = input_flatten.dtype
        if input_flatten.is_cuda and torch.is_autocast_enabled("cuda"):
            dtype = torch.get_autocast_dtype("cuda")
        x2d = input_flatten.to(dtype=dtype, memory_format=torch.contiguous_format).reshape(N * S, -1)
        value_all = F.linear(x2d, weight, bias).view(N, S, -1)
        if input_padding_mask is not None:
            value_all = value_all.masked_fill(input_padding_mask[..., None], float(0))
        value_all = value_all.view(N, S, self.K, self.n_heads, self.d_model // self.n_heads)
        shared = SharedGradBuffer()
        views = _SplitStackedValue.apply(value_all, shared)
        return views, value_all, shared

    def value_for(self, index: int, input_flatten: torch.Tensor, input_padding_mask: Optional[torch.Tensor]):
        """Layer ``index``'s (view, value_all, shared); the stacked GEMM runs when a forward pass presents a memory tensor
        (or mask) that differs from the cached one, and the cache is dropped once all K layers have been served."""
        # the key also covers the projection weights (an optimizer step in between bumps their version counters), and a
        # layer index that does not advance means a new forward pass began -- e.g. after a pass that served fewer than K
        # layers (exception, early exit, pruned layers): never hand out a projection computed with stale weights or
        # whose autograd graph has been freed
        wkey = tuple(t._version for m in self.modules for t in (m.value_proj.weight, m.value_proj.bias))
        key = (self._tensor_key(input_flatten), self._tensor_key(input_padding_mask), torch.is_grad_enabled(), wkey)
        if self._views is None or key != self._key or index <= self._last_index:
            self._clear()
            self._views, self._value_all, self._shared = self.project(input_flatten, input_padding_mask)
            self._src, self._mask = input_flatten, input_padding_mask        # keep them alive: the key holds addresses
            self._key = key
        out = (self._views[index], self._value_all, self._shared)
        self._last_index = index
        self._served += 1
        if self._served >= self.K:
            self._clear()
        return out


def share_value_proj(modules: Sequence[torch.nn.Module]) -> StackedValueProj:
    """Make the given ``MSDeformAttn`` modules (the decoder layers' ``cross_attn``, in layer order) share one stacked
    ``value_proj`` GEMM per forward pass.  Returns the shared state; ``unshare_value_proj`` undoes it."""
    proj = StackedValueProj(modules)
    for i, m in enumerate(proj.modules):
        m._stacked_value = (proj, i)
    return proj


def unshare_value_proj(modules: Sequence[torch.nn.Module]) -> None:
    for m in modules:
        m._stacked_value = None
