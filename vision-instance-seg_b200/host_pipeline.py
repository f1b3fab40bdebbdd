"""Forward+backward of the operator for operands that live in (pinned) HOST memory.

`HostPipeline.submit()` enqueues, for one call of the path, the host→device copies of (value,
sampling_locations, attention_weights, grad_output), `ms_deform_attn_forward` + `ms_deform_attn_backward`
through the extension-level API, and the device→host copies of (output, grad_value, grad_sampling_loc,
grad_attn_weight) — on three CUDA streams with multi-buffered device staging, so that the H2D of one piece, the kernels
of the previous one and the D2H of the one before overlap (PCIe is full duplex).  Images are independent in this
operator (output row (b, q) depends only on value[b], loc[b, q], attn[b, q]), so a call is cut along the batch into
`chunks` pieces that flow through the pipeline one after the other: the fill / drain bubble is one piece, not one
call, and the device staging is a fraction of the call's footprint.  Nothing is cached between calls: every submit
moves all of its bytes.  This is the end-to-end entry point `bench.py` times as `e2e`.
"""
from __future__ import annotations

import torch

from . import MultiScaleDeformableAttention as MSDA


class HostPipeline:
    """``fused=False`` (default): the upstream extension API, operands (value, sampling_locations, attention_weights,
    grad_output) -> results (output, grad_value, grad_sampling_loc, grad_attn_weight).

    ``fused=True``: the fused pre-op entry points (include/msda_b200.h ``msda_fused_forward / backward``), operands (value,
    reference_points, sampling_offsets, attn_logits, grad_output) -> results (output, grad_value, grad_sampling_offsets,
    grad_attn_logits).  With offsets / logits in the 16-bit value dtype -- what the two Linears emit under autocast -- the
    per-point auxiliary traffic over PCIe is 6 bytes each way instead of 12."""

    def __init__(self, spatial_shapes: torch.Tensor, level_start_index: torch.Tensor, device, im2col_step: int = 128,
                 depth: int = 3, chunks: int = 4, fused: bool = False):
        self.fused = bool(fused)
        self.device = torch.device(device)
        self.shapes = spatial_shapes.to(self.device, torch.int64).contiguous()
        self.lsi = level_start_index.to(self.device, torch.int64).contiguous()
        self.im2col_step = im2col_step
        self.depth = depth
        self.chunks = max(1, int(chunks))
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.slots = [None] * depth              # device staging of the four inputs
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, host_in, host_out):
        """host_in = (value, sampling_locations, attention_weights, grad_output) pinned CPU tensors -- or, with
        ``fused=True``, (value, reference_points, sampling_offsets, attn_logits, grad_output); host_out = (output,
        grad_value, grad of the location-like operand, grad of the weight-like operand) pinned CPU tensors to fill."""
        n = host_in[0].shape[0]
        pieces = min(self.chunks, n)
        base, rem = divmod(n, pieces)
        start = 0
        for i in range(pieces):
            stop = start + base + (1 if i < rem else 0)
            self._submit_piece([h[start:stop] for h in host_in], [h[start:stop] for h in host_out])
            start = stop

    def _submit_piece(self, host_in, host_out):
        k = self.count % self.depth
        first_use = self.count < self.depth
        self.count += 1
        with torch.cuda.device(self.device):
            if self.slots[k] is None or any(s.shape != h.shape or s.dtype != h.dtype for s, h in zip(self.slots[k], host_in)):
                self.slots[k] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_in]
            dev_in = self.slots[k]
            with torch.cuda.stream(self.s_in):
                if not first_use:
                    self.s_in.wait_event(self.ev_run[k])        # kernels of the previous user of this slot are done
                for d, h in zip(dev_in, host_in):
                    d.copy_(h, non_blocking=True)
                    self.h2d_bytes += h.numel() * h.element_size()
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                if self.fused:
                    v, ref, off, logits, go = dev_in
                    out = MSDA.ms_deform_attn_fused_forward(v, self.shapes, self.lsi, ref, off, logits, self.im2col_step)
                    gv, gl, ga = MSDA.ms_deform_attn_fused_backward(v, self.shapes, self.lsi, ref, off, logits, go,
                                                                    self.im2col_step)
                else:
                    v, loc, attn, go = dev_in
                    out = MSDA.ms_deform_attn_forward(v, self.shapes, self.lsi, loc, attn, self.im2col_step)
                    gv, gl, ga = MSDA.ms_deform_attn_backward(v, self.shapes, self.lsi, loc, attn, go, self.im2col_step)
                self.ev_run[k].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[k])
                for h, d in zip(host_out, (out, gv, gl, ga)):
                    d.record_stream(self.s_out)
                    h.copy_(d, non_blocking=True)
                    self.d2h_bytes += h.numel() * h.element_size()
                self.ev_out[k].record(self.s_out)

    def wait_all(self):
        """Make the caller's current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.s_in)
        cur.wait_stream(self.s_run)
        cur.wait_stream(self.s_out)

    def synchronize(self):
        self.s_in.synchronize()
        self.s_run.synchronize()
        self.s_out.synchronize()
