"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — parity unpinned against reference tests.

CPU restatement of the reference's algorithm for the hot path.

What it follows.  The reference reaches the operator only through detectron2's ``build_model(cfg)``
(``/root/reference/training/maskdino/train_full.py:308``, ``evaluate.py:109``, ``visualize.py:251``)
after ``sys.path.insert(0, MASKDINO_PATH)`` (``train_full.py:15-16``); the operator's source is the
third-party, un-pinned IDEA-Research/MaskDINO checkout, module
``maskdino/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py`` (a descendant of
Deformable-DETR ``models/ops``).  Its published algorithm, restated here:

* ``ms_deform_attn_core_pytorch``: per level, ``F.grid_sample(value_l, 2*loc-1, bilinear, zeros,
  align_corners=False)``, multiply by the attention weights, sum over levels x points.
* ``ms_deform_attn_scalar_numpy``: the same arithmetic in the form the upstream CUDA kernel uses
  (``ms_deform_im2col_cuda.cuh``): pixel coords ``h_im = y*H - 0.5``, ``w_im = x*W - 0.5``; a sample is
  skipped unless ``h_im > -1 and w_im > -1 and h_im < H and w_im < W``; each of the four corners is
  zero-padded individually.  Pure-Python loops: small cases only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _shapes_list(value_spatial_shapes):
    if isinstance(value_spatial_shapes, torch.Tensor):
        return [(int(h), int(w)) for h, w in value_spatial_shapes.tolist()]
    return [(int(h), int(w)) for h, w in value_spatial_shapes]


def ms_deform_attn_core_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); shapes (L,2) rows (H,W); loc (N,Lq,M,L,P,2) last dim (x,y) in [0,1];
    attn (N,Lq,M,L,P) -> (N,Lq,M*D).  Differentiable (torch autograd supplies oracle gradients)."""
    N_, S_, M_, D_ = value.shape
    _, Lq_, M_, L_, P_, _ = sampling_locations.shape
    shapes = _shapes_list(value_spatial_shapes)
    value_list = value.split([H_ * W_ for H_, W_ in shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for lid_, (H_, W_) in enumerate(shapes):
        # N_, H_*W_, M_, D_ -> N_, H_*W_, M_*D_ -> N_, M_*D_, H_*W_ -> N_*M_, D_, H_, W_
        value_l_ = value_list[lid_].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, H_, W_)
        # N_, Lq_, M_, P_, 2 -> N_, M_, Lq_, P_, 2 -> N_*M_, Lq_, P_, 2
        sampling_grid_l_ = sampling_grids[:, :, :, lid_].transpose(1, 2).flatten(0, 1)
        # N_*M_, D_, Lq_, P_
        sampling_value_l_ = F.grid_sample(value_l_, sampling_grid_l_, mode="bilinear",
                                          padding_mode="zeros", align_corners=False)
        sampling_value_list.append(sampling_value_l_)
    # (N_, Lq_, M_, L_, P_) -> (N_, M_, Lq_, L_, P_) -> (N_*M_, 1, Lq_, L_*P_)
    attention_weights = attention_weights.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights).sum(-1)
    return output.view(N_, M_ * D_, Lq_).transpose(1, 2).contiguous()


def ms_deform_attn_oracle_grads(value, value_spatial_shapes, sampling_locations, attention_weights,
                                grad_output, dtype=torch.float64):
    """Forward + the three gradients through torch autograd of the restatement, computed in ``dtype``
    on the CPU.  Returns (out, grad_value, grad_sampling_loc, grad_attn_weight) as ``dtype`` tensors."""
    v = value.detach().to("cpu", dtype).requires_grad_(True)
    loc = sampling_locations.detach().to("cpu", dtype).requires_grad_(True)
    aw = attention_weights.detach().to("cpu", dtype).requires_grad_(True)
    out = ms_deform_attn_core_pytorch(v, value_spatial_shapes, loc, aw)
    out.backward(grad_output.detach().to("cpu", dtype))
    return out.detach(), v.grad, loc.grad, aw.grad


def ms_deform_attn_scalar_numpy(value, value_spatial_shapes, level_start_index, sampling_locations,
                                attention_weights, grad_output=None):
    """Scalar restatement in upstream-kernel form (float64 numpy, Python loops; tiny shapes only).

    Returns ``out`` or, when ``grad_output`` is given, ``(out, grad_value, grad_loc, grad_attn)`` using
    the upstream col2im formulas: ``grad_loc_x = W * sum_c(grad_w_weight_c * top_grad_c) * attn``,
    ``grad_loc_y = H * ...``, ``grad_attn = sum_c(bilinear_c * top_grad_c)`` and
    ``grad_value[corner] += corner_weight * attn * top_grad``."""
    v = np.asarray(value, dtype=np.float64)
    loc = np.asarray(sampling_locations, dtype=np.float64)
    aw = np.asarray(attention_weights, dtype=np.float64)
    shapes = _shapes_list(value_spatial_shapes)
    lsi = [int(x) for x in (level_start_index.tolist() if hasattr(level_start_index, "tolist") else level_start_index)]
    N_, S_, M_, D_ = v.shape
    _, Lq_, _, L_, P_, _ = loc.shape
    out = np.zeros((N_, Lq_, M_, D_))
    want_grad = grad_output is not None
    if want_grad:
        go = np.asarray(grad_output, dtype=np.float64).reshape(N_, Lq_, M_, D_)
        gv = np.zeros_like(v)
        gl = np.zeros_like(loc)
        ga = np.zeros_like(aw)
    for b in range(N_):
        for q in range(Lq_):
            for m in range(M_):
                for l, (H_, W_) in enumerate(shapes):
                    for p in range(P_):
                        x, y = loc[b, q, m, l, p]
                        a = aw[b, q, m, l, p]
                        h_im = y * H_ - 0.5
                        w_im = x * W_ - 0.5
                        if not (h_im > -1 and w_im > -1 and h_im < H_ and w_im < W_):
                            continue
                        h_low = int(np.floor(h_im))
                        w_low = int(np.floor(w_im))
                        lh, lw = h_im - h_low, w_im - w_low
                        hh, hw = 1 - lh, 1 - lw
                        corners = ((h_low, w_low, hh * hw, -hw, -hh), (h_low, w_low + 1, hh * lw, -lw, hh),
                                   (h_low + 1, w_low, lh * hw, hw, -lh), (h_low + 1, w_low + 1, lh * lw, lw, lh))
                        samp = np.zeros(D_)
                        gh = np.zeros(D_)
                        gw = np.zeros(D_)
                        for (hy, wx, wgt, dh, dw) in corners:
                            if 0 <= hy < H_ and 0 <= wx < W_:
                                vv = v[b, lsi[l] + hy * W_ + wx, m]
                                samp += wgt * vv
                                if want_grad:
                                    gh += dh * vv
                                    gw += dw * vv
                                    gv[b, lsi[l] + hy * W_ + wx, m] += wgt * a * go[b, q, m]
                        out[b, q, m] += a * samp
                        if want_grad:
                            ga[b, q, m, l, p] = float(np.dot(samp, go[b, q, m]))
                            gl[b, q, m, l, p, 0] = W_ * a * float(np.dot(gw, go[b, q, m]))
                            gl[b, q, m, l, p, 1] = H_ * a * float(np.dot(gh, go[b, q, m]))
    out = out.reshape(N_, Lq_, M_ * D_)
    if want_grad:
        return out, gv, gl, ga
    return out


def msdeformattn_preop_pytorch(reference_points, sampling_offsets, attention_logits, value_spatial_shapes):
    """The arithmetic upstream ``MSDeformAttn.forward`` (``modules/ms_deform_attn.py``) performs between its Linears
    and ``MSDeformAttnFunction.apply``: softmax of the attention logits over L*P, and
    ``sampling_locations = reference_points[:, :, None, :, None, :] + sampling_offsets / (W_l, H_l)`` for 2-d
    reference points, ``reference_points[..., :2] + sampling_offsets / P * reference_points[..., 2:] * 0.5`` for
    4-d reference boxes.  reference_points (N,Lq,L,2|4); sampling_offsets (N,Lq,M,L,P,2); logits (N,Lq,M,L*P)."""
    N_, Lq_, M_, L_, P_, _ = sampling_offsets.shape
    attention_weights = F.softmax(attention_logits.reshape(N_, Lq_, M_, L_ * P_), -1).view(N_, Lq_, M_, L_, P_)
    shapes = torch.as_tensor(_shapes_list(value_spatial_shapes), dtype=torch.long, device=sampling_offsets.device)
    if reference_points.shape[-1] == 2:
        offset_normalizer = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        sampling_locations = reference_points[:, :, None, :, None, :] \
            + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
    elif reference_points.shape[-1] == 4:
        sampling_locations = reference_points[:, :, None, :, None, :2] \
            + sampling_offsets / P_ * reference_points[:, :, None, :, None, 2:] * 0.5
    else:
        raise ValueError("Last dim of reference_points must be 2 or 4")
    return sampling_locations, attention_weights


def ms_deform_attn_fused_oracle_grads(value, value_spatial_shapes, reference_points, sampling_offsets,
                                      attention_logits, grad_output, dtype=torch.float64):
    """Oracle of the opt-in fused operator: pre-op + core, gradients by torch autograd, in ``dtype`` on the CPU.
    Returns (out, grad_value, grad_sampling_offsets, grad_attention_logits)."""
    v = value.detach().to("cpu", dtype).requires_grad_(True)
    ref = reference_points.detach().to("cpu", dtype)
    off = sampling_offsets.detach().to("cpu", dtype).requires_grad_(True)
    lg = attention_logits.detach().to("cpu", dtype).requires_grad_(True)
    loc, aw = msdeformattn_preop_pytorch(ref, off, lg, value_spatial_shapes)
    out = ms_deform_attn_core_pytorch(v, value_spatial_shapes, loc, aw)
    out.backward(grad_output.detach().to("cpu", dtype))
    return out.detach(), v.grad, off.grad, lg.grad
