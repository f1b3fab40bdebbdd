"""Opt-in tiled kernels of the dense call site (csrc/msda_tiled.cuh) against the CPU oracle.

Same operator, same C ABI (msda_forward / msda_backward through the upstream-shaped Python module); only
``msda_set_tiled_mode(1)`` differs.  Cases cover what the tiling logic branches on: pyramids and shapes that do not nest,
levels narrower than a tile, L*P that does not divide the CTA, fp16 and bf16, sampling patterns that stay inside the
windows (encoder-like), leave them for the overflow rows (large offsets) or ignore them altogether (uniform locations:
every point takes the global-memory path).  Tolerance: 2e-2 of the reference's largest magnitude, per tensor
(max-abs-error / max-abs-reference, see tests/helpers.py), the bf16 gate of BASELINE.json.
"""
import pytest
import torch

from oracle import ms_deform_attn_oracle_grads
from tests.helpers import rel_to_max

pytestmark = pytest.mark.gpu

PYR = [(32, 32), (16, 16), (8, 8), (4, 4)]
CASES = [
    ("pyramid_bf16", PYR, 2, torch.bfloat16, "encoder", dict()),
    ("pyramid_f16", PYR, 2, torch.float16, "encoder", dict()),
    ("pyramid_uniform", PYR, 2, torch.bfloat16, "uniform", dict()),
    ("pyramid_big_offsets", PYR, 1, torch.bfloat16, "encoder", dict(offset_sigma_px=5.0)),
    ("odd_shapes", [(25, 38), (13, 19), (7, 10)], 2, torch.bfloat16, "encoder", dict()),
    ("odd_shapes_uniform", [(25, 38), (13, 19), (7, 10)], 1, torch.bfloat16, "uniform", dict()),
    ("single_level", [(20, 24)], 2, torch.bfloat16, "encoder", dict()),
    ("non_nested", [(8, 8), (24, 40), (5, 3)], 1, torch.bfloat16, "encoder", dict()),
    ("heads4_points2", PYR, 2, torch.bfloat16, "encoder", dict(n_heads=4, n_points=2)),
    ("points3", PYR, 1, torch.bfloat16, "encoder", dict(n_points=3)),
    ("levels5_points8", PYR + [(2, 2)], 1, torch.bfloat16, "encoder", dict(n_points=8)),
    ("wide", [(6, 200), (3, 100)], 1, torch.bfloat16, "encoder", dict()),
    ("tall", [(300, 5), (150, 3)], 1, torch.bfloat16, "encoder", dict()),
    ("cfg3_one_image", [(128, 128), (64, 64), (32, 32), (16, 16)], 1, torch.bfloat16, "encoder", dict()),
]


@pytest.fixture
def tiled_mode():
    import vision_instance_seg_b200 as b200
    prev = b200.set_tiled_mode(True)
    yield
    b200.set_tiled_mode(prev)


@pytest.mark.parametrize("name,shapes,batch,dtype,kind,kw", CASES, ids=[c[0] for c in CASES])
def test_tiled_kernels_match_oracle(tiled_mode, name, shapes, batch, dtype, kind, kw):
    from vision_instance_seg_b200 import MSDeformAttnFunction, workloads
    dev = "cuda"
    make = workloads.make_encoder_inputs if kind == "encoder" else workloads.make_uniform_inputs
    if kind != "encoder":
        kw = {k: v for k, v in kw.items() if k != "offset_sigma_px"}
    value, ss, lsi, loc, attn = make(shapes, batch, dtype, seed=11, device=dev, **kw)
    heads = value.shape[2]
    g = torch.Generator(device=dev).manual_seed(5)
    go = torch.randn(batch, loc.shape[1], heads * 32, generator=g, device=dev).to(dtype)
    v = value.clone().requires_grad_(True)
    lo = loc.clone().requires_grad_(True)
    at = attn.clone().requires_grad_(True)
    out = MSDeformAttnFunction.apply(v, ss, lsi, lo, at, 128)
    out.backward(go)
    torch.cuda.synchronize()
    ref = ms_deform_attn_oracle_grads(value.float().cpu(), ss.cpu(), loc.cpu(), attn.cpu(), go.float().cpu())
    # grad_sampling_loc is compared away from integer pixel lines (incl. the gate at -1 and H / W), where the bilinear
    # derivative jumps and two correct implementations may pick different cells (same mask as tests/test_gpu_parity.py)
    wh = torch.stack([ss[:, 1], ss[:, 0]], -1).to(torch.float32)[None, None, None, :, None, :]
    px = loc * wh - 0.5
    smooth = ((px - px.round()).abs() > 1e-3).all(-1, keepdim=True).expand_as(loc).cpu()
    for what, got, want in (("out", out, ref[0]), ("grad_value", v.grad, ref[1]), ("grad_loc", lo.grad, ref[2]),
                            ("grad_attn", at.grad, ref[3])):
        got, want = got.detach().float().cpu(), want.float()
        if what == "grad_loc":
            got, want = got * smooth, want * smooth
        err = rel_to_max(got, want)
        assert err < 2e-2, f"{name} {what}: {err:.3e}"
    # the direct kernels on the same inputs: same arithmetic up to summation order, so no masking is needed
    import vision_instance_seg_b200 as b200
    b200.set_tiled_mode(False)
    v2 = value.clone().requires_grad_(True)
    lo2 = loc.clone().requires_grad_(True)
    at2 = attn.clone().requires_grad_(True)
    out2 = MSDeformAttnFunction.apply(v2, ss, lsi, lo2, at2, 128)
    out2.backward(go)
    torch.cuda.synchronize()
    b200.set_tiled_mode(True)
    assert rel_to_max(out, out2) < 1e-2 and rel_to_max(v.grad, v2.grad) < 2e-2
    assert rel_to_max(lo.grad, lo2.grad) < 1e-4, f"{name}: tiled and direct grad_loc differ"
    assert rel_to_max(at.grad, at2.grad) < 1e-4, f"{name}: tiled and direct grad_attn differ"


def test_tiled_and_direct_kernels_agree_on_the_aux_gradients(tiled_mode):
    """grad_sampling_loc / grad_attn_weight are computed in fp32 from exact 16-bit products in both paths: they agree to
    fp32 rounding, far inside the bf16 gate."""
    import vision_instance_seg_b200 as b200
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    from vision_instance_seg_b200 import workloads
    value, ss, lsi, loc, attn = workloads.make_encoder_inputs(PYR, 2, torch.bfloat16, seed=3, device="cuda")
    go = torch.randn(2, loc.shape[1], 256, device="cuda").to(torch.bfloat16)
    tiled = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    b200.set_tiled_mode(False)
    direct = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
    b200.set_tiled_mode(True)
    torch.cuda.synchronize()
    assert rel_to_max(tiled[1], direct[1]) < 1e-5
    assert rel_to_max(tiled[2], direct[2]) < 1e-5


def test_tiled_mode_leaves_other_call_sites_alone(tiled_mode):
    """Lq != S (decoder) and fp32 values keep using the direct kernels: same launch counts as with the mode off."""
    import vision_instance_seg_b200 as b200
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    from vision_instance_seg_b200 import workloads
    lib = b200.load_library()
    value, ss, lsi, loc, attn = workloads.make_decoder_inputs(PYR, 2, torch.bfloat16, queries=20, device="cuda")
    MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
    n_on = lib.msda_last_launch_count()
    b200.set_tiled_mode(False)
    MSDA.ms_deform_attn_forward(value, ss, lsi, loc, attn, 128)
    n_off = lib.msda_last_launch_count()
    b200.set_tiled_mode(True)
    assert n_on == n_off == 1


HYBRID_CASES = [c for c in CASES if c[0] in ("pyramid_bf16", "pyramid_f16", "pyramid_uniform", "pyramid_big_offsets", "odd_shapes",
                                             "non_nested", "points3", "cfg3_one_image")]


@pytest.mark.parametrize("split", [1, 8, 30])
@pytest.mark.parametrize("name,shapes,batch,dtype,kind,kw", HYBRID_CASES, ids=[c[0] for c in HYBRID_CASES])
def test_hybrid_backward_matches_oracle(name, shapes, batch, dtype, kind, kw, split):
    """Mode 2 (msda_set_tiled_mode(2)): the direct backward kernel keeps grad_sampling_loc / grad_attn_weight and the
    grad_value reductions of the levels that expect at most `split` corner rows per pixel row; the sorting kernel of the
    tiled backward produces grad_value of the coarser levels.  split = 1: every level is sorted; 8: the finest level of
    a /8../64 pyramid stays direct; 30: the two finest.  Same oracle, same 2e-2 gate (max-abs-error / max-abs-reference);
    the aux gradients come from the direct kernel unchanged, so they must be bit-identical to mode 0."""
    import vision_instance_seg_b200 as b200
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    from vision_instance_seg_b200 import workloads
    dev = "cuda"
    make = workloads.make_encoder_inputs if kind == "encoder" else workloads.make_uniform_inputs
    if kind != "encoder":
        kw = {k: v for k, v in kw.items() if k != "offset_sigma_px"}
    value, ss, lsi, loc, attn = make(shapes, batch, dtype, seed=13, device=dev, **kw)
    g = torch.Generator(device=dev).manual_seed(6)
    go = torch.randn(batch, loc.shape[1], value.shape[2] * 32, generator=g, device=dev).to(dtype)
    prev_mode, prev_split = b200.set_tiled_mode(2), b200.set_hybrid_split(split)
    try:
        lib = b200.load_library()
        hyb = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
        launches = lib.msda_last_launch_count()
        torch.cuda.synchronize()
        b200.set_tiled_mode(0)
        direct = MSDA.ms_deform_attn_backward(value, ss, lsi, loc, attn, go, 128)
        torch.cuda.synchronize()
    finally:
        b200.set_tiled_mode(prev_mode)
        b200.set_hybrid_split(prev_split)
    assert launches == 4, "zero + max|grad_out|, direct kernel, sorting kernel, rounding pass"
    ref = ms_deform_attn_oracle_grads(value.float().cpu(), ss.cpu(), loc.cpu(), attn.cpu(), go.float().cpu())
    assert rel_to_max(hyb[0].float().cpu(), ref[1].float()) < 2e-2, f"{name} split {split}: grad_value"
    assert torch.equal(hyb[1], direct[1]) and torch.equal(hyb[2], direct[2])
    assert rel_to_max(hyb[0], direct[0]) < 2e-2
