"""The compiled torch extension (csrc/torch_binding.cpp) against the ctypes route: both call the same C-ABI entry points, so
everything that is deterministic must be bit-identical; grad_value (atomics: order-dependent rounding) within the dtype's
gate.  Also the behaviours a binding can get wrong: non-default stream, a device other than the current one is not
available on a 1-GPU box, so the device guard is exercised through `torch.cuda.device`; dtype conversions of the aux
tensors; error codes turned into RuntimeError."""
import pytest
import torch

from tests.helpers import rel_to_max

pytestmark = pytest.mark.gpu


def _both_routes(fn):
    from vision_instance_seg_b200 import _lib
    ext = _lib.torch_extension()
    if ext is None:
        pytest.skip("torch extension not available here; the ctypes route serves the same C ABI")
    a = fn()
    saved = _lib._torch_ext
    _lib._torch_ext = None
    try:
        b = fn()
    finally:
        _lib._torch_ext = saved
    torch.cuda.synchronize()
    return a, b


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, torch.float64])
@pytest.mark.parametrize("kind", ["encoder", "decoder"])
def test_extension_and_ctypes_routes_agree(dtype, kind):
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    from vision_instance_seg_b200 import workloads as W
    shapes = [(12, 16), (6, 8), (3, 4)]
    if kind == "encoder":
        v, ss, lsi, loc, attn = W.make_encoder_inputs(shapes, 2, dtype, seed=5, device="cuda")
    else:
        v, ss, lsi, loc, attn = W.make_decoder_inputs(shapes, 2, dtype, queries=17, seed=5, device="cuda")
    if dtype == torch.float64:
        loc, attn = loc.double(), attn.double()
    go = torch.randn(2, loc.shape[1], v.shape[2] * v.shape[3], device="cuda").to(dtype)
    side = torch.cuda.Stream()

    def run():
        with torch.cuda.stream(side):
            out = MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 128)
            gv, gl, ga = MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128)
        side.synchronize()
        return out, gv, gl, ga

    (o1, gv1, gl1, ga1), (o2, gv2, gl2, ga2) = _both_routes(run)
    assert torch.equal(o1, o2) and torch.equal(gl1, gl2) and torch.equal(ga1, ga2)
    assert o1.dtype == dtype and gv1.dtype == dtype and gl1.dtype == loc.dtype and ga1.dtype == attn.dtype
    assert rel_to_max(gv1, gv2) < (2e-2 if dtype in (torch.bfloat16, torch.float16) else 1e-5)


def test_extension_converts_aux_dtypes_and_reports_errors():
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    from vision_instance_seg_b200 import workloads as W
    v, ss, lsi, loc, attn = W.make_encoder_inputs([(8, 8), (4, 4)], 3, torch.bfloat16, seed=2, device="cuda")
    go = torch.randn(3, loc.shape[1], 256, device="cuda")          # fp32 grad_output for bf16 values: converted
    (a, b) = _both_routes(lambda: MSDA.ms_deform_attn_backward(v, ss, lsi, loc.bfloat16(), attn.bfloat16(), go, 128))
    assert a[1].dtype == torch.bfloat16 and a[2].dtype == torch.bfloat16
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    for route_result in _both_routes(lambda: _raises(lambda: MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 2))):
        assert "im2col_step" in route_result              # N = 3 is not a multiple of min(N, 2)
    int32_shapes = _both_routes(lambda: MSDA.ms_deform_attn_forward(v, ss.int(), lsi.int(), loc, attn, 128))
    assert torch.equal(int32_shapes[0], int32_shapes[1])
    empty = _both_routes(lambda: MSDA.ms_deform_attn_forward(v, ss, lsi, loc[:, :0], attn[:, :0], 128))
    assert empty[0].shape == empty[1].shape == (3, 0, 256)


def _raises(fn):
    try:
        fn()
    except RuntimeError as e:
        return str(e)
    return ""
