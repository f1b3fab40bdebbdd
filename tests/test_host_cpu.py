"""CPU: host-side logic, the C-ABI surface and the fail-loudly contract (no compute calls without a GPU)."""
import ctypes
import math
import os
import re

import pytest
import torch

import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import MSDeformAttn, MSDeformAttnFunction, _lib, workloads as W
from tests.helpers import lsi_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = "\n".join(open(os.path.join(ROOT, "include", h)).read() for h in ("msda_b200.h", "msda_encoder_b200.h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_and_library_exports_every_symbol(built_library):
    declared = _declared_functions()
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), (declared, _lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(built_library)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    assert pkg.load_library().msda_abi_version() == 5


def test_error_strings(built_library):
    lib = pkg.load_library()
    seen = set()
    for code in range(-6, 1):
        msg = lib.msda_error_string(code).decode()
        assert msg and msg != "unknown msda error"
        seen.add(msg)
    assert len(seen) == 7
    assert lib.msda_error_string(-99).decode() == "unknown msda error"


def test_argument_validation_without_gpu(built_library):
    """Validation happens before any CUDA call, so the negative codes are observable on a CPU-only box."""
    lib = pkg.load_library()
    buf = (ctypes.c_char * 256)()
    p = ctypes.addressof(buf)
    p = (p + 15) & ~15
    ok_dims = (1, 4, 1, 4, 1, 1, 1)
    assert lib.msda_forward(None, p, p, p, p, p, *ok_dims, 0, 64, None) == -1
    assert lib.msda_forward(p, p, p, p, p, p, 0, 4, 1, 4, 1, 1, 1, 0, 64, None) == -2
    assert lib.msda_forward(p, p, p, p, p, p, 1, 4, 1, 4, 1, 33, 1, 0, 64, None) == -2
    assert lib.msda_forward(p, p, p, p, p, p, *ok_dims, 9, 64, None) == -3
    assert lib.msda_forward(p + 4, p, p, p, p, p, *ok_dims, 0, 64, None) == -4
    assert lib.msda_forward(p, p, p, p, p, p, 6, 4, 1, 4, 1, 1, 1, 0, 4, None) == -5     # 6 % 4 != 0
    assert lib.msda_backward(p, p, p, p, p, p, p, p, p, None, 0, *ok_dims, 2, 64, 0, None) == -6  # bf16 needs scratch
    # (N, S, M, D, Lq, L, P, dtype, flags): fp32 accumulation = one float per value element ...
    assert lib.msda_backward_scratch_bytes(2, 10, 8, 32, 7, 2, 4, _lib.MSDA_BF16, _lib.MSDA_BWD_GRAD_VALUE_FP32_ACCUM) == 2 * 10 * 8 * 32 * 4
    # ... bucketed fp16 accumulation (default) = 256-byte control block + N * rows_bound * M * D halves,
    # rows_bound = L * (ceil(Lq*P/depth) + 1) + 2*S
    rows = 2 * ((7 * 4 + 31) // 32 + 1) + 2 * 10
    ng = _lib.MSDA_BWD_NO_CLUSTER_GUARD
    assert lib.msda_backward_scratch_bytes(2, 10, 8, 32, 7, 2, 4, _lib.MSDA_BF16, ng) == 256 + 2 * rows * 8 * 32 * 2
    rows4 = 2 * ((7 * 4 + 3) // 4 + 1) + 2 * 10
    assert lib.msda_backward_scratch_bytes(2, 10, 8, 32, 7, 2, 4, _lib.MSDA_F16, _lib.accum_depth_flag(4) | ng) == 256 + 2 * rows4 * 8 * 32 * 2
    # ... with the cluster guard (default for calls of <= 4 Mi sampling points): + per-(4-pixel run, head) counters,
    # rounded up to 256 bytes, and a payload large enough for either accumulation mode
    counters = -(-(2 * ((10 + 3) // 4 + 1) * 8 * 4) // 256) * 256
    assert lib.msda_backward_scratch_bytes(2, 10, 8, 32, 7, 2, 4, _lib.MSDA_BF16, 0) == \
        256 + counters + max(2 * rows * 8 * 32 * 2, 2 * 10 * 8 * 32 * 4)
    # a call too large for the density pass is sized as without the guard
    big = (16, 21760, 8, 32, 21760, 4, 4)
    assert lib.msda_backward_scratch_bytes(*big, _lib.MSDA_BF16, 0) == lib.msda_backward_scratch_bytes(*big, _lib.MSDA_BF16, ng)
    assert lib.msda_backward_scratch_bytes(2, 10, 8, 30, 7, 2, 4, _lib.MSDA_BF16, 0) == 2 * 10 * 8 * 30 * 4   # compat kernels
    assert lib.msda_backward_scratch_bytes(2, 10, 8, 32, 7, 2, 4, _lib.MSDA_F32, 0) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("MSDA_B200_LIBRARY", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load_library()


def test_cpu_tensors_are_rejected_not_silently_computed():
    ss = torch.tensor([(4, 4)], dtype=torch.long)
    value = torch.randn(1, 16, 2, 8)
    loc = torch.rand(1, 3, 2, 1, 2, 2)
    attn = torch.rand(1, 3, 2, 1, 2)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        MSDeformAttnFunction.apply(value, ss, lsi_of(ss), loc, attn, 64)
    m = MSDeformAttn(d_model=16, n_levels=1, n_heads=2, n_points=2)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        m(torch.randn(1, 3, 16), torch.rand(1, 3, 1, 2), torch.randn(1, 16, 16), ss, lsi_of(ss))
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDeformAttnFunction.apply(value.transpose(2, 3).contiguous().transpose(2, 3), ss, lsi_of(ss), loc, attn, 64)


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "vision-instance-seg_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "grid_sample" not in text, f


def test_module_surface_matches_upstream():
    m = MSDeformAttn()
    assert (m.d_model, m.n_levels, m.n_heads, m.n_points, m.im2col_step) == (256, 4, 8, 4, 128)
    keys = set(m.state_dict().keys())
    assert keys == {f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
                    for p in ("weight", "bias")}
    assert m.sampling_offsets.weight.shape == (8 * 4 * 4 * 2, 256)
    assert m.attention_weights.weight.shape == (8 * 4 * 4, 256)
    assert sum(p.numel() for p in m.parameters()) == 230272
    with pytest.raises(ValueError, match="divisible"):
        MSDeformAttn(d_model=100, n_heads=8)
    with pytest.warns(UserWarning, match="power of 2"):
        MSDeformAttn(d_model=240, n_heads=8)


def test_module_initialisation():
    m = MSDeformAttn(d_model=64, n_levels=3, n_heads=4, n_points=2)
    assert float(m.sampling_offsets.weight.abs().max()) == 0.0
    assert float(m.attention_weights.weight.abs().max()) == 0.0 and float(m.attention_weights.bias.abs().max()) == 0.0
    assert float(m.value_proj.bias.abs().max()) == 0.0 and float(m.output_proj.bias.abs().max()) == 0.0
    bias = m.sampling_offsets.bias.detach().view(4, 3, 2, 2)
    for h in range(4):
        th = h * 2 * math.pi / 4
        d = torch.tensor([math.cos(th), math.sin(th)])
        d = d / d.abs().max()
        for p in range(2):
            assert torch.allclose(bias[h, :, p], (d * (p + 1)).expand(3, 2), atol=1e-6)
    assert torch.allclose(bias, W.init_offset_pattern(4, 3, 2), atol=1e-6)
    bound = math.sqrt(6.0 / (64 + 64))
    assert float(m.value_proj.weight.abs().max()) <= bound + 1e-6


def test_shape_builders():
    ss = W.make_spatial_shapes([(64, 64), (32, 32), (16, 16)])
    lsi = W.make_level_start_index(ss)
    assert lsi.tolist() == [0, 4096, 5120] and ss.dtype == torch.long and lsi.dtype == torch.long
    ref = W.get_reference_points(ss, torch.ones(2, 3, 2))
    assert tuple(ref.shape) == (2, 5376, 3, 2)
    assert torch.allclose(ref[0, 0, 0], torch.tensor([0.5 / 64, 0.5 / 64]))
    assert torch.allclose(ref[0, 4096 + 33, 1], torch.tensor([1.5 / 32, 1.5 / 32]))
    vr = torch.full((1, 3, 2), 0.5)
    ref_h = W.get_reference_points(ss, vr)
    assert torch.allclose(ref_h[0, 0, 0], torch.tensor([0.5 / 64, 0.5 / 64]))       # (c / (vr*W)) * vr
    ab = W.algorithmic_bytes(16, 21760, 21760, 8, 32, 4, 4, 2)
    assert ab["points"] == 44564480 and ab["fwd"] == 20 * ab["points"] and ab["bwd"] == 36 * ab["points"]


def test_synthetic_workloads_on_cpu():
    v, ss, lsi, loc, attn = W.make_encoder_inputs([(8, 8), (4, 4)], 2, torch.float32, n_heads=2, head_dim=8, device="cpu")
    assert tuple(v.shape) == (2, 80, 2, 8) and tuple(loc.shape) == (2, 80, 2, 2, 4, 2)
    assert torch.allclose(attn.sum((-1, -2)), torch.ones(2, 80, 2), atol=1e-5)
    v, ss, lsi, loc, attn = W.make_decoder_inputs([(8, 8), (4, 4)], 2, torch.bfloat16, queries=5, n_heads=2, head_dim=8, device="cpu")
    assert v.dtype == torch.bfloat16 and loc.dtype == torch.float32 and tuple(loc.shape) == (2, 5, 2, 2, 4, 2)


def test_install_as_upstream_extension_binds_the_upstream_import_name(monkeypatch):
    """Upstream's ms_deform_attn_func.py does `import MultiScaleDeformableAttention as MSDA`; after the helper that import
    resolves to this package's stand-in (same two functions), so an unmodified checkout needs no extension build."""
    import importlib
    import sys
    monkeypatch.delitem(sys.modules, "MultiScaleDeformableAttention", raising=False)
    with pytest.raises(ImportError):
        importlib.import_module("MultiScaleDeformableAttention")
    mod = pkg.install_as_upstream_extension()
    try:
        import MultiScaleDeformableAttention as MSDA
        assert MSDA is mod is pkg.MultiScaleDeformableAttention
        assert callable(MSDA.ms_deform_attn_forward) and callable(MSDA.ms_deform_attn_backward)
        # the upstream autograd function body, verbatim in spirit: it only needs these two entry points
        with pytest.raises(RuntimeError):          # CPU tensors: same error class as upstream's AT_ERROR
            MSDA.ms_deform_attn_forward(torch.zeros(1, 4, 1, 8), torch.tensor([[2, 2]]), torch.tensor([0]),
                                        torch.zeros(1, 1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1), 64)
    finally:
        sys.modules.pop("MultiScaleDeformableAttention", None)


def test_decoder_reference_points_input():
    g = torch.Generator().manual_seed(0)
    boxes = torch.rand(2, 5, 4, generator=g)
    vr = torch.tensor([[[1.0, 1.0], [0.5, 0.75], [0.25, 1.0]], [[0.8, 0.9], [1.0, 1.0], [0.6, 0.7]]])
    out = W.decoder_reference_points_input(boxes, vr)
    assert out.shape == (2, 5, 3, 4)
    assert torch.allclose(out[1, 3, 2], boxes[1, 3] * torch.tensor([0.6, 0.7, 0.6, 0.7]))
    pts = W.decoder_reference_points_input(boxes[..., :2], vr)
    assert pts.shape == (2, 5, 3, 2) and torch.allclose(pts[0, 0, 1], boxes[0, 0, :2] * torch.tensor([0.5, 0.75]))
    with pytest.raises(ValueError):
        W.decoder_reference_points_input(torch.rand(1, 2, 3), vr[:1])


def test_torch_extension_is_built_and_mirrors_the_ctypes_route(built_library):
    """csrc/torch_binding.cpp: the compiled torch extension over the same C ABI loads next to the library, exposes upstream's
    two functions and refuses CPU / non-contiguous tensors with upstream's messages -- before any CUDA call."""
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    ext = _lib.torch_extension()
    if ext is None:        # optional by design: build() removes an extension that fails its self-test in this environment
        pytest.skip("torch extension not available here; the ctypes route serves the same C ABI")
    assert ext.abi_version() == pkg.load_library().msda_abi_version()
    ext.raise_for_code(0)
    for code in (-1, -5, -6):           # a C-ABI error code becomes the same RuntimeError as on the ctypes route
        with pytest.raises(RuntimeError) as info:
            ext.raise_for_code(code)
        with pytest.raises(RuntimeError) as want:
            _lib.check(code, "raise_for_code")
        assert str(want.value) in str(info.value)
    value = torch.zeros(1, 4, 1, 4)
    shapes = torch.tensor([[2, 2]])
    lsi = torch.zeros(1, dtype=torch.long)
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    attn = torch.zeros(1, 1, 1, 1, 1)
    for route in ("extension", "ctypes"):
        saved = _lib._torch_ext
        if route == "ctypes":
            _lib._torch_ext = None
        try:
            with pytest.raises(RuntimeError, match="value must be a CUDA tensor"):
                MSDA.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64)
            with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):
                MSDA.ms_deform_attn_forward(value.transpose(1, 3), shapes, lsi, loc, attn, 64)
            with pytest.raises(RuntimeError, match="grad_output must be a CUDA tensor|value must be a CUDA tensor"):
                MSDA.ms_deform_attn_backward(value, shapes, lsi, loc, attn, torch.zeros(1, 1, 4), 64)
        finally:
            _lib._torch_ext = saved
