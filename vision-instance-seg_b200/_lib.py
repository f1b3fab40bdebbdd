"""ctypes binding of the C ABI declared in include/msda_b200.h (libmsda_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Loading is lazy so that
``import vision_instance_seg_b200`` works on a machine without the built library; any *use* of the
operator without it raises ``RuntimeError`` — there is deliberately no fallback implementation.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libmsda_b200.so"
_lock = threading.RLock()
_lib = None

MSDA_F32, MSDA_F64, MSDA_BF16, MSDA_F16 = 0, 1, 2, 3
MSDA_BWD_DEFAULT = 0
MSDA_BWD_GRAD_VALUE_FP32_ACCUM = 2
MSDA_BWD_NO_SPARSE_DIRECT = 4
MSDA_BWD_NO_CLUSTER_GUARD = 8


def accum_depth_flag(depth: int) -> int:
    """flags bits that override the fp16 accumulation depth (see include/msda_b200.h)."""
    return (int(depth) & 0xFFFF) << 8

#: every symbol include/msda_b200.h declares (tests check that the built library exports all of them)
EXPORTED_SYMBOLS = (
    "msda_abi_version",
    "msda_error_string",
    "msda_forward",
    "msda_backward_scratch_bytes",
    "msda_backward",
    "msda_forward_strided",
    "msda_backward_strided",
    "msda_set_tiled_mode",
    "msda_set_hybrid_split",
    "msda_last_launch_count",
    "msda_total_launch_count",
    "msda_profile_enable",
    "msda_profile_collect",
    "msda_fused_supported",
    "msda_fused_forward",
    "msda_fused_backward",
    # include/msda_encoder_b200.h
    "msda_enc_add_cast",
    "msda_enc_add_layernorm_forward",
    "msda_enc_add_layernorm_backward_scratch_bytes",
    "msda_enc_add_layernorm_backward",
    "msda_enc_colsum_scratch_bytes",
    "msda_enc_colsum",
    "msda_enc_relu_bwd_colsum",
)


def library_path() -> str:
    return os.environ.get("MSDA_B200_LIBRARY", os.path.join(_HERE, _LIB_NAME))


def _declare(lib):
    vp, i64p, i, sz = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.msda_abi_version.restype = i
    lib.msda_abi_version.argtypes = []
    lib.msda_error_string.restype = ctypes.c_char_p
    lib.msda_error_string.argtypes = [i]
    lib.msda_set_tiled_mode.restype = i
    lib.msda_set_tiled_mode.argtypes = [i]
    lib.msda_set_hybrid_split.restype = i
    lib.msda_set_hybrid_split.argtypes = [i]
    lib.msda_last_launch_count.restype = i
    lib.msda_last_launch_count.argtypes = []
    lib.msda_forward.restype = i
    lib.msda_forward.argtypes = [vp, i64p, i64p, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]
    lib.msda_backward_scratch_bytes.restype = sz
    lib.msda_backward_scratch_bytes.argtypes = [i, i, i, i, i, i, i, i, i]
    lib.msda_backward.restype = i
    lib.msda_backward.argtypes = [vp, i64p, i64p, vp, vp, vp, vp, vp, vp, vp, sz,
                                  i, i, i, i, i, i, i, i, i, i, vp]
    lib.msda_forward_strided.restype = i
    lib.msda_forward_strided.argtypes = [vp, ctypes.c_longlong, i64p, i64p, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]
    lib.msda_backward_strided.restype = i
    lib.msda_backward_strided.argtypes = [vp, ctypes.c_longlong, i64p, i64p, vp, vp, vp, vp, ctypes.c_longlong, vp, vp, vp, sz,
                                          i, i, i, i, i, i, i, i, i, i, vp]
    lib.msda_fused_supported.restype = i
    lib.msda_fused_supported.argtypes = [i, i]
    lib.msda_fused_forward.restype = i
    lib.msda_fused_forward.argtypes = [vp, i64p, i64p, vp, i, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp]
    lib.msda_fused_backward.restype = i
    lib.msda_fused_backward.argtypes = [vp, i64p, i64p, vp, i, vp, vp, vp, vp, vp, vp, vp, sz,
                                        i, i, i, i, i, i, i, i, i, i, i, vp]
    ll, fl = ctypes.c_longlong, ctypes.c_float
    lib.msda_enc_add_cast.restype = i
    lib.msda_enc_add_cast.argtypes = [vp, vp, vp, sz, vp]
    lib.msda_enc_add_layernorm_forward.restype = i
    lib.msda_enc_add_layernorm_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ll, i, fl, vp]
    lib.msda_enc_add_layernorm_backward_scratch_bytes.restype = sz
    lib.msda_enc_add_layernorm_backward_scratch_bytes.argtypes = [i]
    lib.msda_enc_add_layernorm_backward.restype = i
    lib.msda_enc_add_layernorm_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, ll, i, vp]
    lib.msda_enc_colsum_scratch_bytes.restype = sz
    lib.msda_enc_colsum_scratch_bytes.argtypes = [i]
    lib.msda_enc_colsum.restype = i
    lib.msda_enc_colsum.argtypes = [vp, vp, vp, sz, ll, ll, ll, ll, i, vp]
    lib.msda_enc_relu_bwd_colsum.restype = i
    lib.msda_enc_relu_bwd_colsum.argtypes = [vp, vp, vp, vp, sz, ll, i, vp]
    lib.msda_total_launch_count.restype = ctypes.c_longlong
    lib.msda_total_launch_count.argtypes = []
    lib.msda_profile_enable.restype = i
    lib.msda_profile_enable.argtypes = [i]
    lib.msda_profile_collect.restype = i
    lib.msda_profile_collect.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), i]
    return lib


def profile_collect(max_records: int = 65536):
    """Return [(ms, kind), ...] for every kernel timed since msda_profile_enable(1) (see include/msda_b200.h)."""
    lib = load_library()
    ms = (ctypes.c_float * max_records)()
    kinds = (ctypes.c_int * max_records)()
    n = lib.msda_profile_collect(ms, kinds, max_records)
    return [(float(ms[k]), int(kinds[k])) for k in range(n)]


def load_library():
    """Load (once) and return the ctypes handle of libmsda_b200.so; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = library_path()
            if not os.path.exists(path):
                raise RuntimeError(
                    f"{_LIB_NAME} not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). This operator has no CPU or PyTorch fallback.")
            _lib = _declare(ctypes.CDLL(path))
    return _lib


_torch_ext = None
_torch_ext_tried = False


def torch_extension():
    """The compiled torch extension over the same C ABI (csrc/torch_binding.cpp -> _msda_torch*.so next to the library;
    built by ``__graft_entry__.build()``), or None: not built, switched off with MSDA_B200_NO_TORCH_EXT=1, or the library
    path is overridden (the extension links the in-tree library).  Callers fall back to the ctypes route, which reaches
    the very same entry points -- there is no other implementation behind either."""
    global _torch_ext, _torch_ext_tried
    if _torch_ext_tried:
        return _torch_ext
    with _lock:
        if not _torch_ext_tried:
            ext = None
            if os.environ.get("MSDA_B200_NO_TORCH_EXT") != "1" and "MSDA_B200_LIBRARY" not in os.environ \
                    and os.path.exists(library_path()):
                try:
                    import importlib
                    import torch  # noqa: F401  (its shared libraries must be loaded first)
                    want = load_library().msda_abi_version()
                    ext = importlib.import_module(__package__ + "._msda_torch")
                    if ext.abi_version() != want:
                        ext = None
                except ImportError:
                    ext = None
            _torch_ext = ext
            _torch_ext_tried = True
    return _torch_ext


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().msda_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")
