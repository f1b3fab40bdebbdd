"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of the scalar C restatement (oracle/msda_oracle.c).
Parity unpinned against reference tests (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "msda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def forward(value, shapes, lsi, loc, attn):
    v = np.ascontiguousarray(value, dtype=np.float64)
    sh = np.ascontiguousarray(shapes, dtype=np.int64)
    ls = np.ascontiguousarray(lsi, dtype=np.int64)
    lo = np.ascontiguousarray(loc, dtype=np.float64)
    at = np.ascontiguousarray(attn, dtype=np.float64)
    N, S, M, D = v.shape
    _, Lq, _, L, P, _ = lo.shape
    out = np.empty((N, Lq, M * D), dtype=np.float64)
    _load().msda_oracle_forward(_p(v), _p(sh), _p(ls), _p(lo), _p(at), _p(out), N, S, M, D, Lq, L, P)
    return out


def backward(value, shapes, lsi, loc, attn, grad_out):
    v = np.ascontiguousarray(value, dtype=np.float64)
    sh = np.ascontiguousarray(shapes, dtype=np.int64)
    ls = np.ascontiguousarray(lsi, dtype=np.int64)
    lo = np.ascontiguousarray(loc, dtype=np.float64)
    at = np.ascontiguousarray(attn, dtype=np.float64)
    go = np.ascontiguousarray(grad_out, dtype=np.float64)
    N, S, M, D = v.shape
    _, Lq, _, L, P, _ = lo.shape
    gv, gl, ga = np.empty_like(v), np.empty_like(lo), np.empty_like(at)
    _load().msda_oracle_backward(_p(v), _p(sh), _p(ls), _p(lo), _p(at), _p(go), _p(gv), _p(gl), _p(ga),
                                 N, S, M, D, Lq, L, P)
    return gv, gl, ga
