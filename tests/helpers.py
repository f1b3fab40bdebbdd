"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_cases():
    """Operator-level fixtures (tests/golden/make_golden.py)."""
    return [n for n in _all_golden() if not n.startswith("module_")]


def module_golden_cases():
    """Module-level fixtures (tests/golden/make_golden_module.py): whole MSDeformAttn.forward incl. the pre-op."""
    return [n for n in _all_golden() if n.startswith("module_")]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def lsi_of(shapes: torch.Tensor) -> torch.Tensor:
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def rel_to_max(got, want) -> float:
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    denom = max(float(want.abs().max()), 1e-30)
    return float((got - want).abs().max()) / denom


def random_problem(N, M, D, Lq, shapes, P, seed=0, dtype=torch.float64, loc_range=(-0.1, 1.1)):
    g = torch.Generator().manual_seed(seed)
    ss = torch.as_tensor(shapes, dtype=torch.long)
    S = int(ss.prod(1).sum())
    L = len(shapes)
    value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64).to(dtype)
    lo, hi = loc_range
    loc = (torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype)
    attn = torch.rand(N, Lq, M, L, P, generator=g, dtype=torch.float64) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dtype)
    grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64).to(dtype)
    return value, ss, lsi_of(ss), loc, attn, grad_out
