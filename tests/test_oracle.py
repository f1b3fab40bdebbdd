"""CPU: pin the oracle (three restatements) against the committed golden vectors and against each other."""
import numpy as np
import pytest
import torch

from oracle import ms_deform_attn_core_pytorch, ms_deform_attn_oracle_grads, ms_deform_attn_scalar_numpy
from oracle import c_oracle
from tests.helpers import golden_cases, load_golden, module_golden_cases, random_problem, rel_to_max


@pytest.mark.parametrize("name", golden_cases())
def test_torch_oracle_matches_golden(name):
    g = load_golden(name)
    out, gv, gl, ga = ms_deform_attn_oracle_grads(torch.from_numpy(g["value"]), torch.from_numpy(g["shapes"]),
                                                  torch.from_numpy(g["loc"]), torch.from_numpy(g["attn"]),
                                                  torch.from_numpy(g["grad_out"]))
    assert rel_to_max(out, g["out"]) < 1e-12
    assert rel_to_max(gv, g["grad_value"]) < 1e-12
    assert rel_to_max(gl, g["grad_loc"]) < 1e-12
    assert rel_to_max(ga, g["grad_attn"]) < 1e-12


@pytest.mark.parametrize("name", golden_cases())
def test_c_oracle_matches_golden(name):
    g = load_golden(name)
    out = c_oracle.forward(g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"])
    gv, gl, ga = c_oracle.backward(g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"], g["grad_out"])
    assert rel_to_max(out, g["out"]) < 1e-12
    assert rel_to_max(gv, g["grad_value"]) < 1e-12
    assert rel_to_max(gl, g["grad_loc"]) < 1e-11
    assert rel_to_max(ga, g["grad_attn"]) < 1e-12


@pytest.mark.parametrize("name", ["upstream_test_tiny", "edge_locations"])
def test_scalar_numpy_matches_golden(name):
    g = load_golden(name)
    out, gv, gl, ga = ms_deform_attn_scalar_numpy(g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"],
                                                  g["grad_out"])
    assert rel_to_max(out, g["out"]) < 1e-12
    assert rel_to_max(gv, g["grad_value"]) < 1e-12
    assert rel_to_max(gl, g["grad_loc"]) < 1e-11
    assert rel_to_max(ga, g["grad_attn"]) < 1e-12


def test_upstream_forward_check_float_and_double():
    """Restated upstream ops/test.py check_forward_equal_with_pytorch_{double,float}: seed 3, tiny shapes."""
    g = load_golden("upstream_test_tiny")
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 1e-6)):
        out = ms_deform_attn_core_pytorch(torch.from_numpy(g["value"]).to(dtype), torch.from_numpy(g["shapes"]),
                                          torch.from_numpy(g["loc"]).to(dtype), torch.from_numpy(g["attn"]).to(dtype))
        assert out.dtype == dtype and tuple(out.shape) == (1, 2, 4)
        assert rel_to_max(out, g["out"]) < tol


def test_oracle_gradcheck_fp64():
    value, ss, _, loc, attn, _ = random_problem(1, 2, 4, 3, [(5, 4), (3, 2)], 2, seed=4, loc_range=(0.1, 0.9))
    # keep sampling points away from integer pixel lines (bilinear is only piecewise smooth)
    for l, (H, W) in enumerate(ss.tolist()):
        px = loc[:, :, :, l] * torch.tensor([W, H], dtype=torch.float64) - 0.5
        frac = px - px.floor()
        px = px.floor() + frac.clamp(0.1, 0.9)
        loc[:, :, :, l] = (px + 0.5) / torch.tensor([W, H], dtype=torch.float64)
    value.requires_grad_(True), loc.requires_grad_(True), attn.requires_grad_(True)
    assert torch.autograd.gradcheck(lambda v, lo, a: ms_deform_attn_core_pytorch(v, ss, lo, a), (value, loc, attn),
                                    eps=1e-6, atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("seed,shapes,D,P", [(0, [(6, 4), (3, 2)], 2, 2), (1, [(7, 5)], 30, 3), (2, [(4, 4), (2, 2), (1, 1)], 32, 4)])
def test_three_restatements_agree(seed, shapes, D, P):
    value, ss, lsi, loc, attn, go = random_problem(2, 3, D, 5, shapes, P, seed=seed, loc_range=(-0.3, 1.3))
    t = ms_deform_attn_oracle_grads(value, ss, loc, attn, go)
    c_out = c_oracle.forward(value.numpy(), ss.numpy(), lsi.numpy(), loc.numpy(), attn.numpy())
    c_g = c_oracle.backward(value.numpy(), ss.numpy(), lsi.numpy(), loc.numpy(), attn.numpy(), go.numpy())
    s = ms_deform_attn_scalar_numpy(value.numpy(), ss, lsi, loc.numpy(), attn.numpy(), go.numpy())
    for a, b, c in zip(t, (c_out,) + c_g, s):
        assert rel_to_max(b, a) < 1e-11
        assert rel_to_max(c, a) < 1e-11


def test_empty_attention_and_out_of_range_points_give_zero():
    value, ss, lsi, loc, attn, go = random_problem(1, 2, 8, 4, [(4, 4)], 2, seed=9)
    loc[:] = 2.5                                        # everything outside (-1, H) x (-1, W)
    out = c_oracle.forward(value.numpy(), ss.numpy(), lsi.numpy(), loc.numpy(), attn.numpy())
    assert np.all(out == 0)
    assert float(ms_deform_attn_core_pytorch(value, ss, loc, attn).abs().max()) == 0.0
    gv, gl, ga = c_oracle.backward(value.numpy(), ss.numpy(), lsi.numpy(), loc.numpy(), attn.numpy(), go.numpy())
    assert np.all(gv == 0) and np.all(gl == 0) and np.all(ga == 0)


@pytest.mark.parametrize("name", module_golden_cases())
def test_preop_oracle_composition_matches_module_golden(name):
    """Pins oracle.msdeformattn_preop_pytorch (softmax + sampling-location arithmetic, 2-d and 4-d references) together
    with the core against whole-module vectors produced by transformers' independent MSDeformAttn module."""
    import torch.nn.functional as F
    from oracle import msdeformattn_preop_pytorch
    g = load_golden(name)
    t = {k: torch.from_numpy(v) if isinstance(v, np.ndarray) and v.ndim else v for k, v in g.items()}
    M, P = int(g["heads"]), int(g["points"])
    L = t["shapes"].shape[0]
    prm = {k[len("param."):]: t[k].clone().requires_grad_(True) for k in t if k.startswith("param.")}
    query = t["query"].clone().requires_grad_(True)
    src = t["src"].clone().requires_grad_(True)
    N, Lq, C = query.shape
    S = src.shape[1]
    value = F.linear(src, prm["value_proj.weight"], prm["value_proj.bias"])
    if bool(g["masked"]):
        value = value.masked_fill(t["padding"][..., None], 0.0)
    off = F.linear(query, prm["sampling_offsets.weight"], prm["sampling_offsets.bias"]).view(N, Lq, M, L, P, 2)
    logits = F.linear(query, prm["attention_weights.weight"], prm["attention_weights.bias"]).view(N, Lq, M, L * P)
    loc, aw = msdeformattn_preop_pytorch(t["ref"], off, logits, t["shapes"])
    core = ms_deform_attn_core_pytorch(value.view(N, S, M, C // M), t["shapes"], loc, aw)
    out = F.linear(core, prm["output_proj.weight"], prm["output_proj.bias"])
    out.backward(t["grad_out"])
    assert rel_to_max(out, g["out"]) < 1e-12
    assert rel_to_max(query.grad, g["grad_query"]) < 1e-11 and rel_to_max(src.grad, g["grad_src"]) < 1e-11
    for k, p in prm.items():
        assert rel_to_max(p.grad, g["grad." + k]) < 1e-11, k
