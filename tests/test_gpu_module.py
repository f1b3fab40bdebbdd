"""GPU: the MSDeformAttn nn.Module (encoder and decoder call patterns) against the same module arithmetic on
the CPU with the oracle core in place of the CUDA function."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ms_deform_attn_core_pytorch
from tests.helpers import lsi_of, rel_to_max

pytestmark = pytest.mark.gpu


def reference_module_forward(m, query, reference_points, input_flatten, ss, padding_mask=None):
    """Upstream MSDeformAttn.forward arithmetic, restated with the oracle core (CPU, module's dtype)."""
    N, Lq, _ = query.shape
    _, S, _ = input_flatten.shape
    value = F.linear(input_flatten, m.value_proj.weight, m.value_proj.bias)
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = value.view(N, S, m.n_heads, m.d_model // m.n_heads)
    off = F.linear(query, m.sampling_offsets.weight, m.sampling_offsets.bias).view(N, Lq, m.n_heads, m.n_levels, m.n_points, 2)
    aw = F.linear(query, m.attention_weights.weight, m.attention_weights.bias).view(N, Lq, m.n_heads, m.n_levels * m.n_points)
    aw = aw.softmax(-1).view(N, Lq, m.n_heads, m.n_levels, m.n_points)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([ss[..., 1], ss[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / m.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
    out = ms_deform_attn_core_pytorch(value, ss, loc, aw)
    return F.linear(out, m.output_proj.weight, m.output_proj.bias)


@pytest.fixture(scope="module")
def ops(built_library):
    assert torch.cuda.is_available()
    import vision_instance_seg_b200 as pkg
    return pkg


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("with_mask", [False, True])
def test_module_forward_backward_matches_cpu_restatement(ops, ref_dim, with_mask):
    from vision_instance_seg_b200 import workloads as W
    torch.manual_seed(ref_dim * 10 + with_mask)
    shapes = [(12, 10), (6, 5), (3, 3)]
    ss = W.make_spatial_shapes(shapes)
    lsi = W.make_level_start_index(ss)
    S = int(ss.prod(1).sum())
    N = 2
    m_cpu = ops.MSDeformAttn(d_model=64, n_levels=3, n_heads=4, n_points=4).double()
    with torch.no_grad():   # make the zero-initialised projections non-trivial
        m_cpu.sampling_offsets.weight.normal_(0, 0.05)
        m_cpu.attention_weights.weight.normal_(0, 0.5)
        m_cpu.attention_weights.bias.normal_(0, 0.5)
    m_gpu = ops.MSDeformAttn(d_model=64, n_levels=3, n_heads=4, n_points=4).double()
    m_gpu.load_state_dict(m_cpu.state_dict())
    m_gpu = m_gpu.cuda()
    src = torch.randn(N, S, 64, dtype=torch.float64)
    if ref_dim == 2:
        Lq = S
        query = src + 0.1 * torch.randn(N, S, 64, dtype=torch.float64)
        ref = W.get_reference_points(ss, torch.ones(N, 3, 2)).double()
    else:
        Lq = 13
        query = torch.randn(N, Lq, 64, dtype=torch.float64)
        ref = torch.cat([torch.rand(N, Lq, 1, 2, dtype=torch.float64).expand(-1, -1, 3, -1),
                         torch.rand(N, Lq, 1, 2, dtype=torch.float64).expand(-1, -1, 3, -1) * 0.4 + 0.05], -1)
    mask = None
    if with_mask:
        mask = torch.zeros(N, S, dtype=torch.bool)
        mask[0, ::7] = True
        mask[1, -20:] = True
    q_c = query.clone().requires_grad_(True)
    s_c = src.clone().requires_grad_(True)
    out_c = reference_module_forward(m_cpu, q_c, ref, s_c, ss, mask)
    g = torch.randn_like(out_c)
    out_c.backward(g)
    q_g = query.cuda().requires_grad_(True)
    s_g = src.cuda().requires_grad_(True)
    out_g = m_gpu(q_g, ref.cuda(), s_g, ss.cuda(), lsi.cuda(), mask.cuda() if mask is not None else None)
    out_g.backward(g.cuda())
    assert rel_to_max(out_g, out_c) < 1e-10
    assert rel_to_max(q_g.grad, q_c.grad) < 1e-9
    assert rel_to_max(s_g.grad, s_c.grad) < 1e-9
    for (n, p_g), (_, p_c) in zip(m_gpu.named_parameters(), m_cpu.named_parameters()):
        assert rel_to_max(p_g.grad, p_c.grad) < 1e-9, n


def test_module_fp32_default_shape_and_autocast_bf16(ops):
    """d_model 256 / 8 heads / 4 levels / 4 points in fp32, and under torch.autocast(bfloat16) where the value
    projection arrives as bf16 while locations / weights are computed from bf16 Linears."""
    from vision_instance_seg_b200 import workloads as W
    torch.manual_seed(0)
    shapes = [(16, 16), (8, 8), (4, 4), (2, 2)]
    ss = W.make_spatial_shapes(shapes)
    lsi = W.make_level_start_index(ss)
    S = int(ss.prod(1).sum())
    m = ops.MSDeformAttn().cuda()
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.2)
    src = torch.randn(2, S, 256)
    ref = W.get_reference_points(ss, torch.ones(2, 4, 2))
    out = m(src.cuda(), ref.cuda(), src.cuda(), ss.cuda(), lsi.cuda())
    m_cpu = ops.MSDeformAttn()
    m_cpu.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    want = reference_module_forward(m_cpu, src, ref, src, ss)
    assert rel_to_max(out, want) < 1e-4           # includes four fp32 GEMMs (TF32 off by default for matmul)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out_bf = m(src.cuda(), ref.cuda(), src.cuda(), ss.cuda(), lsi.cuda())
    assert out_bf.dtype == torch.bfloat16
    assert rel_to_max(out_bf, want) < 5e-2


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_module_matches_independent_module_golden(ops, fused, dtype):
    """Whole-module drop-in check: state_dict from transformers' independent MSDeformAttn module (same parameter names)
    loads into ours, and forward / every gradient match its float64 results (tests/golden/make_golden_module.py)."""
    import numpy as np
    from tests.helpers import load_golden, module_golden_cases
    tol = 1e-9 if dtype == torch.float64 else 2e-5
    for name in module_golden_cases():
        g = load_golden(name)
        M, P, L = int(g["heads"]), int(g["points"]), g["shapes"].shape[0]
        C = g["query"].shape[-1]
        m = ops.MSDeformAttn(d_model=C, n_levels=L, n_heads=M, n_points=P).double()     # keep the float64 parameters exact
        missing, unexpected = m.load_state_dict({k[len("param."):]: torch.from_numpy(g[k]) for k in g if k.startswith("param.")})
        assert not missing and not unexpected
        m = m.to("cuda", dtype)
        ops.set_fused_preop(m, fused)            # float64 has no fused kernels: the module composes the plain operator
        q = torch.from_numpy(g["query"]).to("cuda", dtype).requires_grad_(True)
        s = torch.from_numpy(g["src"]).to("cuda", dtype).requires_grad_(True)
        mask = torch.from_numpy(g["padding"]).cuda() if bool(g["masked"]) else None
        out = m(q, torch.from_numpy(g["ref"]).to("cuda", dtype), s, torch.from_numpy(g["shapes"]).cuda(),
                torch.from_numpy(g["level_start_index"]).cuda(), mask)
        out.backward(torch.from_numpy(g["grad_out"]).to("cuda", dtype))
        assert rel_to_max(out, g["out"]) < tol, name
        assert rel_to_max(q.grad, g["grad_query"]) < tol and rel_to_max(s.grad, g["grad_src"]) < tol, name
        for k, p in m.named_parameters():
            assert rel_to_max(p.grad, g["grad." + k]) < tol, (name, k)


def test_unmodified_upstream_autograd_function_runs_on_the_installed_extension(ops):
    """The body of upstream's ops/functions/ms_deform_attn_func.py::MSDeformAttnFunction, restated with its own call
    pattern (`import MultiScaleDeformableAttention as MSDA`, positional arguments, three returned gradients), runs
    unchanged once `install_as_upstream_extension()` has been called, and matches the oracle."""
    import sys
    from torch.autograd import Function
    from torch.autograd.function import once_differentiable
    from oracle import ms_deform_attn_oracle_grads
    from tests.helpers import random_problem
    ops.install_as_upstream_extension()
    try:
        import MultiScaleDeformableAttention as MSDA

        class UpstreamStyleFunction(Function):
            @staticmethod
            def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
                ctx.im2col_step = im2col_step
                output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                                     attention_weights, ctx.im2col_step)
                ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
                return output

            @staticmethod
            @once_differentiable
            def backward(ctx, grad_output):
                value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
                grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
                    value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, grad_output,
                    ctx.im2col_step)
                return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None

        value, ss, lsi, loc, attn, go = random_problem(2, 8, 32, 23, [(9, 7), (4, 5)], 4, seed=31, dtype=torch.float32)
        v = value.cuda().requires_grad_(True)
        lo = loc.cuda().requires_grad_(True)
        at = attn.cuda().requires_grad_(True)
        out = UpstreamStyleFunction.apply(v, ss.cuda(), lsi.cuda(), lo, at, 64)
        out.backward(go.cuda())          # upstream passes grad_output as autograd hands it over (contiguous here)
        want = ms_deform_attn_oracle_grads(value.double(), ss, loc.double(), attn.double(), go.double())
        for a, b in zip((out, v.grad, lo.grad, at.grad), want):
            assert rel_to_max(a, b) < 1e-5
    finally:
        sys.modules.pop("MultiScaleDeformableAttention", None)
