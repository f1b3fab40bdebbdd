"""CPU tests of the host logic behind ``share_value_proj``: the stacked projection, the split
into per-layer views and the shared gradient buffer protocol.  The sampling kernels are not involved (no GPU here): a
stand-in per-layer function with the same buffer protocol as ``MSDeformAttnStackedFunction`` takes their place."""
import torch
from torch.autograd import Function

from vision_instance_seg_b200.modules import MSDeformAttn, StackedValueProj, share_value_proj, unshare_value_proj
from vision_instance_seg_b200.modules.stacked_value_proj import SharedGradBuffer, _SplitStackedValue


def _modules(K, seed=0):
    torch.manual_seed(seed)
    return [MSDeformAttn(64, 2, 4, 2).double() for _ in range(K)]


def test_stacked_projection_equals_per_module_value_proj_and_gradients_flow_through_the_cat():
    K, N, S = 3, 2, 10
    mods = _modules(K)
    proj = StackedValueProj(mods)
    src = torch.randn(N, S, 64, dtype=torch.double, requires_grad=True)
    mask = torch.rand(N, S) < 0.3
    views, value_all, _ = proj.project(src, mask)
    assert value_all.shape == (N, S, K, 4, 16) and len(views) == K
    loss = 0
    for i, (m, v) in enumerate(zip(mods, views)):
        want = m.value_proj(src).masked_fill(mask[..., None], 0.0).view(N, S, 4, 16)
        assert torch.allclose(v, want, atol=1e-12)
        assert v.data_ptr() == value_all.data_ptr() + i * 64 * 8 and v.stride() == (S * K * 64, K * 64, 16, 1)
        loss = loss + (v * (i + 1)).sin().sum()
    loss.backward()                         # foreign consumers of the views: the split's copy path
    g_src = src.grad.clone()
    g_w = [m.value_proj.weight.grad.clone() for m in mods]
    src.grad = None
    for m in mods:
        m.zero_grad()
    loss = sum((m.value_proj(src).masked_fill(mask[..., None], 0.0) * (i + 1)).sin().sum() for i, m in enumerate(mods))
    loss.backward()
    assert torch.allclose(src.grad, g_src, atol=1e-10)
    for m, g in zip(mods, g_w):
        assert torch.allclose(m.value_proj.weight.grad, g, atol=1e-10)


class _FakeLayerOp(Function):
    """y = 2 * value_view, with MSDeformAttnStackedFunction's gradient protocol (writes its slice of the shared buffer)."""

    @staticmethod
    def forward(ctx, value_view, value_all, shared, layer):
        ctx.shared, ctx.layer = shared, layer
        ctx.save_for_backward(value_all)
        return value_view * 2

    @staticmethod
    def backward(ctx, g):
        (value_all,) = ctx.saved_tensors
        buf = ctx.shared.slice_for(value_all, ctx.layer)
        if buf is None:
            return g * 2, None, None, None
        buf[:, :, ctx.layer] = g * 2
        return buf[:, :, ctx.layer], None, None, None


def test_shared_gradient_buffer_is_passed_through_without_copies():
    K = 3
    value_all = torch.randn(2, 5, K, 4, 8, dtype=torch.double, requires_grad=True)
    shared = SharedGradBuffer()
    seen = []
    value_all.register_hook(lambda g: seen.append(g))
    views = _SplitStackedValue.apply(value_all * 1.0, shared)
    outs = [_FakeLayerOp.apply(views[i], value_all.detach(), shared, i) for i in (0, 2)]     # layer 1 never used
    bufs = []
    orig = shared.slice_for

    def spy(like, layer):
        b = orig(like, layer)
        bufs.append(b)
        return b
    shared.slice_for = spy
    (outs[0].sum() + 3 * outs[1].sum()).backward()
    g = seen[0]
    assert torch.equal(g[:, :, 0], torch.full_like(g[:, :, 0], 2.0))
    assert torch.equal(g[:, :, 1], torch.zeros_like(g[:, :, 1]))           # never sampled: zero-filled
    assert torch.equal(g[:, :, 2], torch.full_like(g[:, :, 2], 6.0))
    assert bufs[0] is bufs[1] and shared.buf is None and not shared.written   # one buffer per pass, released afterwards


def test_layer_differentiated_twice_gets_a_private_gradient():
    value_all = torch.randn(1, 3, 2, 2, 8, dtype=torch.double, requires_grad=True)
    shared = SharedGradBuffer()
    views = _SplitStackedValue.apply(value_all * 1.0, shared)
    a = _FakeLayerOp.apply(views[0], value_all.detach(), shared, 0)
    b = _FakeLayerOp.apply(views[0], value_all.detach(), shared, 0)
    (a.sum() + b.sum()).backward()
    assert torch.equal(value_all.grad[:, :, 0], torch.full_like(value_all.grad[:, :, 0], 4.0))
    assert torch.equal(value_all.grad[:, :, 1], torch.zeros_like(value_all.grad[:, :, 1]))


def test_value_cache_is_per_forward_pass():
    K = 3
    mods = _modules(K, seed=1)
    proj = share_value_proj(mods)
    assert all(m._stacked_value == (proj, i) for i, m in enumerate(mods))
    src = torch.randn(1, 6, 64, dtype=torch.double)
    calls = []
    orig = proj.project
    proj.project = lambda s, m: (calls.append(1), orig(s, m))[1]
    for i in range(K):
        proj.value_for(i, src, None)
    assert len(calls) == 1 and proj._views is None            # one GEMM per pass; cache dropped after K layers
    proj.value_for(0, src, None)
    src2 = src.clone()
    proj.value_for(1, src2, None)                              # a different memory tensor: new projection
    assert len(calls) == 3
    src2.add_(1.0)                                             # in-place change of the memory: new projection
    proj.value_for(2, src2, None)
    assert len(calls) == 4
    # the decoder layers each present a fresh `memory.transpose(0, 1)` view of the same memory: one projection
    memory = torch.randn(6, 1, 64, dtype=torch.double, requires_grad=True)
    n = len(calls)
    for i in range(K):
        proj.value_for(i, memory.transpose(0, 1), None)
    assert len(calls) == n + 1
    unshare_value_proj(mods)
    assert all(m._stacked_value is None for m in mods)


def test_cpu_inputs_take_the_plain_path_and_still_raise_without_a_gpu():
    """share_value_proj must not open a CPU path: on CPU tensors the module behaves as upstream does (and the operator
    itself raises, see test_host_cpu.py)."""
    import pytest
    mods = _modules(2, seed=2)
    share_value_proj(mods)
    q = torch.randn(1, 3, 64, dtype=torch.double)
    src = torch.randn(1, 5, 64, dtype=torch.double)
    ss = torch.tensor([(2, 2), (1, 1)])
    with pytest.raises(RuntimeError):
        mods[0](q, torch.rand(1, 3, 2, 2, dtype=torch.double), src, ss, torch.tensor([0, 4]))
