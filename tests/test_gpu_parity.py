"""GPU parity tests proper: the CUDA path (through the public Python surface -> ctypes -> C ABI) against the
CPU oracle on identical inputs.  Tolerances (BASELINE.json north_star): 1e-5 relative in fp32, 2e-2 in bf16;
fp64 (compatibility kernels, used for gradcheck) is held to 1e-10.  Relative = max-abs error / max-abs of
the oracle tensor."""
import pytest
import torch

from oracle import ms_deform_attn_core_pytorch, ms_deform_attn_oracle_grads
from tests.helpers import golden_cases, load_golden, lsi_of, random_problem, rel_to_max

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-10, torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 5e-3}


@pytest.fixture(scope="module")
def ops(built_library):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import vision_instance_seg_b200 as pkg
    pkg.load_library()
    return pkg


def run_cuda(ops, value, ss, lsi, loc, attn, grad_out, dtype, im2col_step=64):
    dev = torch.device("cuda:0")
    aux = torch.float64 if dtype == torch.float64 else torch.float32
    v = value.to(dev, dtype).requires_grad_(True)
    lo = loc.to(dev, aux).requires_grad_(True)
    at = attn.to(dev, aux).requires_grad_(True)
    out = ops.MSDeformAttnFunction.apply(v, ss.to(dev), lsi.to(dev), lo, at, im2col_step)
    out.backward(grad_out.to(dev, dtype))
    torch.cuda.synchronize()
    return out.detach(), v.grad, lo.grad, at.grad, (v, lo, at)


def oracle_on_rounded_inputs(value, ss, loc, attn, grad_out, dtype):
    """The oracle sees exactly the numbers the kernel sees (inputs rounded to the kernel's dtypes), in fp64."""
    aux = torch.float64 if dtype == torch.float64 else torch.float32
    return ms_deform_attn_oracle_grads(value.to(dtype).double(), ss, loc.to(aux).double(), attn.to(aux).double(),
                                       grad_out.to(dtype).double())


def check(ops, problem, dtype, tol=None, names=("out", "grad_value", "grad_loc", "grad_attn")):
    value, ss, lsi, loc, attn, go = problem
    got = run_cuda(ops, value, ss, lsi, loc, attn, go, dtype)
    want = oracle_on_rounded_inputs(value, ss, loc, attn, go, dtype)
    tol = TOL[dtype] if tol is None else tol
    for name, g, w in zip(names, got, want):
        err = rel_to_max(g, w)
        assert err < tol, f"{name} ({dtype}): rel-to-max error {err:.3e} >= {tol}"
    assert got[0].dtype == dtype and got[1].dtype == dtype
    assert got[2].dtype == got[4][1].dtype and got[3].dtype == got[4][2].dtype


# ---------------------------------------------------------------------------------------------------
# golden fixtures (pins produced by transformers' independent implementation, see tests/golden)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.bfloat16])
def test_golden_vectors(ops, name, dtype):
    g = load_golden(name)
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    got = run_cuda(ops, t["value"], t["shapes"], t["level_start_index"], t["loc"], t["attn"], t["grad_out"], dtype)
    if dtype == torch.bfloat16:
        # compare against the oracle fed with the bf16-rounded inputs: the golden outputs belong to fp64 inputs
        want = oracle_on_rounded_inputs(t["value"], t["shapes"], t["loc"], t["attn"], t["grad_out"], dtype)
    else:
        want = (t["out"], t["grad_value"], t["grad_loc"], t["grad_attn"])
    for name_, a, b in zip(("out", "grad_value", "grad_loc", "grad_attn"), got, want):
        assert rel_to_max(a, b) < TOL[dtype], name_


def test_upstream_check_forward_equal_with_pytorch(ops):
    """Restated upstream ops/test.py: N,M,D=1,2,2; Lq,L,P=2,2,2; shapes [(6,4),(3,2)]; seed 3; im2col_step 2;
    the float variant is held to upstream's own allclose(rtol=1e-2, atol=1e-3) and to our 1e-5."""
    g = load_golden("upstream_test_tiny")
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    dev = "cuda:0"
    for dtype in (torch.float64, torch.float32):
        out = ops.MSDeformAttnFunction.apply(t["value"].to(dev, dtype), t["shapes"].to(dev), t["level_start_index"].to(dev),
                                             t["loc"].to(dev, dtype), t["attn"].to(dev, dtype), 2)
        ref = ms_deform_attn_core_pytorch(t["value"].to(dtype), t["shapes"], t["loc"].to(dtype), t["attn"].to(dtype))
        assert torch.allclose(out.cpu(), ref, rtol=1e-2, atol=1e-3)
        assert rel_to_max(out, t["out"]) < TOL[dtype]


@pytest.mark.parametrize("channels", [30, 32, 64, 71, 1025, 2048, 3096])
def test_upstream_check_gradient_numerical(ops, channels):
    """Restated upstream check_gradient_numerical: torch.autograd.gradcheck of the fp64 CUDA path."""
    value, ss, lsi, loc, attn, _ = random_problem(1, 2, channels, 2, [(6, 4), (3, 2)], 2, seed=3, loc_range=(0.05, 0.95))
    for l, (H, W) in enumerate(ss.tolist()):     # stay off the integer pixel lines where bilinear has a kink
        scale = torch.tensor([W, H], dtype=torch.float64)
        px = loc[:, :, :, l] * scale - 0.5
        px = px.floor() + (px - px.floor()).clamp(0.15, 0.85)
        loc[:, :, :, l] = (px + 0.5) / scale
    dev = "cuda:0"
    v = (value * 0.01).to(dev).requires_grad_(True)
    lo = loc.to(dev).requires_grad_(True)
    at = attn.to(dev).requires_grad_(True)
    func = ops.MSDeformAttnFunction.apply
    assert torch.autograd.gradcheck(func, (v, ss.to(dev), lsi.to(dev), lo, at, 2), eps=1e-6, atol=1e-7, rtol=1e-5,
                                    nondet_tol=1e-12)


# ---------------------------------------------------------------------------------------------------
# kernel-variant matrix on random problems
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("D", [16, 32, 64, 128])
def test_vector_kernels_all_head_dims(ops, dtype, D):
    check(ops, random_problem(2, 3, D, 37, [(9, 7), (5, 4), (2, 3)], 4, seed=D), dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("D", [1, 8, 30, 71, 200])
def test_compat_kernels_any_head_dim(ops, dtype, D):
    check(ops, random_problem(2, 2, D, 11, [(6, 5), (3, 3)], 3, seed=100 + D), dtype)


@pytest.mark.parametrize("L,P", [(1, 1), (3, 3), (2, 5), (5, 8), (4, 4)])
def test_level_point_counts_including_unaligned_rows(ops, L, P):
    shapes = [(10, 12), (7, 5), (4, 4), (3, 2), (1, 1)][:L]
    check(ops, random_problem(2, 4, 32, 19, shapes, P, seed=L * 10 + P), torch.float32)
    check(ops, random_problem(2, 4, 32, 19, shapes, P, seed=L * 10 + P), torch.bfloat16)


@pytest.mark.parametrize("N,Lq,M", [(1, 1, 1), (1, 3, 3), (3, 5, 7), (2, 65, 8)])
def test_ragged_tails(ops, N, Lq, M):
    """pair counts that do not fill the last warp / CTA"""
    check(ops, random_problem(N, M, 32, Lq, [(5, 5), (3, 3)], 4, seed=N + Lq + M), torch.float32)
    check(ops, random_problem(N, M, 32, Lq, [(5, 5), (3, 3)], 4, seed=N + Lq + M), torch.bfloat16)


def test_all_points_outside_give_exact_zeros(ops):
    value, ss, lsi, loc, attn, go = random_problem(2, 8, 32, 9, [(4, 4), (2, 2)], 4, seed=1)
    loc[:] = -3.0
    out, gv, gl, ga, _ = run_cuda(ops, value, ss, lsi, loc, attn, go, torch.float32)
    for t in (out, gv, gl, ga):
        assert float(t.abs().max()) == 0.0


def test_zero_attention_rows(ops):
    value, ss, lsi, loc, attn, go = random_problem(2, 8, 32, 9, [(4, 4), (2, 2)], 4, seed=2)
    attn[:, ::2] = 0
    check(ops, (value, ss, lsi, loc, attn, go), torch.float32)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configurations
# ---------------------------------------------------------------------------------------------------
def test_config1_and_2_fp32_forward_backward(ops):
    """configs[0]/[1]: batch 2, 512x512 -> levels 64/32/16, d_model 256, 8 heads, 4 points, fp32, fwd+bwd."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg1_512_fp32"]
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], cfg["batch"], torch.float32, device="cpu", seed=1234)
    go = torch.randn(v.shape[0], loc.shape[1], 256, generator=torch.Generator().manual_seed(5))
    check(ops, (v, ss, lsi, loc, attn, go), torch.float32)
    v, ss, lsi, loc, attn = W.make_uniform_inputs(cfg["shapes"], cfg["batch"], torch.float32, device="cpu", seed=77)
    check(ops, (v, ss, lsi, loc, attn, go), torch.float32)


def test_config3_encoder_shape_bf16_one_image(ops):
    """configs[2] geometry (1024^2, 4 levels, 21 760 queries) at batch 1 so the oracle finishes in seconds."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg3_swinl_1024_bf16"]
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], 1, torch.bfloat16, device="cpu", seed=1234)
    go = torch.randn(1, loc.shape[1], 256, generator=torch.Generator().manual_seed(6))
    value, ss_, lsi_, loc_, attn_, go_ = v.float(), ss, lsi, loc, attn, go
    got = run_cuda(ops, value, ss_, lsi_, loc_, attn_, go_, torch.bfloat16)
    want = ms_deform_attn_oracle_grads(value.to(torch.bfloat16).float(), ss_, loc_, attn_, go_.to(torch.bfloat16).float(),
                                       dtype=torch.float32)
    # grad_loc is compared away from integer pixel lines (see test_config5_encoder_shape_bf16_one_image)
    wh = torch.stack([ss_[:, 1], ss_[:, 0]], -1).to(torch.float32)[None, None, None, :, None, :]
    px = loc_ * wh - 0.5
    smooth = ((px - px.round()).abs() > 1e-3).all(-1, keepdim=True).expand_as(loc_)
    for name, a, b in zip(("out", "grad_value", "grad_loc", "grad_attn"), got, want):
        a, b = a.detach().float().cpu(), b.float()
        if name == "grad_loc":
            a, b = a * smooth, b * smooth
        assert rel_to_max(a, b) < 2e-2, name


def test_config5_encoder_shape_bf16_one_image(ops):
    """configs[4] geometry (2048^2 -> 256/128/64/32, 87 040 queries) at batch 1 against the oracle (fp32 on the CPU)."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg5_2048_bf16"]
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], 1, torch.bfloat16, device="cpu", seed=77)
    go = torch.randn(1, loc.shape[1], 256, generator=torch.Generator().manual_seed(8))
    got = run_cuda(ops, v.float(), ss, lsi, loc, attn, go, torch.bfloat16)
    want = ms_deform_attn_oracle_grads(v.float(), ss, loc, attn, go.to(torch.bfloat16).float(), dtype=torch.float32)
    # 11 M sampling points: a handful sit within float rounding of an integer pixel line, where the bilinear derivative
    # jumps and the fp32 oracle and the kernel may legitimately pick different cells; grad_loc is compared away from them
    wh = torch.stack([ss[:, 1], ss[:, 0]], -1).to(torch.float32)[None, None, None, :, None, :]
    px = loc * wh - 0.5
    smooth = ((px - px.round()).abs() > 1e-3).all(-1, keepdim=True).expand_as(loc)
    for name, a, b in zip(("out", "grad_value", "grad_loc", "grad_attn"), got, want):
        a, b = a.detach().float().cpu(), b.float()
        if name == "grad_loc":
            a, b = a * smooth, b * smooth
        assert rel_to_max(a, b) < 2e-2, name


def test_config4_decoder_shape_bf16(ops):
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg4_decoder_300q_bf16"]
    v, ss, lsi, loc, attn = W.make_decoder_inputs(cfg["shapes"], 2, torch.bfloat16, queries=300, device="cpu", seed=3)
    go = torch.randn(2, 300, 256, generator=torch.Generator().manual_seed(7))
    check(ops, (v.float(), ss, lsi, loc, attn, go), torch.bfloat16)


def test_full_size_properties_config3(ops):
    """At BASELINE.json's full cfg3 size (N=16) the oracle is too slow; check size-independent properties:
    (1) value == 1 and interior sampling -> output == sum(attn) == 1; (2) linearity in value;
    (3) checksum of checksums: sum_s grad_value[b,:,m,d] == sum_q grad_out[b,q,m,d] for interior sampling;
    (4) batch permutation equivariance (bit-exact forward)."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg3_swinl_1024_bf16"]
    dev = "cuda:0"
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], 16, torch.bfloat16, device=dev, seed=99)
    f = ops.MSDeformAttnFunction.apply
    # interior locations: clamp so that every footprint is fully inside every level
    lo_in = loc.clamp(0.04, 0.96).contiguous()
    ones = torch.ones_like(v)
    out = f(ones, ss, lsi, lo_in, attn, 128)
    assert float((out.float() - 1).abs().max()) < 1e-2           # bf16 rounding of ~1.0
    # linearity (fp32 accumulate, one bf16 rounding at the end)
    v2 = torch.randn_like(v)
    a = f(v, ss, lsi, loc, attn, 128).float()
    b = f(v2, ss, lsi, loc, attn, 128).float()
    c = f((v.float() * 0.5 + v2.float() * 0.25).to(torch.bfloat16), ss, lsi, loc, attn, 128).float()
    assert rel_to_max(c, 0.5 * a + 0.25 * b) < 2e-2
    # checksum of checksums through backward
    vv = ones.clone().requires_grad_(True)
    go = torch.randn(16, loc.shape[1], 256, device=dev, dtype=torch.bfloat16)
    f(vv, ss, lsi, lo_in, attn, 128).backward(go)
    lhs = vv.grad.float().sum(1)                                   # (N, M, D)
    rhs = go.float().view(16, -1, 8, 32).sum(1)
    assert rel_to_max(lhs, rhs) < 2e-2
    # batch permutation
    perm = torch.randperm(16, device=dev)
    a_p = f(v[perm].contiguous(), ss, lsi, loc[perm].contiguous(), attn[perm].contiguous(), 128).float()
    assert torch.equal(a_p, a[perm])


def test_full_size_properties_config5(ops):
    """BASELINE configs[4] geometry (2048^2 -> 256/128/64/32, 87 040 queries) at N = 4: the same size-independent checks
    as cfg3, plus agreement between the plain and the fused entry points on a constant value field."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg5_2048_bf16"]
    dev = "cuda:0"
    N = 4
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], N, torch.bfloat16, device=dev, seed=5)
    f = ops.MSDeformAttnFunction.apply
    lo_in = loc.clamp(0.02, 0.98).contiguous()
    ones = torch.ones_like(v)
    assert float((f(ones, ss, lsi, lo_in, attn, 128).float() - 1).abs().max()) < 1e-2
    vv = ones.clone().requires_grad_(True)
    go = torch.randn(N, loc.shape[1], 256, device=dev, dtype=torch.bfloat16)
    f(vv, ss, lsi, lo_in, attn, 128).backward(go)
    assert rel_to_max(vv.grad.float().sum(1), go.float().view(N, -1, 8, 32).sum(1)) < 2e-2
    a = f(v, ss, lsi, loc, attn, 128)
    perm = torch.randperm(N, device=dev)
    assert torch.equal(f(v[perm].contiguous(), ss, lsi, loc[perm].contiguous(), attn[perm].contiguous(), 128), a[perm])
    # query permutation equivariance (rows are independent): bit-exact forward, gradients of the permuted rows follow
    qperm = torch.randperm(loc.shape[1], device=dev)
    assert torch.equal(f(v, ss, lsi, loc[:, qperm].contiguous(), attn[:, qperm].contiguous(), 128), a[:, qperm])


# ---------------------------------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------------------------------
def test_error_behaviour_matches_upstream(ops):
    dev = "cuda:0"
    value, ss, lsi, loc, attn, go = random_problem(6, 2, 32, 5, [(4, 4)], 2, seed=5, dtype=torch.float32)
    f = ops.MSDeformAttnFunction.apply
    args = [value.to(dev), ss.to(dev), lsi.to(dev), loc.to(dev), attn.to(dev)]
    f(*args, 128)                      # min(N, step) = 6
    f(*args, 3)
    with pytest.raises(RuntimeError, match="im2col_step"):
        f(*args, 4)
    with pytest.raises(RuntimeError, match="contiguous"):
        f(args[0].transpose(0, 1).contiguous().transpose(0, 1), *args[1:], 128)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        f(args[0].cpu(), *args[1:], 128)


def test_forward_is_deterministic_and_stream_safe(ops):
    dev = "cuda:0"
    value, ss, lsi, loc, attn, go = random_problem(2, 8, 32, 301, [(16, 16), (8, 8)], 4, seed=8, dtype=torch.float32)
    f = ops.MSDeformAttnFunction.apply
    args = [value.to(dev), ss.to(dev), lsi.to(dev), loc.to(dev), attn.to(dev)]
    a = f(*args, 128)
    b = f(*args, 128)
    assert torch.equal(a, b)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        c = f(*args, 128)
    s.synchronize()
    assert torch.equal(a, c)


def test_grad_loc_and_grad_attn_are_deterministic(ops):
    value, ss, lsi, loc, attn, go = random_problem(2, 8, 32, 64, [(8, 8), (4, 4)], 4, seed=12, dtype=torch.float32)
    r1 = run_cuda(ops, value, ss, lsi, loc, attn, go, torch.float32)
    r2 = run_cuda(ops, value, ss, lsi, loc, attn, go, torch.float32)
    assert torch.equal(r1[2], r2[2]) and torch.equal(r1[3], r2[3])


def _with_flags(flags, fn):
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    old = MSDA.backward_flags
    try:
        MSDA.backward_flags = flags
        return fn()
    finally:
        MSDA.backward_flags = old


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("gscale", [1.0, 1e-6, 3e4])
def test_fp16_bucket_accumulation_any_gradient_magnitude(ops, dtype, gscale):
    """Default 16-bit backward: grad_value accumulates in a scaled, bucketed fp16 buffer.  It must hold the 2e-2
    gate for any magnitude of grad_output (the scale is derived on the device), agree with the fp32-accumulation
    mode, and change nothing else."""
    if dtype == torch.float16 and gscale > 100:
        gscale = 50.0          # keep grad_value itself inside the fp16 range of the output dtype
    value, ss, lsi, loc, attn, go = random_problem(2, 8, 32, 2000, [(16, 16), (8, 8), (2, 2)], 4, seed=21)
    go = go * gscale
    fast = run_cuda(ops, value, ss, lsi, loc, attn, go, dtype)
    exact = _with_flags(2, lambda: run_cuda(ops, value, ss, lsi, loc, attn, go, dtype))
    assert torch.equal(exact[0], fast[0]) and torch.equal(exact[2], fast[2]) and torch.equal(exact[3], fast[3])
    want = oracle_on_rounded_inputs(value, ss, loc, attn, go, dtype)
    assert torch.isfinite(fast[1]).all()
    assert rel_to_max(fast[1], want[1]) < 2e-2
    assert rel_to_max(exact[1], want[1]) < 2e-2
    assert rel_to_max(fast[1], want[1]) < 2.5 * max(rel_to_max(exact[1], want[1]), 2e-3)


@pytest.mark.parametrize("mean", [0.0, 1.0])
def test_fp16_bucket_accumulation_deep_reductions(ops, mean):
    """2 000 queries x 4 points onto a 2x2 level = ~2 000 adds per element: the bucket layout must keep the error
    near the fp32-accumulation level even for same-sign gradients; one single bucket (depth 65535) must not."""
    value, ss, lsi, loc, attn, go = random_problem(1, 8, 32, 2000, [(2, 2)], 4, seed=33, loc_range=(0.0, 1.0))
    go = go + mean
    want = oracle_on_rounded_inputs(value, ss, loc, attn, go, torch.bfloat16)[1]
    e_default = rel_to_max(run_cuda(ops, value, ss, lsi, loc, attn, go, torch.bfloat16)[1], want)
    e_exact = rel_to_max(_with_flags(2, lambda: run_cuda(ops, value, ss, lsi, loc, attn, go, torch.bfloat16))[1], want)
    e_single = rel_to_max(_with_flags(65535 << 8, lambda: run_cuda(ops, value, ss, lsi, loc, attn, go, torch.bfloat16))[1], want)
    assert e_exact < 5e-3 and e_default < 8e-3, (e_exact, e_default, e_single)
    assert e_default < e_single or e_single < 5e-3, (e_default, e_single)


def test_fp16_bucket_accumulation_cannot_overflow(ops):
    """Adversarial: every sampling point of every query hits the same pixel with weight 1 and every grad_output
    element is the same large value; the device-derived scale must keep the fp16 accumulators finite."""
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    dev = "cuda:0"
    N, M, D, Lq, L, P = 1, 8, 32, 6000, 1, 4
    ss = torch.tensor([(4, 4)], dtype=torch.long)
    value = torch.ones(N, 16, M, D, dtype=torch.bfloat16, device=dev)
    loc = torch.full((N, Lq, M, L, P, 2), 0.375, device=dev)          # exact centre of pixel (1, 1)
    attn = torch.full((N, Lq, M, L, P), 0.25, device=dev)
    go = torch.full((N, Lq, M * D), 1000.0, dtype=torch.bfloat16, device=dev)
    for flags in (0, 8 << 8):
        gv, gl, ga = _with_flags(flags, lambda: MSDA.ms_deform_attn_backward(value, ss.to(dev), lsi_of(ss).to(dev), loc, attn, go, 128))
        assert torch.isfinite(gv).all()
        want = 1000.0 * Lq                                                 # all of it lands on pixel (1, 1)
        got = gv.float().view(16, M, D)[5]
        # thousands of *identical* addends are the worst case of round-to-nearest accumulation (every add rounds
        # the same way): the bound here is finiteness plus a few percent, not the 2e-2 gate of realistic inputs
        assert float((got - want).abs().max()) / want < 5e-2
        assert float(gv.float().view(16, M, D)[[0, 1, 2, 3, 4, 6, 7]].abs().max()) == 0.0


def test_bf16_aux_inputs_are_accepted_and_grads_keep_their_dtype(ops):
    dev = "cuda:0"
    value, ss, lsi, loc, attn, go = random_problem(1, 8, 32, 10, [(8, 8)], 4, seed=31, dtype=torch.float32)
    v = value.to(dev, torch.bfloat16).requires_grad_(True)
    lo = loc.to(dev, torch.bfloat16).requires_grad_(True)
    at = attn.to(dev, torch.bfloat16).requires_grad_(True)
    out = ops.MSDeformAttnFunction.apply(v, ss.to(dev), lsi.to(dev), lo, at, 64)
    out.sum().backward()
    assert lo.grad.dtype == torch.bfloat16 and at.grad.dtype == torch.bfloat16 and v.grad.dtype == torch.bfloat16


def test_host_pipeline_matches_device_resident_path(ops):
    """HostPipeline (pinned host operands, 3 streams, double-buffered) == direct extension-level calls."""
    from vision_instance_seg_b200.host_pipeline import HostPipeline
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA
    dev = torch.device("cuda:0")
    problems = [random_problem(2, 8, 32, 150, [(16, 16), (8, 8)], 4, seed=40 + i, dtype=torch.float32) for i in range(5)]
    ss, lsi = problems[0][1], problems[0][2]
    pipe = HostPipeline(ss, lsi, dev)
    outs = []
    for value, _, _, loc, attn, go in problems:
        h_in = [t.to(dt).pin_memory() for t, dt in ((value, torch.bfloat16), (loc, torch.float32), (attn, torch.float32), (go, torch.bfloat16))]
        h_out = [torch.empty(2, 150, 256, dtype=torch.bfloat16).pin_memory(), torch.empty(value.shape, dtype=torch.bfloat16).pin_memory(),
                 torch.empty(loc.shape, dtype=torch.float32).pin_memory(), torch.empty(attn.shape, dtype=torch.float32).pin_memory()]
        pipe.submit(h_in, h_out)
        outs.append((h_in, h_out))
    pipe.synchronize()
    assert pipe.h2d_bytes == sum(t.numel() * t.element_size() for h_in, _ in outs for t in h_in)
    for h_in, h_out in outs:
        d = [t.to(dev) for t in h_in]
        out = MSDA.ms_deform_attn_forward(d[0], ss.to(dev), lsi.to(dev), d[1], d[2], 128)
        gv, gl, ga = MSDA.ms_deform_attn_backward(d[0], ss.to(dev), lsi.to(dev), d[1], d[2], d[3], 128)
        assert torch.equal(out.cpu(), h_out[0])
        assert rel_to_max(h_out[1], gv) < 1e-2          # fp32 atomics: order-dependent in the last bits, then bf16 rounding
        assert torch.equal(gl.cpu(), h_out[2]) and torch.equal(ga.cpu(), h_out[3])


# ---------------------------------------------------------------------------------------------------
# round 2: non-finite values, full-size per-image comparison
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_points_that_fail_the_gate_never_touch_value(ops, dtype, bad):
    """Upstream skips a sampling point unless -1 < h_im < H and -1 < w_im < W; it never loads a pixel for it.  Plant
    NaN / Inf at the four corner pixels of every level -- the pixels a clamped address of an outside point would hit --
    keep every valid footprint away from them, and send half of the points far outside: all four results must stay
    finite and equal to the oracle's (which runs on the same planted values)."""
    D = 32 if dtype != torch.float64 else 6
    value, ss, lsi, loc, attn, go = random_problem(2, 4, D, 40, [(12, 10), (6, 7), (5, 5)], 4, seed=21, loc_range=(0.3, 0.7))
    g = torch.Generator().manual_seed(2)
    outside = torch.rand(loc.shape[:-1], generator=g) < 0.5
    far = torch.where(torch.rand(loc.shape, generator=g) < 0.5, torch.full_like(loc, -0.75), torch.full_like(loc, 1.75))
    loc = torch.where(outside[..., None], far, loc)
    value = value.clone()
    for (H, W), start in zip(ss.tolist(), lsi.tolist()):
        for y, x in ((0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1)):
            value[:, start + y * W + x] = bad
    got = run_cuda(ops, value, ss, lsi, loc, attn, go, dtype)
    want = oracle_on_rounded_inputs(value, ss, loc, attn, go, dtype)
    for name, a, b in zip(("out", "grad_value", "grad_loc", "grad_attn"), got, want):
        assert torch.isfinite(b).all(), f"oracle {name} not finite: the test inputs are wrong"
        assert torch.isfinite(a).all(), f"{name}: non-finite values leaked from pixels no valid point samples"
        assert rel_to_max(a, b) < TOL[dtype], name


def _compare_per_image(ops, v, ss, lsi, loc, attn, go, smooth_eps=1e-3):
    """CUDA on the whole batch, oracle image by image (fp32 on the CPU, seconds per image)."""
    got = run_cuda(ops, v.float(), ss, lsi, loc, attn, go, torch.bfloat16)
    wh = torch.stack([ss[:, 1], ss[:, 0]], -1).to(torch.float32)[None, None, None, :, None, :]
    worst = {}
    for b in range(v.shape[0]):
        want = ms_deform_attn_oracle_grads(v[b:b + 1].float(), ss, loc[b:b + 1], attn[b:b + 1],
                                           go[b:b + 1].to(torch.bfloat16).float(), dtype=torch.float32)
        px = loc[b:b + 1] * wh - 0.5
        smooth = ((px - px.round()).abs() > smooth_eps).all(-1, keepdim=True).expand_as(loc[b:b + 1])
        for name, a, w in zip(("out", "grad_value", "grad_loc", "grad_attn"), got, want):
            a, w = a[b:b + 1].detach().float().cpu(), w.float()
            if name == "grad_loc":          # away from integer pixel lines, where the bilinear derivative jumps
                a, w = a * smooth, w * smooth
            worst[name] = max(worst.get(name, 0.0), rel_to_max(a, w))
    return worst


def test_config3_full_batch_against_the_oracle_per_image(ops):
    """BASELINE.json configs[2] at its full size (N = 16, 21 760 queries, bf16): every image against the oracle."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg3_swinl_1024_bf16"]
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], cfg["batch"], torch.bfloat16, device="cpu", seed=99)
    go = torch.randn(cfg["batch"], loc.shape[1], 256, generator=torch.Generator().manual_seed(9))
    worst = _compare_per_image(ops, v, ss, lsi, loc, attn, go)
    for name, err in worst.items():
        assert err < 2e-2, f"{name}: {err:.3e}"


def test_config4_full_batch_against_the_oracle_per_image(ops):
    """BASELINE.json configs[3] at its full size (N = 16, 300 box queries, bf16), default backward mode."""
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS["cfg4_decoder_300q_bf16"]
    v, ss, lsi, loc, attn = W.make_decoder_inputs(cfg["shapes"], cfg["batch"], torch.bfloat16, queries=300, device="cpu", seed=5)
    go = torch.randn(cfg["batch"], 300, 256, generator=torch.Generator().manual_seed(10))
    worst = _compare_per_image(ops, v, ss, lsi, loc, attn, go)
    for name, err in worst.items():
        assert err < 2e-2, f"{name}: {err:.3e}"


@pytest.mark.parametrize("dtype,D,L,P", [(torch.bfloat16, 16, 2, 80), (torch.float32, 32, 3, 200), (torch.float16, 128, 1, 700)])
def test_many_points_per_query_run_on_the_compatibility_kernels(ops, dtype, D, L, P):
    """the vector kernels stage a warp's L*P rows in shared memory, which caps L*P (~130 for 16-bit D = 16,
    ~530 for fp32 D = 32); upstream has no such limit, so larger L*P must run (on the any-D kernels), not fail."""
    shapes = [(9, 7), (5, 4), (3, 3)][:L]
    check(ops, random_problem(1, 2, D, 5, shapes, P, seed=31), dtype, tol=TOL[dtype] if dtype != torch.float16 else 5e-3)
