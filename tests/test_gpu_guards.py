"""GPU: guard-band checks through the raw C ABI (compute-sanitizer is not available on the pool).  Every output / scratch
buffer is carved out of a larger canary-filled allocation; after the call the canaries on both sides must be intact, for
ragged sizes that exercise the tail warps / tail rows of every kernel family."""
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096          # bytes on each side


@pytest.fixture(scope="module")
def lib(built_library):
    assert torch.cuda.is_available()
    import vision_instance_seg_b200 as pkg
    return pkg.load_library()


class Guarded:
    """`nbytes` usable bytes (16-byte aligned) between two canary bands."""

    def __init__(self, nbytes, fill=0xA5):
        self.nbytes = nbytes
        self.raw = torch.full((nbytes + 2 * GUARD,), fill, dtype=torch.uint8, device="cuda")
        self.fill = fill

    @property
    def ptr(self):
        return self.raw.data_ptr() + GUARD

    def view(self, dtype, shape):
        return self.raw[GUARD:GUARD + self.nbytes].view(dtype).view(shape)

    def intact(self):
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.nbytes:]
        return bool((lo == self.fill).all()) and bool((hi == self.fill).all())


DT = {torch.float32: 0, torch.float64: 1, torch.bfloat16: 2, torch.float16: 3}


@pytest.mark.parametrize("dtype,D", [(torch.float32, 32), (torch.bfloat16, 32), (torch.float16, 64), (torch.bfloat16, 16),
                                     (torch.float32, 30), (torch.bfloat16, 71), (torch.float64, 8), (torch.bfloat16, 128)])
@pytest.mark.parametrize("flags", [0, 2])
def test_operator_stays_inside_its_buffers(lib, dtype, D, flags):
    from vision_instance_seg_b200 import workloads as W
    shapes = [(11, 7), (5, 3), (2, 2)]
    N, M, Lq, L, P = 3, 3, 29, 3, 4               # 261 pairs: ragged last warp for every lane-group width
    v, ss, lsi, loc, attn = W.make_uniform_inputs(shapes, N, dtype, queries=Lq, n_heads=M, head_dim=D, device="cuda")
    aux = torch.float64 if dtype == torch.float64 else torch.float32
    loc, attn = loc.to(aux).contiguous(), attn.to(aux).contiguous()
    S = v.shape[1]
    es, ea = v.element_size(), loc.element_size()
    out = Guarded(N * Lq * M * D * es)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.msda_forward(v.data_ptr(), ss.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(), out.ptr,
                          N, S, M, D, Lq, L, P, DT[dtype], 64, st)
    assert rc == 0
    go = torch.randn(N, Lq, M * D, device="cuda").to(dtype)
    gv, gl, ga = Guarded(v.numel() * es), Guarded(loc.numel() * ea), Guarded(attn.numel() * ea)
    nscratch = lib.msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, DT[dtype], flags)
    scratch = Guarded(max(nscratch, 16))
    rc = lib.msda_backward(v.data_ptr(), ss.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(), go.data_ptr(),
                           gv.ptr, gl.ptr, ga.ptr, scratch.ptr if nscratch else None, nscratch,
                           N, S, M, D, Lq, L, P, DT[dtype], 64, flags, st)
    assert rc == 0
    torch.cuda.synchronize()
    for name, g in (("output", out), ("grad_value", gv), ("grad_loc", gl), ("grad_attn", ga), ("scratch", scratch)):
        assert g.intact(), f"{name}: canary overwritten"
    # and the results written inside are complete (no canary bytes left in fully-overwritten outputs)
    assert torch.isfinite(out.view(dtype, (N, Lq, M * D)).float()).all()
    assert torch.isfinite(gl.view(aux, tuple(loc.shape)).float()).all()
    assert torch.isfinite(gv.view(dtype, tuple(v.shape)).float()).all()


@pytest.mark.parametrize("aux16", [False, True])
@pytest.mark.parametrize("R", [2, 4])
def test_fused_operator_stays_inside_its_buffers(lib, aux16, R):
    shapes = torch.tensor([(11, 7), (5, 3), (2, 2)], dtype=torch.long, device="cuda")
    lsi = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    N, M, D, Lq, L, P = 3, 3, 32, 29, 3, 4
    dtype = torch.bfloat16
    adt = torch.bfloat16 if aux16 else torch.float32
    v = torch.randn(N, S, M, D, device="cuda").to(dtype)
    ref = torch.rand(N, Lq, L, R, device="cuda")
    off = torch.randn(N, Lq, M, L, P, 2, device="cuda").to(adt)
    lg = torch.randn(N, Lq, M, L * P, device="cuda").to(adt)
    go = torch.randn(N, Lq, M * D, device="cuda").to(dtype)
    st = torch.cuda.current_stream().cuda_stream
    out = Guarded(N * Lq * M * D * 2)
    rc = lib.msda_fused_forward(v.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), ref.data_ptr(), R, off.data_ptr(), lg.data_ptr(),
                                out.ptr, N, S, M, D, Lq, L, P, DT[dtype], DT[adt], 64, st)
    assert rc == 0
    gv, goff, glg = Guarded(v.numel() * 2), Guarded(off.numel() * off.element_size()), Guarded(lg.numel() * lg.element_size())
    nscratch = lib.msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, DT[dtype], 0)
    scratch = Guarded(nscratch)
    rc = lib.msda_fused_backward(v.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), ref.data_ptr(), R, off.data_ptr(), lg.data_ptr(),
                                 go.data_ptr(), gv.ptr, goff.ptr, glg.ptr, scratch.ptr, nscratch,
                                 N, S, M, D, Lq, L, P, DT[dtype], DT[adt], 64, 0, st)
    assert rc == 0
    torch.cuda.synchronize()
    for name, g in (("output", out), ("grad_value", gv), ("grad_offsets", goff), ("grad_logits", glg), ("scratch", scratch)):
        assert g.intact(), f"{name}: canary overwritten"
    assert torch.isfinite(goff.view(adt, tuple(off.shape)).float()).all() and torch.isfinite(glg.view(adt, tuple(lg.shape)).float()).all()


@pytest.mark.parametrize("rows,C", [(1, 128), (37, 256), (1031, 384), (7, 1024)])
def test_encoder_glue_kernels_stay_inside_their_buffers(lib, rows, C):
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(rows, C, device="cuda")
    d16 = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    y, y16, mean, rstd = Guarded(rows * C * 4), Guarded(rows * C * 2), Guarded(rows * 4), Guarded(rows * 4)
    assert lib.msda_enc_add_layernorm_forward(x.data_ptr(), d16.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.ptr, y16.ptr,
                                              mean.ptr, rstd.ptr, rows, C, 1e-5, st) == 0
    dx, dd, dg, db = Guarded(rows * C * 4), Guarded(rows * C * 2), Guarded(C * 4), Guarded(C * 4)
    part = Guarded(lib.msda_enc_add_layernorm_backward_scratch_bytes(C))
    gy = torch.randn(rows, C, device="cuda")
    assert lib.msda_enc_add_layernorm_backward(gy.data_ptr(), d16.data_ptr(), x.data_ptr(), d16.data_ptr(), mean.ptr, rstd.ptr,
                                               gamma.data_ptr(), dx.ptr, dd.ptr, dg.ptr, db.ptr, part.ptr, part.nbytes,
                                               rows, C, st) == 0
    q16 = Guarded(rows * C * 2)
    assert lib.msda_enc_add_cast(x.data_ptr(), gy.data_ptr(), q16.ptr, rows * C, st) == 0
    cs, cscr = Guarded(C * 4), Guarded(lib.msda_enc_colsum_scratch_bytes(C))
    assert lib.msda_enc_colsum(d16.data_ptr(), cs.ptr, cscr.ptr, cscr.nbytes, 1, rows, 0, rows, C, st) == 0
    g16 = Guarded(rows * C * 2)
    g16.view(torch.bfloat16, (rows, C)).copy_(torch.randn(rows, C, device="cuda").to(torch.bfloat16))
    cs2 = Guarded(C * 4)
    assert lib.msda_enc_relu_bwd_colsum(g16.ptr, d16.data_ptr(), cs2.ptr, cscr.ptr, cscr.nbytes, rows, C, st) == 0
    torch.cuda.synchronize()
    for name, g in (("y", y), ("y16", y16), ("mean", mean), ("rstd", rstd), ("dx", dx), ("ddelta", dd), ("dgamma", dg), ("dbeta", db),
                    ("partials", part), ("q16", q16), ("colsum", cs), ("colsum scratch", cscr), ("relu grad", g16), ("relu colsum", cs2)):
        assert g.intact(), f"{name}: canary overwritten"
    assert torch.isfinite(y.view(torch.float32, (rows, C))).all() and torch.isfinite(dg.view(torch.float32, (C,))).all()
