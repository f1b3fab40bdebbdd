"""GPU: the encoder harness end to end (6-layer-style stack of MSDeformAttn self-attention + FFN, the calling pattern
the reference triggers through build_model(cfg)) against the same modules on the CPU with the oracle core patched in."""
import copy

import pytest
import torch

from oracle import ms_deform_attn_core_pytorch
from tests.helpers import rel_to_max

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_library):
    assert torch.cuda.is_available()
    import vision_instance_seg_b200 as pkg
    pkg.load_library()
    return pkg


class _OracleFunction:
    """Stands in for MSDeformAttnFunction on the CPU copy: torch autograd differentiates the oracle core."""

    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return ms_deform_attn_core_pytorch(value, shapes, loc, attn)


def _inputs(shapes, N, C, seed, padded):
    g = torch.Generator().manual_seed(seed)
    srcs = [torch.randn(N, C, h, w, generator=g) for h, w in shapes]
    pos = [torch.randn(N, C, h, w, generator=g) * 0.1 for h, w in shapes]
    masks = [torch.zeros(N, h, w, dtype=torch.bool) for h, w in shapes]
    if padded:
        for m in masks:
            m[1, :, (m.shape[2] * 3) // 4:] = True
            m[1, (m.shape[1] * 2) // 3:, :] = True
    return srcs, masks, pos


@pytest.mark.parametrize("padded", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_encoder_matches_cpu_restatement(ops, monkeypatch, padded, fused):
    from vision_instance_seg_b200.modules import encoder as E
    from vision_instance_seg_b200.modules import ms_deform_attn as MOD
    torch.manual_seed(7)
    enc_cpu = E.MSDeformAttnTransformerEncoderOnly(d_model=64, nhead=4, num_encoder_layers=2, dim_feedforward=128,
                                                   dropout=0.0, num_feature_levels=3, enc_n_points=4)
    with torch.no_grad():
        for layer in enc_cpu.encoder.layers:      # leave the zero init so that offsets / weights depend on the query
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
            layer.self_attn.attention_weights.weight.normal_(0, 0.3)
    enc_gpu = copy.deepcopy(enc_cpu).cuda()
    assert ops.set_fused_preop(enc_gpu, fused) == 2
    shapes = [(12, 20), (6, 10), (3, 5)]          # not multiples of 32: the padding masks are honoured
    srcs, masks, pos = _inputs(shapes, 2, 64, 3, padded)

    monkeypatch.setattr(MOD, "MSDeformAttnFunction", _OracleFunction)
    srcs_c = [s.clone().requires_grad_(True) for s in srcs]
    mem_c, ss_c, lsi_c = enc_cpu(srcs_c, masks, pos)
    g = torch.randn_like(mem_c)
    mem_c.backward(g)
    monkeypatch.undo()

    srcs_g = [s.cuda().requires_grad_(True) for s in srcs]
    mem_g, ss_g, lsi_g = enc_gpu(srcs_g, [m.cuda() for m in masks], [p.cuda() for p in pos])
    mem_g.backward(g.cuda())
    torch.cuda.synchronize()
    assert ss_g.tolist() == ss_c.tolist() and lsi_g.tolist() == lsi_c.tolist()
    assert rel_to_max(mem_g, mem_c) < 1e-4
    for a, b in zip(srcs_g, srcs_c):
        assert rel_to_max(a.grad, b.grad) < 1e-4
    for (name, pg), (_, pc) in zip(enc_gpu.named_parameters(), enc_cpu.named_parameters()):
        assert rel_to_max(pg.grad, pc.grad) < 2e-4, name


def test_encoder_bf16_autocast_runs_the_16bit_kernels(ops):
    from vision_instance_seg_b200.modules import encoder as E
    torch.manual_seed(1)
    enc = E.MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=2, dim_feedforward=512,
                                               dropout=0.0, num_feature_levels=3, enc_n_points=4).cuda()
    shapes = [(32, 32), (16, 16), (8, 8)]
    srcs, masks, pos = _inputs(shapes, 2, 256, 5, False)
    srcs = [s.cuda() for s in srcs]
    pos = [p.cuda() for p in pos]
    lib = ops.load_library()
    ref, _, _ = enc(srcs, None, pos)
    n0 = lib.msda_total_launch_count()
    for fused in (False, True):
        ops.set_fused_preop(enc, fused)
        enc.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            mem, _, _ = enc(srcs, None, pos)
        mem.float().square().mean().backward()
        torch.cuda.synchronize()
        assert mem.dtype == torch.bfloat16 or mem.dtype == torch.float32
        assert rel_to_max(mem.float(), ref) < 5e-2
        assert all(torch.isfinite(p.grad).all() for p in enc.parameters() if p.grad is not None)
    assert lib.msda_total_launch_count() - n0 >= 2 * 2 * 2      # our kernels ran (2 layers x fwd+bwd x 2 modes)
