"""CPU: the detectron2-free encoder harness — naming, sizes and call-site metadata.
No operator call happens here (the product has no CPU path); the GPU run of the harness is in test_gpu_encoder.py."""
import pytest
import torch

import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import workloads as W
from vision_instance_seg_b200.distributed import ENCODER_GRAD_ELEMENTS
from vision_instance_seg_b200.modules.encoder import (MSDeformAttnTransformerEncoder, MSDeformAttnTransformerEncoderLayer,
                                                      MSDeformAttnTransformerEncoderOnly, encoder_state_dict_keys)


def test_parameter_names_and_count_match_maskdino_swinl_settings():
    enc = MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=2048,
                                             dropout=0.0, num_feature_levels=4, enc_n_points=4)
    sd = enc.state_dict()
    assert sorted(sd.keys()) == sorted(encoder_state_dict_keys(6))
    assert sum(p.numel() for p in enc.parameters()) == 6 * 1282176 + 1024 == ENCODER_GRAD_ELEMENTS
    layer = enc.encoder.layers[0]
    assert sum(p.numel() for p in layer.self_attn.parameters()) == 230272
    assert isinstance(layer.self_attn, pkg.MSDeformAttn) and layer.self_attn.im2col_step == 128


def test_checkpoint_with_upstream_prefix_loads():
    enc = MSDeformAttnTransformerEncoderOnly(d_model=64, nhead=4, num_encoder_layers=2, dim_feedforward=128,
                                             num_feature_levels=3, enc_n_points=2)
    prefix = "sem_seg_head.pixel_decoder.transformer."
    ckpt = {prefix + k: torch.randn_like(v) for k, v in enc.state_dict().items()}
    ckpt["sem_seg_head.pixel_decoder.input_proj.0.0.weight"] = torch.zeros(1)       # unrelated keys are ignored by the strip
    stripped = {k[len(prefix):]: v for k, v in ckpt.items() if k.startswith(prefix)}
    missing, unexpected = enc.load_state_dict(stripped, strict=True)
    assert not missing and not unexpected
    assert torch.equal(enc.encoder.layers[1].self_attn.value_proj.weight, ckpt[prefix + "encoder.layers.1.self_attn.value_proj.weight"])


def test_layers_are_independent_copies_with_upstream_init():
    enc = MSDeformAttnTransformerEncoderOnly(d_model=64, nhead=4, num_encoder_layers=3, dim_feedforward=128,
                                             num_feature_levels=3, enc_n_points=2)
    a, b = enc.encoder.layers[0], enc.encoder.layers[2]
    assert a.linear1.weight.data_ptr() != b.linear1.weight.data_ptr()
    for layer in enc.encoder.layers:        # MSDeformAttn._reset_parameters wins over the blanket xavier init
        assert float(layer.self_attn.sampling_offsets.weight.detach().abs().max()) == 0.0
        assert float(layer.self_attn.attention_weights.weight.detach().abs().max()) == 0.0
        bias = layer.self_attn.sampling_offsets.bias.view(4, 3, 2, 2)
        assert torch.allclose(bias[:, :, 1], 2 * bias[:, :, 0])


def test_valid_ratio_and_reference_points():
    mask = torch.zeros(2, 8, 10, dtype=torch.bool)
    mask[1, 6:, :] = True
    mask[1, :, 5:] = True
    vr = MSDeformAttnTransformerEncoderOnly.get_valid_ratio(mask)
    assert torch.allclose(vr, torch.tensor([[1.0, 1.0], [0.5, 0.75]]))
    ss = W.make_spatial_shapes([(8, 10), (4, 5)])
    ratios = torch.stack([vr, vr], 1)                                   # (N, L, 2)
    ref = MSDeformAttnTransformerEncoder.get_reference_points(ss, ratios, "cpu")
    assert ref.shape == (2, 8 * 10 + 4 * 5, 2, 2)
    assert torch.equal(ref, W.get_reference_points(ss, ratios))          # same builder as the synthetic workloads use
    # image 0 is unpadded: the reference point of pixel (y, x) of level 0 is its centre in [0, 1]
    assert torch.allclose(ref[0, 0 * 10 + 3, 0], torch.tensor([3.5 / 10, 0.5 / 8]))
    # image 1: centres are normalised by the valid extent then scaled back by the sampled level's valid ratio
    assert torch.allclose(ref[1, 2 * 10 + 1, 1], torch.tensor([1.5 / (0.5 * 10) * 0.5, 2.5 / (0.75 * 8) * 0.75]))


def test_flattening_metadata_without_operator_call():
    """forward() builds spatial_shapes / level_start_index / masks exactly as the op expects; the encoder stack itself is
    stubbed so that no kernel is needed."""
    enc = MSDeformAttnTransformerEncoderOnly(d_model=32, nhead=4, num_encoder_layers=1, dim_feedforward=64,
                                             num_feature_levels=3, enc_n_points=2)
    seen = {}

    class FakeEncoder(torch.nn.Module):
        def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None, level_embed=None, spatial_shapes_list=None):
            seen.update(src=src, ss=spatial_shapes, lsi=level_start_index, vr=valid_ratios, pos=pos, mask=padding_mask)
            return src

    enc.encoder = FakeEncoder()
    shapes = [(12, 20), (6, 10), (3, 5)]
    srcs = [torch.randn(2, 32, h, w) for h, w in shapes]
    pos = [torch.randn(2, 32, h, w) for h, w in shapes]
    masks = [torch.zeros(2, h, w, dtype=torch.bool) for h, w in shapes]
    masks[0][1, :, 15:] = True
    masks[1][1, :, 8:] = True      # ceil-style down-sampling of the padding, as interpolate would give
    masks[2][1, :, 4:] = True
    memory, ss, lsi = enc(srcs, masks, pos)
    assert ss.tolist() == [list(s) for s in shapes] and ss.dtype == torch.long
    assert lsi.tolist() == [0, 240, 300]
    assert memory.shape == (2, 240 + 60 + 15, 32)
    assert seen["mask"] is not None and seen["mask"].shape == (2, 315) and int(seen["mask"].sum()) == 12 * 5 + 6 * 2 + 3 * 1
    assert torch.allclose(seen["vr"][1], torch.tensor([[0.75, 1.0], [0.8, 1.0], [0.8, 1.0]]))
    assert torch.allclose(seen["pos"][:, :240], pos[0].flatten(2).transpose(1, 2) + enc.level_embed[0])
    # sizes that are multiples of 32 skip the masks altogether (upstream's enable_mask rule)
    srcs32 = [torch.randn(1, 32, 32, 64), torch.randn(1, 32, 32, 32)]
    enc2 = MSDeformAttnTransformerEncoderOnly(d_model=32, nhead=4, num_encoder_layers=1, dim_feedforward=64,
                                              num_feature_levels=2, enc_n_points=2)
    enc2.encoder = FakeEncoder()
    enc2(srcs32, [torch.ones(1, 32, 64, dtype=torch.bool), torch.ones(1, 32, 32, dtype=torch.bool)], [torch.zeros_like(s) for s in srcs32])
    assert seen["mask"] is None and torch.equal(seen["vr"], torch.ones(1, 2, 2))


def test_cpu_forward_raises_without_fallback():
    enc = MSDeformAttnTransformerEncoderOnly(d_model=32, nhead=4, num_encoder_layers=1, dim_feedforward=64,
                                             num_feature_levels=1, enc_n_points=2)
    src = [torch.randn(1, 32, 4, 4)]
    with pytest.raises(RuntimeError):
        enc(src, None, [torch.zeros_like(src[0])])
