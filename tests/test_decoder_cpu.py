"""CPU: metadata of the decoder harness (decoder call site) -- names, parameter counts, the sine
embedding and the reference-box arithmetic.  The sampling core needs a GPU (tests/test_gpu_decoder.py)."""
import math

import torch

from vision_instance_seg_b200.modules import (MSDeformAttn, build_decoder, gen_sineembed_for_position,
                                              set_shared_value_proj)
from vision_instance_seg_b200 import workloads as W


def test_decoder_layout_and_parameter_names():
    dec = build_decoder(num_decoder_layers=9)
    assert len(dec.layers) == 9
    names = {n for n, _ in dec.named_parameters()}
    for i in (0, 8):
        for lin in ("sampling_offsets", "attention_weights", "value_proj", "output_proj"):
            assert f"layers.{i}.cross_attn.{lin}.weight" in names and f"layers.{i}.cross_attn.{lin}.bias" in names
        for n in ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "norm1.weight", "norm2.weight", "norm3.weight",
                  "linear1.weight", "linear2.bias"):
            assert f"layers.{i}.{n}" in names
    assert "ref_point_head.layers.0.weight" in names and "ref_point_head.layers.1.bias" in names and "norm.weight" in names
    per_layer = sum(p.numel() for p in dec.layers[0].parameters())
    # cross_attn 230 272 + self_attn 263 168 + 3 LayerNorms 1 536 + FFN (256*2048 + 2048 + 2048*256 + 256)
    assert per_layer == 230272 + 263168 + 1536 + 2 * 256 * 2048 + 2048 + 256
    assert all(isinstance(l.cross_attn, MSDeformAttn) for l in dec.layers)


def test_sine_embedding_and_reference_boxes():
    boxes = torch.rand(5, 2, 4)
    e = gen_sineembed_for_position(boxes)
    assert e.shape == (5, 2, 512)
    # first block embeds y, second x (DAB-DETR order); channel 0 is sin(2*pi*coord / 10000^0)
    assert torch.allclose(e[:, :, 0], torch.sin(boxes[:, :, 1] * 2 * math.pi), atol=1e-6)
    assert torch.allclose(e[:, :, 128], torch.sin(boxes[:, :, 0] * 2 * math.pi), atol=1e-6)
    assert torch.allclose(e[:, :, 129], torch.cos(boxes[:, :, 0] * 2 * math.pi), atol=1e-6)
    assert gen_sineembed_for_position(boxes[..., :2]).shape == (5, 2, 256)
    # the boxes every layer hands to cross_attn: reference * cat(valid_ratios, valid_ratios), batch-first here
    vr = torch.rand(2, 4, 2) * 0.5 + 0.5
    got = W.decoder_reference_points_input(boxes.transpose(0, 1), vr)
    want = (boxes[:, :, None] * torch.cat([vr, vr], -1)[None, :]).transpose(0, 1)
    assert torch.allclose(got, want)


def test_set_shared_value_proj_wires_every_cross_attention_in_layer_order():
    dec = build_decoder(num_decoder_layers=4, dim_feedforward=64)
    keys = list(dec.state_dict().keys())
    proj = set_shared_value_proj(dec)
    assert proj.K == 4 and [l.cross_attn._stacked_value for l in dec.layers] == [(proj, i) for i in range(4)]
    assert list(dec.state_dict().keys()) == keys                   # nothing registered, checkpoints unaffected
    set_shared_value_proj(dec, False)
    assert all(l.cross_attn._stacked_value is None for l in dec.layers)
