import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200.modules import encoder as E

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

torch.manual_seed(11)
C = int(os.environ.get("C", 128)); heads = C // 32
enc = E.MSDeformAttnTransformerEncoderOnly(d_model=C, nhead=heads, num_encoder_layers=int(os.environ.get("LAYERS", 2)), dim_feedforward=2 * C,
                                           dropout=0.0, num_feature_levels=3, enc_n_points=4).cuda()
with torch.no_grad():
    for layer in enc.encoder.layers:
        layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
        layer.self_attn.attention_weights.weight.normal_(0, 0.2)
shapes = [(12, 20), (6, 10), (3, 5)]
g = torch.Generator().manual_seed(4)
srcs = [torch.randn(2, C, h, w, generator=g).cuda() for h, w in shapes]
pos = [(torch.randn(2, C, h, w, generator=g) * 0.1).cuda() for h, w in shapes]
gout = None
res = {}
for mode in ("fp32", "autocast", "fused"):
    pkg.set_fused_encoder_layers(enc, mode == "fused")
    enc.zero_grad()
    xs = [s.clone().requires_grad_(True) for s in srcs]
    if mode == "autocast":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            mem, _, _ = enc(xs, None, pos)
    else:
        mem, _, _ = enc(xs, None, pos)
    if gout is None:
        gout = torch.randn_like(mem)
    mem.float().backward(gout)
    res[mode] = dict(mem=mem.detach().float(), **{f"src{i}": x.grad for i, x in enumerate(xs)},
                     **{n: p.grad.clone() for n, p in enc.named_parameters()})
for k in res["fp32"]:
    print(f"{k:55s} autocast {rel(res['autocast'][k], res['fp32'][k]):.3e}   fused {rel(res['fused'][k], res['fp32'][k]):.3e}")
