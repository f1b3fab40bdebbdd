"""Tiny end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import MSDeformAttnFunction, MSDeformAttnFusedFunction, workloads as W
from vision_instance_seg_b200.modules.encoder import MSDeformAttnTransformerEncoderOnly

dev = "cuda:0"
shapes = [(9, 7), (5, 4), (3, 2)]
for dtype in (torch.float32, torch.bfloat16, torch.float16, torch.float64):
    for D in ((32, 16, 30) if dtype != torch.float64 else (8,)):
        v, ss, lsi, loc, attn = W.make_uniform_inputs(shapes, 2, dtype, queries=13, n_heads=3, head_dim=D, device=dev)
        if dtype == torch.float64:
            loc, attn = loc.double(), attn.double()
        v.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
        out = MSDeformAttnFunction.apply(v, ss, lsi, loc, attn, 64)
        out.backward(torch.randn_like(out))
for dtype, aux in ((torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)):
    for R in (2, 4):
        N, Lq, M, D, L, P = 2, 11, 3, 32, 3, 4
        ss = W.make_spatial_shapes(shapes, dev); lsi = W.make_level_start_index(ss); S = int(ss.prod(1).sum())
        v = torch.randn(N, S, M, D, device=dev).to(dtype).requires_grad_(True)
        ref = torch.rand(N, Lq, L, R, device=dev)
        off = torch.randn(N, Lq, M, L, P, 2, device=dev).to(aux).requires_grad_(True)
        lg = torch.randn(N, Lq, M, L * P, device=dev).to(aux).requires_grad_(True)
        out = MSDeformAttnFusedFunction.apply(v, ss, lsi, ref, off, lg, 64)
        out.backward(torch.randn_like(out))
enc = MSDeformAttnTransformerEncoderOnly(128, 4, 2, 256, 0.0, "relu", 3, 4).to(dev)
pkg.set_fused_encoder_layers(enc, True)
srcs = [torch.randn(2, 128, h, w, device=dev) for h, w in shapes]
mem, _, _ = enc(srcs, None, [torch.zeros_like(s) for s in srcs])
mem.square().mean().backward()
torch.cuda.synchronize()
print("sanitize_small ok")
