"""Kernel-time breakdown of one encoder training step (torch profiler, CUDA activities).  Development aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import workloads as W, distributed as D
from vision_instance_seg_b200.modules.encoder import MSDeformAttnTransformerEncoderOnly

fused = "--fused" in sys.argv
dev = torch.device("cuda:0")
cfg = W.CONFIGS["cfg3_train_step_1024"]
torch.manual_seed(0)
enc = MSDeformAttnTransformerEncoderOnly(256, 8, 6, 2048, 0.0, "relu", 4, 4).to(dev)
pkg.set_fused_preop(enc, fused)
layers = "--fused-layers" in sys.argv
pkg.set_fused_encoder_layers(enc, layers)
buckets = D.GradientBuckets(D.encoder_gradient_groups(enc), device=dev)
opt = torch.optim.AdamW(enc.parameters(), lr=1e-5, fused=True)
srcs, pos = W.make_feature_pyramid(cfg["shapes"], 16, 256, device=dev)

def step():
    buckets.zero()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not layers):
        mem, _, _ = enc(srcs, None, pos)
    mem.float().square().mean().backward()
    buckets.wait()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
