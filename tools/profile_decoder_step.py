"""Kernel-time breakdown of one cfg4 decoder training step (torch profiler, CUDA activities).  Development aid.
    python tools/profile_decoder_step.py [--shared]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import workloads as W
from vision_instance_seg_b200.modules.decoder import build_decoder, set_shared_value_proj

dev = torch.device("cuda:0")
cfg = W.CONFIGS["cfg4_decoder_step_300q"]
torch.manual_seed(0)
dec = build_decoder(256, 8, 9, 2048, 0.0, "relu", 4, 4).to(dev)
if "--shared" in sys.argv:
    set_shared_value_proj(dec)
opt = torch.optim.AdamW(dec.parameters(), lr=1e-5, fused=True)
ss = W.make_spatial_shapes(cfg["shapes"], dev)
lsi = W.make_level_start_index(ss)
S, N, Lq = int(ss.prod(1).sum()), 16, 300
memory = torch.randn(S, N, 256, device=dev)
tgt = torch.randn(Lq, N, 256, device=dev)
refs = torch.randn(Lq, N, 4, device=dev)
vr = torch.ones(N, 4, 2, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    mem = memory.detach().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs, _ = dec(tgt, mem, refpoints_unsigmoid=refs, level_start_index=lsi, spatial_shapes=ss, valid_ratios=vr)
        loss = sum(o.float().square().mean() for o in outs)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes="--shapes" in sys.argv) as prof:
    step()
    torch.cuda.synchronize()
if "--shapes" in sys.argv:
    rows = [e for e in prof.key_averages(group_by_input_shape=True)
            if e.key in ("aten::add_", "aten::copy_", "aten::sum", "aten::mul", "aten::mm", "aten::add", "aten::masked_fill")]
    rows.sort(key=lambda e: -e.self_device_time_total)
    for e in rows[:30]:
        print(f"{e.key:18s} {e.self_device_time_total / 1e3:8.3f} ms  x{e.count:3d}  {str(e.input_shapes)[:150]}")
else:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=80))
