"""SURVEY §8d optional row: upstream's pure-PyTorch core (the oracle restatement: F.grid_sample per level) run on the SAME
B200, forward + backward, next to this repo's kernels.  Development aid (imports oracle/: lives under tests/)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from oracle import ms_deform_attn_core_pytorch
from vision_instance_seg_b200 import MSDeformAttnFunction, workloads as W

dev = "cuda"
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for name, dtype, batch in (("cfg3", torch.bfloat16, 16), ("cfg3", torch.float32, 16), ("cfg1", torch.float32, 2)):
    cfg = W.CONFIGS["cfg3_swinl_1024_bf16" if name == "cfg3" else "cfg1_512_fp32"]
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], batch, dtype, device=dev)
    shapes_list = [tuple(s) for s in cfg["shapes"]]
    go = torch.randn(batch, loc.shape[1], 256, device=dev, dtype=dtype)
    v.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
    lo_t, at_t = (loc.to(dtype), attn.to(dtype)) if dtype != torch.float32 else (loc, attn)
    def stock():
        out = ms_deform_attn_core_pytorch(v, shapes_list, lo_t, at_t)
        torch.autograd.grad(out, (v, loc, attn), go)
    def ours():
        out = MSDeformAttnFunction.apply(v, ss, lsi, loc, attn, 128)
        torch.autograd.grad(out, (v, loc, attn), go)
    t_stock, t_ours = timeit(stock), timeit(ours)
    pts = loc.numel() // 2
    print(json.dumps(dict(case=name, dtype=str(dtype), batch=batch, points=pts, stock_torch_gpu_ms=t_stock, b200_kernels_ms=t_ours,
                          stock_Gpts=pts / t_stock / 1e6, ours_Gpts=pts / t_ours / 1e6, speedup=t_stock / t_ours)), flush=True)
