"""Host-side cost per call of the Python boundary (tiny problem, GPU idle most of the time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, MSDeformAttnFunction, workloads as W

dev = "cuda"
v, ss, lsi, loc, attn = W.make_decoder_inputs([(16, 16), (8, 8)], 1, torch.bfloat16, queries=32, device=dev)
go = torch.randn(1, 32, 256, device=dev, dtype=torch.bfloat16)
def bench(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6
print("forward  wrapper us/call:", round(bench(lambda: MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 128)), 1))
print("backward wrapper us/call:", round(bench(lambda: MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128)), 1))
vv = v.clone().requires_grad_(True)
def fb():
    out = MSDeformAttnFunction.apply(vv, ss, lsi, loc, attn, 128)
    torch.autograd.grad(out, vv, go)
print("autograd fwd+bwd us/call:", round(bench(fb, 1000), 1))
print("torch.empty_like us/call:", round(bench(lambda: torch.empty_like(v)), 1))
