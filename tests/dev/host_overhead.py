"""Host-side cost per call of the Python boundary at the small BASELINE shapes (cfg1: 512^2 fp32; cfg4: 300 queries
bf16), where the kernels take tens of microseconds and the host is the limiter.  Enqueue rate (no sync inside the loop)
and device time (events) per call."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, MSDeformAttnFunction, workloads as W

dev = "cuda"


def bench(fn, n=500):
    for _ in range(30):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return round((t1 - t0) / n * 1e6, 1), round(e0.elapsed_time(e1) / n * 1e3, 1)


for name in ("cfg1_512_fp32", "cfg4_decoder_300q_bf16"):
    cfg = W.CONFIGS[name]
    if cfg["kind"] == "decoder":
        v, ss, lsi, loc, attn = W.make_decoder_inputs(cfg["shapes"], cfg["batch"], cfg["dtype"], queries=cfg["queries"], device=dev)
    else:
        v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], cfg["batch"], cfg["dtype"], device=dev)
    go = torch.randn(v.shape[0], loc.shape[1], 256, device=dev).to(cfg["dtype"])
    row = {"workload": name}
    row["forward_wrapper_host_us,device_us"] = bench(lambda: MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 128))
    row["backward_wrapper_host_us,device_us"] = bench(lambda: MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128))
    vv, ll, aa = v.clone().requires_grad_(True), loc.clone().requires_grad_(True), attn.clone().requires_grad_(True)

    def fb():
        out = MSDeformAttnFunction.apply(vv, ss, lsi, ll, aa, 128)
        torch.autograd.grad(out, (vv, ll, aa), go)
    row["autograd_fwd_bwd_host_us,device_us"] = bench(fb, 300)
    print(json.dumps(row), flush=True)
print(json.dumps({"torch.empty_like_host_us": bench(lambda: torch.empty_like(v))[0]}))
