"""Quick GPU development check: parity at small shapes + timing at cfg3.  Not a test, not the bench."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import MSDeformAttnFunction, workloads as W, MultiScaleDeformableAttention as MSDA
from oracle import ms_deform_attn_oracle_grads

dev = "cuda"

def relerr(got, want):
    want = want.double().cpu(); got = got.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()

def parity(name, maker, dtype, flags=0, **kw):
    MSDA.backward_flags = flags
    v, ss, lsi, loc, attn = maker(dtype=dtype, device=dev, **kw)
    v.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
    out = MSDeformAttnFunction.apply(v, ss, lsi, loc, attn, 128)
    go = torch.randn_like(out)
    out.backward(go)
    torch.cuda.synchronize()
    ref = ms_deform_attn_oracle_grads(v.detach().float(), ss.cpu(), loc.detach(), attn.detach(), go.float())
    r = dict(case=name, dtype=str(dtype), flags=flags, out=relerr(out, ref[0]), gv=relerr(v.grad, ref[1]),
             gl=relerr(loc.grad, ref[2]), ga=relerr(attn.grad, ref[3]))
    print(json.dumps(r), flush=True)

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def timing(name, maker, dtype, flags=0, **kw):
    MSDA.backward_flags = flags
    v, ss, lsi, loc, attn = maker(dtype=dtype, device=dev, **kw)
    N, S, M, D = v.shape; Lq = loc.shape[1]; L, P = loc.shape[3], loc.shape[4]
    go = torch.randn(N, Lq, M * D, device=dev, dtype=dtype)
    t_f = timeit(lambda: MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 128))
    t_b = timeit(lambda: MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128))
    ab = W.algorithmic_bytes(N, S, Lq, M, D, L, P, v.element_size())
    print(json.dumps(dict(case=name, dtype=str(dtype), flags=flags, fwd_ms=t_f, bwd_ms=t_b, points=ab["points"],
                          fwd_GBps=ab["fwd"] / t_f / 1e6, bwd_GBps=ab["bwd"] / t_b / 1e6,
                          fwdbwd_Gpts=ab["points"] / (t_f + t_b) / 1e6)), flush=True)

def timing_fused(name, dtype, shapes, batch, ref_dim=2, queries=None):
    from vision_instance_seg_b200 import MSDeformAttnFusedFunction
    ss = W.make_spatial_shapes(shapes, dev); lsi = W.make_level_start_index(ss)
    S = int(ss.prod(1).sum()); L = len(shapes); M, D, P = 8, 32, 4
    Lq = S if queries is None else queries
    g = torch.Generator(device=dev).manual_seed(7)
    v = torch.randn(batch, S, M, D, generator=g, device=dev).to(dtype)
    if ref_dim == 2:
        ref = W.get_reference_points(ss, torch.ones(batch, L, 2, device=dev), dev).contiguous()
    else:
        ref = torch.cat([torch.rand(batch, Lq, 1, 2, generator=g, device=dev).expand(-1, -1, L, -1),
                         torch.rand(batch, Lq, 1, 2, generator=g, device=dev).expand(-1, -1, L, -1) * 0.45 + 0.05], -1).contiguous()
    off = torch.randn(batch, Lq, M, L, P, 2, generator=g, device=dev) * 2.0
    lg = torch.randn(batch, Lq, M, L * P, generator=g, device=dev)
    go = torch.randn(batch, Lq, M * D, device=dev, dtype=dtype)
    t_f = timeit(lambda: MSDA.ms_deform_attn_fused_forward(v, ss, lsi, ref, off, lg, 128))
    t_b = timeit(lambda: MSDA.ms_deform_attn_fused_backward(v, ss, lsi, ref, off, lg, go, 128))
    pts = batch * Lq * M * L * P
    print(json.dumps(dict(case=name, dtype=str(dtype), fused=True, fwd_ms=t_f, bwd_ms=t_b, points=pts,
                          fwdbwd_Gpts=pts / (t_f + t_b) / 1e6)), flush=True)


if __name__ == "__main__":
    if "--fused-only" in sys.argv:
        c3 = W.CONFIGS["cfg3_swinl_1024_bf16"]["shapes"]
        timing("cfg3", W.make_encoder_inputs, torch.bfloat16, shapes=c3, batch=16)
        timing("cfg3_fp32", W.make_encoder_inputs, torch.float32, shapes=c3, batch=16)
        timing("cfg3_uniform", W.make_uniform_inputs, torch.bfloat16, shapes=c3, batch=16)
        timing_fused("cfg3_fused", torch.bfloat16, c3, 16)
        timing_fused("cfg3_fused_fp32", torch.float32, c3, 16)
        timing("cfg4_dec", W.make_decoder_inputs, torch.bfloat16, shapes=c3, batch=16)
        timing_fused("cfg4_dec_fused", torch.bfloat16, c3, 16, ref_dim=4, queries=300)
        sys.exit(0)
    small = [(16, 16), (8, 8), (4, 4)]
    c1 = W.CONFIGS["cfg1_512_fp32"]["shapes"]; c3 = W.CONFIGS["cfg3_swinl_1024_bf16"]["shapes"]
    parity("small_enc", W.make_encoder_inputs, torch.float32, shapes=small, batch=2)
    parity("small_uni", W.make_uniform_inputs, torch.float32, shapes=small, batch=2)
    parity("small_uni64", W.make_uniform_inputs, torch.float64, shapes=small, batch=2)
    parity("cfg1_enc", W.make_encoder_inputs, torch.float32, shapes=c1, batch=2)
    parity("cfg1_enc", W.make_encoder_inputs, torch.bfloat16, shapes=c1, batch=2)
    parity("cfg1_enc", W.make_encoder_inputs, torch.bfloat16, flags=2, shapes=c1, batch=2)
    parity("cfg1_enc", W.make_encoder_inputs, torch.bfloat16, flags=65535 << 8, shapes=c1, batch=2)
    parity("cfg1_enc", W.make_encoder_inputs, torch.float16, shapes=c1, batch=2)
    parity("cfg1_uni_d30", W.make_uniform_inputs, torch.float32, shapes=small, batch=2, head_dim=30)
    parity("cfg1_uni_d64", W.make_uniform_inputs, torch.float32, shapes=small, batch=2, head_dim=64)
    parity("cfg1_dec", W.make_decoder_inputs, torch.bfloat16, shapes=c1, batch=2)
    parity("cfg3_n2_enc", W.make_encoder_inputs, torch.bfloat16, shapes=c3, batch=2)
    parity("cfg3_n2_enc", W.make_encoder_inputs, torch.bfloat16, flags=2, shapes=c3, batch=2)
    parity("cfg3_n2_enc", W.make_encoder_inputs, torch.bfloat16, flags=65535 << 8, shapes=c3, batch=2)
    parity("cfg3_n2_enc", W.make_encoder_inputs, torch.bfloat16, flags=8 << 8, shapes=c3, batch=2)
    timing("cfg1", W.make_encoder_inputs, torch.float32, shapes=c1, batch=2)
    timing("cfg3", W.make_encoder_inputs, torch.bfloat16, shapes=c3, batch=16)
    timing("cfg3", W.make_encoder_inputs, torch.bfloat16, flags=2, shapes=c3, batch=16)
    timing("cfg3", W.make_encoder_inputs, torch.bfloat16, flags=65535 << 8, shapes=c3, batch=16)
    timing("cfg3", W.make_encoder_inputs, torch.bfloat16, flags=8 << 8, shapes=c3, batch=16)
    timing("cfg3_fp32", W.make_encoder_inputs, torch.float32, shapes=c3, batch=16)
    timing("cfg3_uniform", W.make_uniform_inputs, torch.bfloat16, shapes=c3, batch=16)
    timing("cfg4_dec", W.make_decoder_inputs, torch.bfloat16, shapes=c3, batch=16)
    timing("cfg4_dec_500q", W.make_decoder_inputs, torch.bfloat16, shapes=c3, batch=16, queries=500)
    c5 = W.CONFIGS["cfg5_2048_bf16"]["shapes"]
    for nb in (16, 8, 4, 2):
        timing(f"cfg5_n{nb}", W.make_encoder_inputs, torch.bfloat16, shapes=c5, batch=nb)
