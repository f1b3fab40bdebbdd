"""cfg4 decoder cross-attention stack on a GPU box: 9 x MSDeformAttn(query 300, memory 21 760 tokens, batch 16) forward +
backward under bf16 autocast, (a) as upstream runs it (nine value_proj GEMMs), (b) with share_value_proj (one stacked
GEMM, strided kernels), each with the sparse-level direct accumulation on and off.  Prints one JSON line per variant.

    python tests/dev/decoder_cross_attn.py [--batch 16] [--queries 300] [--iters 20]
"""
import argparse
import copy
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--queries", type=int, default=300)
    ap.add_argument("--layers", type=int, default=9)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import vision_instance_seg_b200 as pkg
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, _lib, workloads as W
    from vision_instance_seg_b200.modules import MSDeformAttn, share_value_proj
    dev = torch.device("cuda:0")
    cfg = W.CONFIGS["cfg4_decoder_300q_bf16"]
    ss = W.make_spatial_shapes(cfg["shapes"], dev)
    lsi = W.make_level_start_index(ss)
    S = int(ss.prod(1).sum())
    N, Lq, K = args.batch, args.queries, args.layers
    g = torch.Generator(device=dev).manual_seed(1234)
    src = torch.randn(N, S, 256, generator=g, device=dev)
    queries = [torch.randn(N, Lq, 256, generator=g, device=dev) for _ in range(K)]
    ctr = torch.rand(N, Lq, 1, 2, generator=g, device=dev).expand(-1, -1, 4, -1)
    wh = torch.rand(N, Lq, 1, 2, generator=g, device=dev).expand(-1, -1, 4, -1) * 0.45 + 0.05
    ref = torch.cat([ctr, wh], -1).contiguous()
    torch.manual_seed(0)
    mods = [MSDeformAttn(256, 4, 8, 4).to(dev) for _ in range(K)]
    for m in mods:
        torch.nn.init.normal_(m.sampling_offsets.weight, std=0.02)
        torch.nn.init.normal_(m.attention_weights.weight, std=0.05)
    shared = copy.deepcopy(mods)
    share_value_proj(shared)

    def step(modules):
        s = src.requires_grad_(True)
        s.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = 0
            for m, q in zip(modules, queries):
                loss = loss + m(q, ref, s, ss, lsi, None).float().square().mean()
        loss.backward()

    def timed(modules, flags):
        old = MSDA.backward_flags
        MSDA.backward_flags = flags
        try:
            for _ in range(args.warmup):
                step(modules)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                step(modules)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.iters
        finally:
            MSDA.backward_flags = old

    for name, modules in (("independent value_proj", mods), ("share_value_proj", shared)):
        for sparse in (False, True):
            flags = 0 if sparse else _lib.MSDA_BWD_NO_SPARSE_DIRECT
            ms = timed(modules, flags)
            print(json.dumps({"what": f"cfg4 decoder cross-attention x{K}, fwd+bwd, bf16 autocast", "variant": name,
                              "sparse_direct": sparse, "batch": N, "queries": Lq, "ms_per_step": round(ms, 3),
                              "ms_per_layer": round(ms / K, 4), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 2)}),
                  flush=True)


if __name__ == "__main__":
    main()
