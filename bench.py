#!/usr/bin/env python
"""bench.py — MSDeformAttn fwd+bwd sampled points/s and % of HBM roofline (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over one batch of synthetic input: the Swin-L MaskDINO pixel-decoder
encoder shape (BASELINE.json configs[2]: 1024x1024 -> 4 levels 128/64/32/16, 21 760 queries, batch 16,
value bf16 + sampling locations / attention weights fp32, 6 layers), i.e. 6 x (forward + backward) of
MSDeformAttnFunction.  N > 1 (torchrun): every rank runs the same per-GPU batch (weak scaling, no data-path
collective) and all-reduces one encoder-sized gradient bucket per step over NCCL on a side stream.

`value`  : device-resident inputs, public autograd API, CUDA events, max over ranks.
`e2e`    : same API, inputs start in pinned HOST memory and results end there, copies inside the timed region.
`roofline`: dominant kernel (backward), algorithmic bytes / event-timed launch duration / measured HBM peak.
`cpu_baseline` and `--impl reference`: the CPU oracle (restated ms_deform_attn_core_pytorch) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "msdeformattn_fwd_bwd_sampled_points_per_sec"
UNIT = "points/s"
WORKLOAD = "cfg3_swinl_1024_bf16"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=None, help="default: --steps")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=1)
    ap.add_argument("--fused-preop", action="store_true",
                    help="train-step workloads: fold softmax + sampling-location arithmetic into the kernels")
    ap.add_argument("--forward-only", action="store_true",
                    help="train-step workloads: time the encoder forward alone under no_grad (inference, reference call stack "
                         "evaluate.py / visualize.py); the metric counts forward sampled points only")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="train-step workloads: capture forward + backward + all-reduce + optimizer in one CUDA graph")
    ap.add_argument("--fused-layers", action="store_true",
                    help="train-step workloads: run every encoder layer as one fused autograd node (implies --fused-preop)")
    ap.add_argument("--shared-value-proj", action="store_true",
                    help="decoder-step workload: one stacked value_proj GEMM for all decoder layers (share_value_proj)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="operator workloads: skip config.secondary (cold / warm L2, fp32, uniform locations, tiled kernels)")
    ap.add_argument("--no-train-step", action="store_true",
                    help="default workload: skip the embedded cfg5 2048^2 training step (config.train_step_cfg5)")
    ap.add_argument("--train-step-batch", type=int, default=16, help="global batch of the embedded cfg5 training step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes(kernel_key):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


def measured_red_ceiling_grows():
    """L2 reduction ceiling for 64-byte packed-fp16 rows (G rows/s): the best `redg_bf16x8_64Brow` / f16x8 figure of the
    committed microbenchmark record, or None."""
    best = None
    try:
        with open(os.path.join(ROOT, "profiles", "microbench_r01.jsonl")) as f:
            for line in f:
                line = line.strip()
                if not line.startswith("{"):
                    continue
                d = json.loads(line)
                if "64Brow" in str(d.get("test", "")) and str(d.get("test", "")).startswith("redg") and "Grows_per_s" in d:
                    best = max(best or 0.0, float(d["Grows_per_s"]))
    except Exception:
        return None
    return best


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:   # region shorter than one sample: fall back to everything sampled
            sm = [float(l.split(",")[0]) for _, l in self.lines if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_step_time(cfg, batch, steps, warmup, min_seconds=0.0):
    """Times the CPU oracle (torch restatement of ms_deform_attn_core_pytorch, fp32, all host threads) on one
    layer's forward+backward of `batch` images of the workload.  Returns (seconds per step, points per step)."""
    import torch
    from oracle import ms_deform_attn_core_pytorch
    from vision_instance_seg_b200 import workloads as W
    torch.set_num_threads(os.cpu_count() or 1)
    v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], batch, torch.float32, device="cpu", seed=4321)
    go = torch.randn(batch, loc.shape[1], v.shape[2] * v.shape[3])
    v.requires_grad_(True), loc.requires_grad_(True), attn.requires_grad_(True)
    pts = loc.numel() // 2

    def step():
        out = ms_deform_attn_core_pytorch(v, ss, loc, attn)
        torch.autograd.grad(out, (v, loc, attn), go)

    for _ in range(warmup):
        step()
    times = []
    while len(times) < steps or sum(times) < min_seconds:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times, pts


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port), host cores only
# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from vision_instance_seg_b200 import workloads as W
    cfg = W.CONFIGS[args.workload]
    batch = args.cpu_sample_batch
    times, pts = cpu_oracle_step_time(cfg, batch, args.steps, args.warmup)
    total = sum(times)
    value = pts * len(times) / total
    cores = os.cpu_count() or 1
    sample = (f"{args.workload}: {batch} image(s) x 1 layer forward+backward per step, fp32, "
              f"ms_deform_attn_core_pytorch restatement (oracle/), torch CPU {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample, "l2_policy": "cpu run"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def binding_resource(ab, bwd_avg_ms, dtype, tiled):
    """What binds the direct backward: 4 corner-row reductions per sampled point against the measured
    L2 reduction ceiling for 64-byte packed-fp16 rows (profiles/microbench_r01.jsonl)."""
    import torch
    if dtype == torch.float32 or tiled:
        return None
    ceiling = measured_red_ceiling_grows()
    rows = 4 * ab["points"] / (bwd_avg_ms * 1e-3) / 1e9
    return {"name": "L2 reduction path (red.global.add.noftz.v4.f16x2, 64-byte rows)", "achieved_Grows_per_s": rows,
            "measured_ceiling_Grows_per_s": ceiling, "ceiling_source": "profiles/microbench_r01.jsonl (best redg_*_64Brow)",
            "frac": rows / ceiling if ceiling else None}


def e2e_fused_16bit(args, cfg, batch, layers, dev, world, bucket, barrier, pts_per_step):
    """`e2e` once more through HostPipeline(fused=True): operands (value, reference_points, raw sampling offsets and
    attention logits in the value dtype, grad_output) start in pinned host memory, results (output, grad_value, grad of
    offsets / logits in the value dtype) end there."""
    import torch
    from vision_instance_seg_b200 import distributed as D, workloads as W
    from vision_instance_seg_b200.host_pipeline import HostPipeline
    dtype = cfg["dtype"]
    shapes = cfg["shapes"]
    ss = W.make_spatial_shapes(shapes, dev)
    lsi = W.make_level_start_index(ss)
    L, S = len(shapes), int(ss.prod(1).sum())
    M, Dh, P = 8, 32, 4
    g = torch.Generator().manual_seed(4242)
    value = torch.randn(batch, S, M, Dh, generator=g).to(dtype).pin_memory()
    ref = W.get_reference_points(ss.cpu(), torch.ones(batch, L, 2)).contiguous().pin_memory()         # (N, S, L, 2) fp32
    off = (torch.randn(batch, S, M, L, P, 2, generator=g) * 2.0).to(dtype).pin_memory()               # pixels, raw
    logits = torch.randn(batch, S, M, L * P, generator=g).to(dtype).pin_memory()
    go = torch.randn(batch, S, M * Dh, generator=g).to(dtype).pin_memory()
    host_in = [value, ref, off, logits, go]
    host_out = [torch.empty((batch, S, M * Dh), dtype=dtype).pin_memory(), torch.empty(tuple(value.shape), dtype=dtype).pin_memory(),
                torch.empty(tuple(off.shape), dtype=dtype).pin_memory(), torch.empty(tuple(logits.shape), dtype=dtype).pin_memory()]
    h2d = sum(t.numel() * t.element_size() for t in host_in) * layers
    d2h = sum(t.numel() * t.element_size() for t in host_out) * layers
    pipe = HostPipeline(ss, lsi, dev, fused=True)

    def step():
        for _ in range(layers):
            pipe.submit(host_in, host_out)
        if bucket is not None:
            bucket.allreduce_async()

    step()
    pipe.wait_all()
    if bucket is not None:
        bucket.wait()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        step()
    pipe.wait_all()
    if bucket is not None:
        bucket.wait()
    e1.record()
    barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), dev) / args.e2e_steps
    assert torch.isfinite(host_out[0].float()).all()
    return {"value": pts_per_step * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "pcie_gbs_per_rank": {"h2d": h2d / (ms * 1e-3) / 1e9, "d2h": d2h / (ms * 1e-3) / 1e9},
            "api": "HostPipeline(fused=True).submit: msda_fused_forward / msda_fused_backward, offsets / logits and their "
                   "gradients in the value dtype"}


def _pair_guarded(bwd_ms, n_calls):
    """With the cluster guard (decoder-sized calls) a backward enqueues its kernel twice -- the fp16 and the fp32
    accumulation pipeline, one of which returns at once -- and both are recorded as the backward kernel: one call's kernel
    time is the sum of the pair, not the mean of a ~60 us and a ~3 us launch."""
    if n_calls and len(bwd_ms) == 2 * n_calls:
        return [a + b for a, b in zip(bwd_ms[0::2], bwd_ms[1::2])]
    return bwd_ms


def secondary_rows(args, cfg, batch, dev, lib):
    """One layer's forward and backward call of the extension-level API, CUDA events around each call (so the backward
    includes its zero / max / rounding passes), median of 7 after 3 warm-ups:
      cold  : 256 MiB written between calls (L2 flushed);  warm: the same inputs back to back (value of one image batch
              does not fit the 126 MB L2 at this size, so `warm` mostly keeps loc / attn / the accumulator tail);
      fp32  : the same geometry with float32 values;       uniform: sampling locations U(-0.05, 1.05) (no locality);
      tiled : the opt-in tiled kernels (msda_set_tiled_mode(1)) on the encoder-like inputs;
      hybrid: mode 2 (direct forward; backward = direct kernel for the fine levels + sorting kernel for the two coarsest)."""
    import torch
    from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, workloads as W
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, cold):
        ts = []
        for r in range(10):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            if r >= 3:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    def row(dtype, kind, tiled=0, cold=True):
        maker = W.make_uniform_inputs if kind == "uniform" else W.make_encoder_inputs
        v, ss, lsi, loc, attn = maker(cfg["shapes"], batch, dtype, seed=777, device=dev)
        go = torch.randn(batch, loc.shape[1], v.shape[2] * v.shape[3], device=dev).to(dtype)
        prev = lib.msda_set_tiled_mode(int(tiled))
        prev_split = lib.msda_set_hybrid_split(30) if int(tiled) == 2 else None
        try:
            f = timed(lambda: MSDA.ms_deform_attn_forward(v, ss, lsi, loc, attn, 128), cold)
            b = timed(lambda: MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128), cold)
        finally:
            lib.msda_set_tiled_mode(prev)
            if prev_split is not None:
                lib.msda_set_hybrid_split(prev_split)
        pts = loc.numel() // 2
        return {"forward_ms": f, "backward_ms": b, "points_per_s": pts / ((f + b) * 1e-3)}

    dtype = cfg["dtype"]
    out = {"what": "one layer, extension-level forward / backward call, CUDA events per call, median of 7",
           "cold_l2": row(dtype, "encoder"), "warm_l2": row(dtype, "encoder", cold=False),
           "uniform_locations": row(dtype, "uniform")}
    if dtype != torch.float32:
        out["fp32_values"] = row(torch.float32, "encoder")
        out["tiled_kernels"] = row(dtype, "encoder", tiled=1)
        out["hybrid_backward"] = row(dtype, "encoder", tiled=2)
    del flush
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import vision_instance_seg_b200 as pkg
    from vision_instance_seg_b200 import MSDeformAttnFunction, _lib, distributed as D, workloads as W

    rank, local_rank, world = D.init_process_group()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = D.bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: unbound"
    lib = pkg.load_library()          # raises if the CUDA library is missing: no fallback

    cfg = W.CONFIGS[args.workload]
    batch = args.batch or cfg["batch"]
    layers = args.layers or cfg.get("layers", 1)
    dtype = cfg["dtype"]
    maker = W.make_decoder_inputs if cfg["kind"] == "decoder" else W.make_encoder_inputs
    extra = {"queries": cfg["queries"]} if cfg["kind"] == "decoder" else {}

    # one independent input set per layer: 6 x 0.89 GB at cfg3 — far larger than the 126 MB L2, so no
    # iteration finds its inputs cached
    sets = []
    for layer in range(layers):
        v, ss, lsi, loc, attn = maker(cfg["shapes"], batch, dtype, seed=1234 + rank + 100 * layer, device=dev, **extra)
        go = torch.randn(batch, loc.shape[1], v.shape[2] * v.shape[3], device=dev, dtype=torch.float32).to(dtype)
        sets.append([v.requires_grad_(True), ss, lsi, loc.requires_grad_(True), attn.requires_grad_(True), go])
    N, S, M, Dh = sets[0][0].shape
    Lq, L, P = sets[0][3].shape[1], sets[0][3].shape[3], sets[0][3].shape[4]
    ab = W.algorithmic_bytes(N, S, Lq, M, Dh, L, P, sets[0][0].element_size())
    pts_per_step = ab["points"] * layers

    bucket = D.GradientBucket(device=dev) if world > 1 else None

    def step():
        for v, ss, lsi, loc, attn, go in sets:
            out = MSDeformAttnFunction.apply(v, ss, lsi, loc, attn, 128)
            torch.autograd.grad(out, (v, loc, attn), go)
        if bucket is not None:
            bucket.allreduce_async()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    if bucket is not None:
        bucket.wait()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    lib.msda_profile_enable(1)
    launches0 = lib.msda_total_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    if bucket is not None:
        bucket.wait()
    e1.record()
    barrier()
    t_wall1 = time.time()
    lib.msda_profile_enable(0)
    launches = lib.msda_total_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_total = D.max_over_ranks(e0.elapsed_time(e1), dev)
    ms_per_step = ms_total / args.steps
    value = pts_per_step * world / (ms_per_step * 1e-3)
    records = _lib.profile_collect()
    fwd_ms = [t for t, k in records if k == 1]
    bwd_ms = _pair_guarded([t for t, k in records if k == 2], len(fwd_ms))
    dots_ms = [t for t, k in records if k == 3]          # tiled mode (MSDA_B200_TILED=1): the backward is two kernels
    scat_ms = [t for t, k in records if k == 4]
    if not bwd_ms and dots_ms and len(dots_ms) == len(scat_ms):
        bwd_ms = [a + b for a, b in zip(dots_ms, scat_ms)]

    # ---- secondary measurements of the same operator: one layer's forward / backward call, event-timed ----
    secondary = None
    if not args.no_secondary and cfg["kind"] == "encoder":
        secondary = secondary_rows(args, cfg, batch, dev, lib)

    # ---- end to end: pinned host -> device -> fwd+bwd -> pinned host, every layer of every step ----
    e2e = None
    if not args.no_e2e:
        from vision_instance_seg_b200.host_pipeline import HostPipeline
        v, ss, lsi, loc, attn, go = sets[0]
        host_in = [t.detach().to("cpu").pin_memory() for t in (v, loc, attn, go)]
        host_out = [torch.empty((N, Lq, M * Dh), dtype=dtype).pin_memory(), torch.empty(tuple(v.shape), dtype=dtype).pin_memory(),
                    torch.empty(tuple(loc.shape), dtype=torch.float32).pin_memory(),
                    torch.empty(tuple(attn.shape), dtype=torch.float32).pin_memory()]
        h2d = sum(t.numel() * t.element_size() for t in host_in) * layers
        d2h = sum(t.numel() * t.element_size() for t in host_out) * layers
        pipe = HostPipeline(ss, lsi, dev)

        def e2e_step():
            for _ in range(layers):
                pipe.submit(host_in, host_out)      # H2D of all four operands, fwd, bwd, D2H of all four results
            if bucket is not None:
                bucket.allreduce_async()

        e2e_step()
        pipe.wait_all()
        if bucket is not None:
            bucket.wait()
        barrier()
        e0.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        pipe.wait_all()
        if bucket is not None:
            bucket.wait()
        e1.record()
        barrier()
        e2e_ms = D.max_over_ranks(e0.elapsed_time(e1), dev) / args.e2e_steps
        # spot-check that the pipelined path produced the same forward result as the device-resident path
        ref_out = MSDeformAttnFunction.apply(v.detach(), ss, lsi, loc.detach(), attn.detach(), 128)
        assert torch.equal(ref_out.cpu(), host_out[0]), "e2e pipeline result differs from the device-resident path"
        e2e = {"value": pts_per_step * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": args.e2e_steps,
               "pcie_gbs_per_rank": {"h2d": h2d / (e2e_ms * 1e-3) / 1e9, "d2h": d2h / (e2e_ms * 1e-3) / 1e9},
               "api": "HostPipeline.submit (pinned host operands -> ms_deform_attn_forward/backward -> pinned host results; "
                      "3 streams, batch cut into 4 pieces, triple-buffered staging)"}
        del host_in, host_out, pipe
        # second figure: the fused pre-op entry points with 16-bit offsets / logits (what the Linears emit under
        # autocast): 6 instead of 12 bytes of auxiliary operands per sampled point in each direction
        if dtype != torch.float32 and cfg["kind"] == "encoder":
            try:
                e2e["fused_16bit_aux"] = e2e_fused_16bit(args, cfg, batch, layers, dev, world, bucket, barrier, pts_per_step)
            except Exception as exc:
                e2e["fused_16bit_aux"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    # ---- BASELINE.json configs[4]: the 2048^2 training step, batch-sharded (strong scaling), on the same ranks ----
    train_step = None
    if args.workload == WORKLOAD and not args.no_train_step:
        del sets
        torch.cuda.empty_cache()
        sub = argparse.Namespace(**vars(args))
        sub.workload, sub.batch, sub.layers = "cfg5_train_step_2048", args.train_step_batch, None
        sub.fused_layers, sub.fused_preop, sub.cuda_graph, sub.forward_only, sub.shared_value_proj = True, False, True, False, False
        sub.no_e2e, sub.steps, sub.warmup = True, 3, 3
        try:
            tl = train_step_line(sub)
        except Exception as exc:       # e.g. out of memory on a smaller GPU: report, do not lose the operator line
            tl = {"error": f"{type(exc).__name__}: {exc}"[:300]} if rank == 0 else None
        if tl is not None and "error" not in tl:
            c = tl["config"]
            train_step = {"workload": c["workload"], "global_batch": c["global_batch"], "per_gpu_batch": c["per_gpu_batch"],
                          "n_gpus": tl["n_gpus"], "scaling": tl["scaling"], "ms_per_step": tl["ms_per_step"],
                          "points_per_s": tl["value"], "points_per_step": c["points_per_step"],
                          "grad_allreduce_bytes": c["grad_allreduce_bytes"], "grad_buckets": c["grad_buckets"],
                          "fused_layers": c["fused_layers"], "cuda_graph": c["cuda_graph"], "gpu_launches_per_step": tl["gpu_launches"] // max(tl["steps"], 1),
                          "msda_share_of_step": c.get("msda_share_of_step"), "msda_ms_per_step": c.get("msda_ms_per_step"),
                          "peak_device_memory_gb": c["peak_device_memory_gb"], "final_loss": c["final_loss"],
                          "note": "strong scaling: efficiency(N) = ms_per_step(1) / (N * ms_per_step(N)); the driver computes it"}
        else:
            train_step = tl

    if rank != 0:
        return 0

    peak, peak_src = measured_peak_gbs()
    bwd_avg = statistics.mean(bwd_ms) if bwd_ms else None
    fwd_avg = statistics.mean(fwd_ms) if fwd_ms else None
    roofline = None
    if bwd_avg:
        achieved = ab["bwd"] / (bwd_avg * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "msda_bwd (backward gather + grad_value scatter)", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes("backward"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": ab["bwd"], "avg_launch_ms": bwd_avg,
                    "launches_timed": len(bwd_ms),
                    "forward": {"achieved": ab["fwd"] / (fwd_avg * 1e-3) / 1e9 if fwd_avg else None,
                                "frac": ab["fwd"] / (fwd_avg * 1e-3) / 1e9 / peak if fwd_avg else None,
                                "algorithmic_bytes_per_launch": ab["fwd"], "avg_launch_ms": fwd_avg,
                                "traffic": ncu_traffic_bytes("forward")},
                    "step_frac": (ab["fwd"] + ab["bwd"]) * layers / (ms_per_step * 1e-3) / 1e9 / peak,
                    # sampled points/s of the forward and the backward kernel on their own
                    "forward_points_per_s": ab["points"] / (fwd_avg * 1e-3) if fwd_avg else None,
                    "backward_points_per_s": ab["points"] / (bwd_avg * 1e-3),
                    # what actually binds the backward: 4 corner-row reductions per sampled point
                    # against the measured L2 reduction ceiling for 64-byte packed-fp16 rows (profiles/microbench_r01.jsonl)
                    "binding_resource": binding_resource(ab, bwd_avg, dtype, bool(dots_ms))}
        if dots_ms:
            roofline["kernel"] = "msda_bwd_dots_tiled + msda_bwd_scatter_tiled (MSDA_B200_TILED=1)"
            roofline["tiled_kernels_ms"] = {"dots": statistics.mean(dots_ms), "scatter": statistics.mean(scat_ms)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        times, pts = cpu_oracle_step_time(cfg, args.cpu_sample_batch, 3, 1, min_seconds=12.0)
        cores = os.cpu_count() or 1
        cpu_baseline = {"value": pts * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{args.workload}: {args.cpu_sample_batch} image(s) x 1 layer fwd+bwd per run, fp32, "
                                  f"{len(times)} timed runs after 1 warm-up ({sum(times):.1f} s of CPU work), oracle "
                                  f"ms_deform_attn_core_pytorch restatement on torch CPU with {cores} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {torch.bfloat16: "bf16", torch.float32: "f32", torch.float16: "f16"}[dtype], "data": "synthetic",
        "config": {"workload": args.workload, "per_gpu_batch": batch, "global_batch": batch * world, "layers": layers,
                   "levels": cfg["shapes"], "queries": Lq, "heads": M, "head_dim": Dh, "points": P,
                   "aux_dtype": "f32", "points_per_step_per_gpu": pts_per_step, "parallelism": f"dp{world}",
                   "grad_allreduce_bytes": D.ENCODER_GRAD_ELEMENTS * 4 if world > 1 else 0, "cpu_binding_rank0": numa,
                   "l2_policy": f"{layers} independent input sets ({layers * (ab['fwd'] + ab['bwd']) / 1e9:.1f} GB touched per step) >> 126 MB L2; no flush needed",
                   "tiled_kernels": os.environ.get("MSDA_B200_TILED") == "1", "secondary": secondary,
                   "train_step_cfg5": train_step},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
# training-step workloads (BASELINE.json configs[4]): 6-layer encoder, batch-sharded, gradient all-reduce, AdamW
# ------------------------------------------------------------------------------------------------------
def run_train_step(args):
    line = train_step_line(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def train_step_line(args):
    """Runs the training-step workload on every rank; returns the JSON line (a dict) on rank 0, None elsewhere."""
    import torch
    import vision_instance_seg_b200 as pkg
    from vision_instance_seg_b200 import _lib, distributed as D, workloads as W
    from vision_instance_seg_b200.modules.encoder import MSDeformAttnTransformerEncoderOnly

    rank, local_rank, world = D.init_process_group()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = pkg.load_library()
    cfg = W.CONFIGS[args.workload]
    global_batch = args.batch or cfg["batch"]
    start, count = D.shard_batch(global_batch, world, rank)
    if count == 0:
        raise SystemExit("global batch smaller than the number of ranks")
    layers = args.layers or cfg["layers"]
    shapes, M, P, C = cfg["shapes"], cfg["heads"], cfg["points"], cfg["d_model"]
    L = len(shapes)
    S = sum(h * w for h, w in shapes)

    torch.manual_seed(1234)                     # identical parameters on every rank
    decoder_step = cfg["kind"] == "decoder_step"
    if decoder_step:
        # BASELINE configs[3]: the deformable decoder (9 layers, `queries` box queries) on a fixed encoder memory; the
        # memory is a differentiable input (its gradient is what the encoder would receive), every layer's output is
        # supervised (DETR-style auxiliary losses)
        from vision_instance_seg_b200.modules.decoder import build_decoder, set_shared_value_proj
        Lq = cfg["queries"]
        enc = build_decoder(d_model=C, nhead=M, num_decoder_layers=layers, dim_feedforward=cfg["d_ffn"], dropout=0.0,
                            num_feature_levels=L, dec_n_points=P).to(dev)
        with torch.no_grad():
            for layer in enc.layers:
                layer.cross_attn.sampling_offsets.weight.normal_(0, 0.01)
                layer.cross_attn.attention_weights.weight.normal_(0, 0.05)
        if args.shared_value_proj:
            set_shared_value_proj(enc)
        pkg.set_fused_preop(enc, args.fused_preop)
        groups = [list(layer.parameters()) for layer in reversed(list(enc.layers))]
        groups.append(list(enc.ref_point_head.parameters()) + list(enc.norm.parameters()))
        buckets = D.GradientBuckets(groups, device=dev)
        gen = torch.Generator().manual_seed(99 + start)
        host_srcs = [torch.randn(S, count, C, generator=gen).pin_memory()]           # encoder memory, sequence first
        tgt = torch.randn(Lq, count, C, generator=gen).to(dev)
        refs = torch.randn(Lq, count, 4, generator=gen).to(dev)
        ss_dev = W.make_spatial_shapes(shapes, dev)
        lsi_dev = W.make_level_start_index(ss_dev)
        vr = torch.ones(count, L, 2, device=dev)
        pos = None
        pts_per_step = global_batch * Lq * M * L * P * layers
    else:
        enc = MSDeformAttnTransformerEncoderOnly(d_model=C, nhead=M, num_encoder_layers=layers, dim_feedforward=cfg["d_ffn"],
                                                 dropout=0.0, num_feature_levels=L, enc_n_points=P).to(dev)
        with torch.no_grad():                       # trained-like projections: offsets / weights depend on the query
            for layer in enc.encoder.layers:
                layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
                layer.self_attn.attention_weights.weight.normal_(0, 0.05)
        pkg.set_fused_preop(enc, args.fused_preop)
        pkg.set_fused_encoder_layers(enc, args.fused_layers)
        buckets = D.GradientBuckets(D.encoder_gradient_groups(enc), device=dev)
        host_srcs, host_pos = W.make_feature_pyramid(shapes, count, C, seed=99 + start, device="cpu", pin=True)
        pos = [t.to(dev) for t in host_pos]
        pts_per_step = global_batch * S * M * L * P * layers
    opt = torch.optim.AdamW(enc.parameters(), lr=1e-5, fused=True, capturable=args.cuda_graph)
    srcs = [t.to(dev) for t in host_srcs]
    loss_host = torch.zeros(1).pin_memory()

    def decoder_body(x):
        memory = x[0].detach().requires_grad_(not args.forward_only)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs, _ = enc(tgt, memory, refpoints_unsigmoid=refs, level_start_index=lsi_dev, spatial_shapes=ss_dev,
                          valid_ratios=vr)
            loss = sum(o.float().square().mean() for o in outs)
        return loss

    def step_body(x):
        if decoder_step:
            if args.forward_only:
                with torch.no_grad():
                    return decoder_body(x)
            buckets.zero()
            loss = decoder_body(x)
            loss.backward()
            buckets.wait()
            opt.step()
            return loss.detach()
        if args.forward_only:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.fused_layers):
                memory, _, _ = enc(x, None, pos)
            return memory.float().square().mean()
        buckets.zero()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.fused_layers):
            memory, _, _ = enc(x, None, pos)       # the fused layers are bf16 by construction
        loss = memory.float().square().mean()
        loss.backward()
        buckets.wait()
        opt.step()
        return loss.detach()

    graph, graph_loss = None, None

    def step(from_host=False):
        if graph is not None:                   # static inputs: refresh them in place, then replay the captured step
            if from_host:
                for d, h in zip(srcs, host_srcs):
                    d.copy_(h, non_blocking=True)
            graph.replay()
            loss = graph_loss
        else:
            loss = step_body([t.to(dev, non_blocking=True) for t in host_srcs] if from_host else srcs)
        if from_host:
            loss_host.copy_(loss.reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    if args.cuda_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):           # warm up off the default stream (allocator, cuBLAS workspaces, NCCL)
            for i in range(3):
                # kernels inside a replayed graph cannot be event-timed: time the sampling kernels of the last eager
                # warm-up step of the same body instead (config.msda_ms_per_step)
                lib.msda_profile_enable(1 if i == 2 else 0)
                step_body(srcs)
            lib.msda_profile_enable(0)
        torch.cuda.current_stream(dev).wait_stream(side)
        barrier()
        eager_records = _lib.profile_collect()
        torch.cuda.empty_cache()                # the capture allocates its own pool: hand the warm-up's cached blocks back first
        launches_a = lib.msda_total_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            graph_loss = step_body(srcs)
        launches_per_replay = lib.msda_total_launch_count() - launches_a
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    lib.msda_profile_enable(0 if args.cuda_graph else 1)     # event pairs cannot be read back out of a captured graph
    launches0 = lib.msda_total_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    t1 = time.time()
    lib.msda_profile_enable(0)
    launches = lib.msda_total_launch_count() - launches0
    if args.cuda_graph:
        launches = launches_per_replay * args.steps          # replayed launches do not pass through the library's counter
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_per_step = D.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    records = _lib.profile_collect()
    fwd_ms = [t for t, k in records if k == 1]
    bwd_ms = [t for t, k in records if k == 2]
    if not bwd_ms:                              # tiled kernels: dots + scatter make one backward
        d3, d4 = [t for t, k in records if k == 3], [t for t, k in records if k == 4]
        bwd_ms = [a + b for a, b in zip(d3, d4)]
    final_loss = float(loss.detach())
    if args.cuda_graph and not (fwd_ms or bwd_ms):
        eager = eager_records
        fwd_ms = [t for t, k in eager if k == 1]
        bwd_ms = [t for t, k in eager if k == 2]
        if not bwd_ms:
            d3, d4 = [t for t, k in eager if k == 3], [t for t, k in eager if k == 4]
            bwd_ms = [a + b for a, b in zip(d3, d4)]
        msda_ms_per_step = sum(fwd_ms) + sum(bwd_ms)
    else:
        msda_ms_per_step = (sum(bwd_ms) + sum(fwd_ms)) / args.steps if (bwd_ms or fwd_ms) else None

    e2e = None
    if not args.no_e2e:
        step(True)
        barrier()
        e0.record()
        for _ in range(args.e2e_steps):
            step(True)
        e1.record()
        barrier()
        e2e_ms = D.max_over_ranks(e0.elapsed_time(e1), dev) / args.e2e_steps
        e2e = {"value": pts_per_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": args.e2e_steps,
               "h2d_bytes_per_step": sum(t.numel() * 4 for t in host_srcs), "d2h_bytes_per_step": 4,
               "api": ("TransformerDecoder.forward/backward + GradientBuckets + AdamW; encoder memory copied from pinned host "
                       "memory every step, loss copied back") if decoder_step else
                      ("MSDeformAttnTransformerEncoderOnly.forward/backward + GradientBuckets + AdamW; feature pyramid copied "
                       "from pinned host memory every step, loss copied back")}
    peak_mem = round(torch.cuda.max_memory_allocated(dev) / 1e9, 2)
    grad_bytes, n_buckets = int(buckets.flat.numel() * 4) if world > 1 else 0, len(buckets.slices)
    del enc, opt, buckets, srcs, graph
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = measured_peak_gbs()
    ab = W.algorithmic_bytes(count, S, cfg["queries"] if decoder_step else S, M, C // M, L, P, 2)
    if args.fused_preop:        # sampling locations / attention weights never reach HBM: raw offsets + logits instead (same sizes)
        pass
    roofline = None
    if fwd_ms and not bwd_ms:
        fwd_avg = statistics.mean(fwd_ms)
        roofline = {"bound": "hbm", "kernel": "msda_fwd (forward gather)", "achieved": ab["fwd"] / (fwd_avg * 1e-3) / 1e9,
                    "peak": peak, "unit": "GB/s", "frac": ab["fwd"] / (fwd_avg * 1e-3) / 1e9 / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": ab["fwd"], "avg_launch_ms": fwd_avg,
                    "launches_timed": len(fwd_ms), "msda_share_of_step": msda_ms_per_step / ms_per_step}
    if bwd_ms:
        bwd_avg, fwd_avg = statistics.mean(bwd_ms), statistics.mean(fwd_ms)
        roofline = {"bound": "hbm", "kernel": "msda_bwd (backward gather + grad_value scatter)",
                    "achieved": ab["bwd"] / (bwd_avg * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": ab["bwd"] / (bwd_avg * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ab["bwd"], "avg_launch_ms": bwd_avg, "launches_timed": len(bwd_ms),
                    "forward": {"avg_launch_ms": fwd_avg, "algorithmic_bytes_per_launch": ab["fwd"],
                                "frac": ab["fwd"] / (fwd_avg * 1e-3) / 1e9 / peak},
                    "msda_share_of_step": msda_ms_per_step / ms_per_step}
    line = {
        "metric": METRIC, "value": pts_per_step / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "global_batch": global_batch, "per_gpu_batch": count, "layers": layers,
                   "levels": shapes, "queries": cfg["queries"] if decoder_step else S, "heads": M, "head_dim": C // M,
                   "points": P, "d_ffn": cfg["d_ffn"], "shared_value_proj": bool(args.shared_value_proj),
                   "fused_preop": bool(args.fused_preop or args.fused_layers), "fused_layers": bool(args.fused_layers),
                   "cuda_graph": bool(args.cuda_graph), "forward_only": bool(args.forward_only),
                   "optimizer": "AdamW(fused)", "autocast": "bf16",
                   "parallelism": f"dp{world}", "grad_allreduce_bytes": grad_bytes,
                   "grad_buckets": n_buckets, "points_per_step": pts_per_step, "final_loss": final_loss,
                   "peak_device_memory_gb": peak_mem,
                   "msda_share_of_step": msda_ms_per_step / ms_per_step if msda_ms_per_step else None,
                   "msda_ms_per_step": msda_ms_per_step,
                   "l2_policy": "activations of one step (GBs) >> 126 MB L2; no flush needed"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
    }
    return line


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def main():
    args = parse_args()
    if args.e2e_steps is None:
        args.e2e_steps = args.steps
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch ourselves one process per GPU
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    from vision_instance_seg_b200 import workloads as W
    try:
        if W.CONFIGS[args.workload]["kind"] in ("train_step", "decoder_step"):
            return run_train_step(args)
        return run_b200(args)
    finally:
        _shutdown()


if __name__ == "__main__":
    sys.exit(main())
