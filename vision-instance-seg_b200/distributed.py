"""Multi-GPU plumbing for the batch-sharded path: one process per GPU, torch.distributed
(NCCL over NVLink 5 / NVSwitch on the B200 box; gloo in the CPU tests).  The operator itself needs no
collective — output row (b, q) depends only on value[b], loc[b, q], attn[b, q] — so ranks own disjoint
images.  The only exchange of a training step is the DDP-style sum-allreduce of the encoder's parameter
gradients (6 layers x 1 282 176 + 1 024 level-embed = 7 694 080 fp32 = 30.8 MB), issued on a side stream so
that it hides under the next layer's backward."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

#: parameters of a 6-layer MaskDINO pixel-decoder encoder (MSDeformAttn + LayerNorms + FFN per layer) + level_embed
ENCODER_GRAD_ELEMENTS = 6 * 1282176 + 1024


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: str | None = None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def shard_batch(global_batch: int, world: int, rank: int):
    """Contiguous image range [start, start + count) owned by `rank` (strong scaling of a fixed global batch);
    remainders go to the lowest ranks."""
    if global_batch < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("invalid shard request")
    base, rem = divmod(global_batch, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class GradientBucket:
    """One flat fp32 bucket standing for the encoder's parameter gradients; `allreduce_async()` averages it
    over the ranks on a side stream (CUDA) so the transfer overlaps the compute that follows."""

    def __init__(self, numel: int = ENCODER_GRAD_ELEMENTS, device="cuda"):
        self.device = torch.device(device)
        self.flat = torch.zeros(numel, dtype=torch.float32, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._work = None

    def allreduce_async(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
                self.flat.div_(world)
        else:
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        elif self._work is not None:
            self._work.wait()
            self.flat.div_(dist.get_world_size())
            self._work = None
