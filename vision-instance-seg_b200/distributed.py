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


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Best effort: pin this process (and therefore the first-touch placement of the pinned host buffers it allocates
    afterwards) to the CPUs that are local to GPU `local_rank`'s PCIe root, read from sysfs.  With one process per GPU and
    host-resident operands (bench.py's `e2e` leg) this keeps every rank's H2D / D2H traffic on its own socket.  Returns a
    short description of what was done; never raises."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"gpu {bus}: local cpus {text or '?'} (no change)"
        os.sched_setaffinity(0, cpus)
        return f"gpu {bus}: bound to cpus {text}"
    except Exception as exc:      # no sysfs, no permission, older torch: run unbound
        return f"unbound ({type(exc).__name__})"


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


class GradientBuckets:
    """Gradients of a module tree as views into one flat buffer, all-reduced group by group *during* backward.

    `groups` is a list of parameter lists (e.g. one per encoder layer, in forward order).  Every parameter's `.grad`
    is a view into `flat`, so autograd accumulates straight into the bucket (no copy in, no copy out).  A
    post-accumulate hook counts a group's parameters down; when the last one has its gradient, the group's slice is
    averaged over the ranks on a side stream (NCCL over NVLink / NVSwitch on the B200 box), overlapping the backward of
    the layers below it.  `wait()` joins the side stream; `zero()` clears the bucket for the next step.  This is the one
    exchange a batch-sharded training step of the path has — the operator itself needs none."""

    def __init__(self, groups, device=None):
        groups = [[p for p in g if p.requires_grad] for g in groups]
        groups = [g for g in groups if g]
        params = [p for g in groups for p in g]
        if not params:
            raise ValueError("no trainable parameters")
        self.device = torch.device(device) if device is not None else params[0].device
        dtype = params[0].dtype
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=dtype, device=self.device)
        self.slices, self._group_of, self._handles, self._view_of = [], {}, [], {}
        off = 0
        for gi, g in enumerate(groups):
            start = off
            for p in g:
                view = self.flat[off:off + p.numel()].view_as(p)
                p.grad = view
                self._view_of[id(p)] = view
                off += p.numel()
                self._group_of[id(p)] = gi
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self.slices.append((start, off))
        self._sizes = [len(g) for g in groups]
        self._pending = list(self._sizes)
        self._params = params
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.launched = 0
        # True while the side stream holds reductions the main stream has not joined yet.  Joining only then keeps the
        # class usable inside a CUDA-graph capture: a capturing stream may not wait on a stream that was never forked
        # from it (cudaErrorStreamCaptureIsolation).
        self._side_pending = False

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _on_grad(self, p):
        # `optimizer.zero_grad()` / `module.zero_grad()` default to set_to_none=True: autograd then hands the parameter a
        # fresh `.grad` tensor that is NOT a view into the bucket, and reducing the bucket slice would silently average
        # zeros.  Adopt such a gradient: copy it into the slice and re-install the view (callers should use `zero()`).
        view = self._view_of[id(p)]
        if p.grad is None:
            raise RuntimeError("GradientBuckets: post-accumulate hook fired without a gradient")
        if p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)
            p.grad = view
        gi = self._group_of[id(p)]
        self._pending[gi] -= 1
        if self._pending[gi] == 0:
            self._reduce(gi)

    def _reduce(self, gi):
        self.launched += 1
        if not self._distributed():
            return
        a, b = self.slices[gi]
        chunk = self.flat[a:b]
        world = dist.get_world_size()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))     # the group's gradients are complete
            with torch.cuda.stream(self.stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM)
                chunk.div_(world)
            self._side_pending = True
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM)
            chunk.div_(world)

    def wait(self):
        """Join the reductions; afterwards every rank holds the averaged gradients in `.grad` / `flat`."""
        if any(n != 0 and n != s for n, s in zip(self._pending, self._sizes)):
            raise RuntimeError("backward left a gradient group incomplete (unused parameters?)")
        self._join()

    def _join(self):
        if self.stream is not None and self._side_pending:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            self._side_pending = False

    def zero(self):
        """Clear the bucket for the next step (use this instead of ``zero_grad(set_to_none=True)``; gradients that were
        detached from the bucket anyway are re-attached here and by the hook)."""
        self._join()
        self.flat.zero_()
        self._reattach()
        self._pending = list(self._sizes)
        self.launched = 0

    def _reattach(self):
        for p in self._params:
            if p.grad is None or p.grad.data_ptr() != self._view_of[id(p)].data_ptr():
                p.grad = self._view_of[id(p)]

    def remove_hooks(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def encoder_gradient_groups(encoder_only):
    """Gradient groups of an MSDeformAttnTransformerEncoderOnly in the order backward completes them:
    last layer first, then ... layer 0, then level_embed."""
    groups = [list(layer.parameters()) for layer in reversed(list(encoder_only.encoder.layers))]
    groups.append([encoder_only.level_embed])
    return groups
