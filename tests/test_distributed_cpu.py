"""CPU, world_size 2 over gloo: the N>1 host logic of the batch-sharded path (sharding, max-over-ranks timing,
gradient-bucket allreduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vision_instance_seg_b200 import distributed as D
    r, lr, w = D.init_process_group("gloo")
    assert (r, w) == (rank, world)
    start, count = D.shard_batch(5, world, rank)
    t_max = D.max_over_ranks(10.0 + rank, "cpu")
    t_sum = D.sum_over_ranks(float(count), "cpu")
    bucket = D.GradientBucket(numel=1000, device="cpu")
    bucket.flat.fill_(float(rank + 1))
    bucket.allreduce_async()
    bucket.wait()
    # per-layer buckets reduced from backward hooks: grads must come out as the mean over the ranks
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3))
    buckets = D.GradientBuckets([list(net[2].parameters()), list(net[0].parameters())], device="cpu")
    x = torch.full((2, 4), float(rank + 1))
    net(x).square().sum().backward()
    buckets.wait()
    local = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3))
    local.load_state_dict(net.state_dict())
    want = None
    for rk in range(world):
        local.zero_grad()
        local(torch.full((2, 4), float(rk + 1))).square().sum().backward()
        gs = torch.cat([p.grad.flatten() for p in list(local[2].parameters()) + list(local[0].parameters())])
        want = gs / world if want is None else want + gs / world
    ok_buckets = bool(torch.allclose(buckets.flat, want, rtol=1e-5, atol=1e-6)) and buckets.launched == 2 \
        and net[0].weight.grad.data_ptr() == buckets.flat[buckets.slices[1][0]:].data_ptr()
    buckets.zero()
    ok_buckets = ok_buckets and float(net[0].weight.grad.abs().max()) == 0.0
    q.put((rank, start, count, t_max, t_sum, float(bucket.flat[0]), float(bucket.flat[-1]), ok_buckets))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 2)]          # 5 images over 2 ranks
    assert all(r[3] == 11.0 for r in res)                            # max over ranks
    assert all(r[4] == 5.0 for r in res)                             # all images accounted for
    assert all(r[5] == 1.5 and r[6] == 1.5 for r in res)             # mean of (1, 2)
    assert all(r[7] for r in res)                                    # hook-driven per-layer buckets


def test_shard_batch_covers_everything():
    from vision_instance_seg_b200.distributed import shard_batch
    for g in (0, 1, 7, 16, 17):
        for w in (1, 2, 4, 8):
            spans = [shard_batch(g, w, r) for r in range(w)]
            assert sum(c for _, c in spans) == g
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
    with pytest.raises(ValueError):
        shard_batch(4, 2, 2)


def test_gradient_buckets_survive_zero_grad_set_to_none():
    """after ``zero_grad(set_to_none=True)`` autograd creates fresh ``.grad`` tensors outside the bucket; the
    hook must adopt them (copy into the slice, re-install the view) instead of reducing a slice of zeros."""
    import torch
    from vision_instance_seg_b200.distributed import GradientBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    buckets = GradientBuckets([list(net[1].parameters()), list(net[0].parameters())], device="cpu")
    x = torch.randn(5, 4)
    net(x).square().sum().backward()
    buckets.wait()
    want = [p.grad.clone() for p in net.parameters()]
    assert all(p.grad.data_ptr() == buckets._view_of[id(p)].data_ptr() for p in net.parameters())
    net.zero_grad(set_to_none=True)                 # what optimizers and modules do by default
    buckets.zero()
    assert all(p.grad is not None and p.grad.data_ptr() == buckets._view_of[id(p)].data_ptr() for p in net.parameters())
    net.zero_grad(set_to_none=True)                 # ... and without calling zero() afterwards
    buckets._pending = list(buckets._sizes)
    net(x).square().sum().backward()
    buckets.wait()
    for p, w in zip(net.parameters(), want):
        assert p.grad.data_ptr() == buckets._view_of[id(p)].data_ptr()
        assert torch.allclose(p.grad, w)
    off = 0
    for p in list(net[1].parameters()) + list(net[0].parameters()):
        assert torch.allclose(buckets.flat[off:off + p.numel()].view_as(p), p.grad)
        off += p.numel()
