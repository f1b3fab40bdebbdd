"""Many-seed soak of the scratch sizing bound: random (Lq, P, level shapes incl. 1-pixel levels, depth) through the raw C ABI
with canary bands around every output and the scratch (see tests/test_gpu_guards.py).  Development aid."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vision_instance_seg_b200 as pkg
from tests.test_gpu_guards import Guarded, DT

lib = pkg.load_library()
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 300
bad = 0
for seed in range(lo, hi):
    rng = random.Random(seed)
    L = rng.randint(1, 6)
    shapes = [(rng.choice([1, 1, 2, 3, 7, 16, 33]), rng.choice([1, 2, 5, 8, 21, 40])) for _ in range(L)]
    N, M, D = rng.randint(1, 3), rng.choice([1, 2, 8]), rng.choice([16, 32, 64, 128])
    Lq, P = rng.choice([1, 7, 64, 300, 1500, 5000]), rng.randint(1, 6)
    depth = rng.choice([0, 1, 4, 32, 1000, 65535])
    flags = (depth & 0xffff) << 8
    dtype = rng.choice([torch.bfloat16, torch.float16])
    ss = torch.tensor(shapes, dtype=torch.long, device="cuda")
    lsi = torch.cat((ss.new_zeros(1), ss.prod(1).cumsum(0)[:-1]))
    S = int(ss.prod(1).sum())
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.randn(N, S, M, D, generator=g, device="cuda").to(dtype)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, device="cuda") * 1.2 - 0.1
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, device="cuda"), -1).view(N, Lq, M, L, P).contiguous()
    go = torch.randn(N, Lq, M * D, generator=g, device="cuda").to(dtype)
    st = torch.cuda.current_stream().cuda_stream
    gv, gl, ga = Guarded(v.numel() * 2), Guarded(loc.numel() * 4), Guarded(attn.numel() * 4)
    nscratch = lib.msda_backward_scratch_bytes(N, S, M, D, Lq, L, P, DT[dtype], flags)
    scratch = Guarded(nscratch)
    rc = lib.msda_backward(v.data_ptr(), ss.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(), go.data_ptr(),
                           gv.ptr, gl.ptr, ga.ptr, scratch.ptr, nscratch, N, S, M, D, Lq, L, P, DT[dtype], 64, flags, st)
    torch.cuda.synchronize()
    ok = rc == 0 and all(b.intact() for b in (gv, gl, ga, scratch)) and bool(torch.isfinite(gv.view(dtype, tuple(v.shape)).float()).all())
    if not ok:
        bad += 1
        print("FAIL", seed, dict(N=N, M=M, D=D, Lq=Lq, L=L, P=P, shapes=shapes, depth=depth, dtype=str(dtype)), "rc", rc, flush=True)
print("done, failures:", bad)
