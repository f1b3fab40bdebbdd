"""Many-seed soak of tests/test_gpu_fuzz.py (development aid; run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tests.test_gpu_fuzz as F
import vision_instance_seg_b200 as pkg

pkg.load_library()
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 1400
bad = 0
for seed in range(lo, hi):
    for dtype in (torch.float32, torch.bfloat16):
        for fn in (F.test_random_shapes_plain_operator, F.test_random_shapes_fused_operator):
            try:
                fn(pkg, seed, dtype)
                torch.cuda.synchronize()
            except AssertionError as e:
                bad += 1
                print("FAIL", fn.__name__, seed, dtype, F._draw(seed), str(e)[:200], flush=True)
            except Exception as e:
                print("ERROR", fn.__name__, seed, dtype, F._draw(seed), repr(e)[:200], flush=True)
                raise
print("done, failures:", bad)
