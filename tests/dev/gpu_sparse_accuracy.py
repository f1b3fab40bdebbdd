"""How accurate is the sparse-level direct accumulation (packed bf16 adds straight into grad_value) when the decoder's
queries cluster?  cfg4 geometry, 300 queries, N = 2; box centres drawn inside a square of side `spread` (fraction of the
image), box sizes U(0.02, 0.1) * spread-ish; same-sign and zero-mean grad_output.  Prints the rel-to-max error of grad_value
against the fp64 oracle for the direct mode and for the bucketed mode (MSDA_BWD_NO_SPARSE_DIRECT)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vision_instance_seg_b200 as pkg
from vision_instance_seg_b200 import MultiScaleDeformableAttention as MSDA, workloads as W
from oracle import ms_deform_attn_oracle_grads
from tests.helpers import rel_to_max

pkg.load_library()
dev = "cuda:0"
cfg = W.CONFIGS["cfg4_decoder_300q_bf16"]
ss = W.make_spatial_shapes(cfg["shapes"])
lsi = W.make_level_start_index(ss)
S, L, M, D, P, N, Lq = int(ss.prod(1).sum()), 4, 8, 32, 4, 2, 300
for spread in (1.0, 0.3, 0.1, 0.03):
    for mean in (0.0, 1.0):
        g = torch.Generator().manual_seed(7)
        value = torch.randn(N, S, M, D, generator=g).to(torch.bfloat16)
        ctr = 0.5 + (torch.rand(N, Lq, 1, 2, generator=g) - 0.5) * spread
        wh = (torch.rand(N, Lq, 1, 2, generator=g) * 0.4 + 0.1) * spread
        off = W.init_offset_pattern(M, L, P)[None, None] + torch.randn(N, Lq, M, L, P, 2, generator=g)
        loc = (ctr[:, :, None, :, None, :] + off / P * wh[:, :, None, :, None, :] * 0.5).contiguous()
        attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
        go = (torch.randn(N, Lq, M * D, generator=g) + mean).to(torch.bfloat16)
        want = ms_deform_attn_oracle_grads(value.double(), ss, loc.double(), attn.double(), go.double())[1]
        row = dict(spread=spread, grad_mean=mean)
        for name, flags in (("direct", 0), ("bucketed", 4), ("fp32_accum", 2)):
            MSDA.backward_flags = flags
            gv = MSDA.ms_deform_attn_backward(value.to(dev), ss.to(dev), lsi.to(dev), loc.to(dev), attn.to(dev), go.to(dev), 64)[0]
            row[name] = round(rel_to_max(gv, want), 5)
        # adds per touched level-0 row: how clustered is it?
        px = (loc[..., 0, :, :] * 128).floor().long().clamp(0, 127)
        idx = (px[..., 1] * 128 + px[..., 0]).flatten()
        cnt = torch.bincount(idx, minlength=128 * 128)
        row["max_points_per_level0_pixel"] = int(cnt.max())
        print(json.dumps(row), flush=True)
