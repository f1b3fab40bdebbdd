"""Per-level error of the fp16-accumulation mode vs the exact fp32 mode (cfg3 geometry, few images)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from vision_instance_seg_b200 import workloads as W, MultiScaleDeformableAttention as MSDA
from oracle import ms_deform_attn_oracle_grads

dev = "cuda"
cfg = W.CONFIGS["cfg3_swinl_1024_bf16"]
for seed in (1, 2, 3):
    for mean in (0.0, 1.0):
        v, ss, lsi, loc, attn = W.make_encoder_inputs(cfg["shapes"], 2, torch.bfloat16, device=dev, seed=seed)
        go = (torch.randn(2, loc.shape[1], 256, device=dev) + mean).to(torch.bfloat16)
        ref = ms_deform_attn_oracle_grads(v.float(), ss.cpu(), loc, attn, go.float(), dtype=torch.float32)[1]
        res = {}
        for flags in (2, 0):
            MSDA.backward_flags = flags
            gv = MSDA.ms_deform_attn_backward(v, ss, lsi, loc, attn, go, 128)[0].float().cpu()
            mx = ref.abs().max()
            per = []
            for l in range(4):
                a = int(lsi[l]); b = int(lsi[l + 1]) if l < 3 else ref.shape[1]
                per.append(float((gv[:, a:b] - ref[:, a:b]).abs().max() / ref[:, a:b].abs().max()))
            res[flags] = dict(total=float((gv - ref).abs().max() / mx), per_level=per)
        print(json.dumps(dict(seed=seed, mean=mean, exact=res[2], f16_buckets=res[0])), flush=True)
