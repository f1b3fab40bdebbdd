"""Write profiles/sass_hot_kernels_rNN.md: opcode histogram + the densest FMA stretch of the hot kernels' SASS.

    python tools/sass_excerpt.py [out.md]          (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision-instance-seg_b200", "libmsda_b200.so")
TARGETS = [
    ("direct forward  msda_fwd_vec_kernel<bf16, 32, false, float>", "msda_fwd_vec_kernelI13__nv_bfloat16Li32ELb0EfE"),
    ("direct backward msda_bwd_vec_kernel<bf16, 32, GV16, !FUSED, float, !SPARSE>", "msda_bwd_vec_kernelI13__nv_bfloat16Li32ELb1ELb0EfLb0E"),
    ("tiled forward   msda_fwd_tiled_kernel<bf16>", "msda_fwd_tiled_kernelI13__nv_bfloat16"),
    ("tiled backward 1 msda_bwd_dots_tiled_kernel<bf16>", "msda_bwd_dots_tiled_kernelI13__nv_bfloat16"),
    ("tiled backward 2 msda_bwd_scatter_tiled_kernel<bf16>", "msda_bwd_scatter_tiled_kernelI13__nv_bfloat16"),
]
FORMS = ("FHFMA", "HFMA2", "LDG", "LDS", "STS", "RED", "ATOMS", "LDGSTS", "UBLKCP", "UTMA", "SHFL", "STG", "BAR")


def main(out_path):
    lines = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(lines) if "Function : " in l]
    out = ["# SASS of the hot kernels of libmsda_b200.so (sm_100a)\n",
           "Produced by `python tools/sass_excerpt.py` (`cuobjdump -sass`) from the committed sources.  Per kernel: opcode\n"
           "histogram of the whole function, the memory / FMA instruction forms it contains, and the 60-instruction stretch with\n"
           "the most FMAs (the gather loop).  No `UTMALDG` / `UBLKCP`: the windows of the tiled kernels are filled with\n"
           "`LDGSTS...ZFILL` (cp.async), see DESIGN.md section 3.7; `UTC*MMA` is absent by design (no tensor cores on this path).\n"]
    for title, key in TARGETS:
        body = None
        for k, i in enumerate(starts):
            if key in lines[i]:
                body = lines[i:starts[k + 1] if k + 1 < len(starts) else len(lines)]
                break
        if body is None:
            out.append(f"\n## {title}\n\nnot found in the library\n")
            continue
        ins = []
        for l in body:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                ins.append((m.group(1), m.group(2).strip()))
        ops, forms = collections.Counter(), collections.Counter()
        for _, t in ins:
            parts = t.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += 1
            if op.startswith(FORMS):
                forms[op] += 1
        out.append(f"\n## {title}\n\n{len(ins)} instructions.  Opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(18)) + "\n")
        out.append("Memory / FMA forms: " + ", ".join(f"`{k}` x{v}" for k, v in sorted(forms.items())) + "\n")
        fma = [i for i, (_, t) in enumerate(ins) if "FHFMA" in t or "HFMA2.MMA" in t or " HFMA2 " in f" {t} "]
        if fma and len(ins) > 60:
            best = max(range(0, len(ins) - 60), key=lambda s: sum(1 for i in fma if s <= i < s + 60))
            out.append("```\n" + "\n".join(f"/*{a}*/ {t} ;" for a, t in ins[best:best + 60]) + "\n```\n")
    with open(out_path, "w") as f:
        f.write("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_hot_kernels_r02.md"))
