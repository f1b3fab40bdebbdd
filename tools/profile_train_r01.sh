#!/bin/bash
# ncu capture of the encoder-layer glue kernels + fused sampling kernels inside a fused training step
# (2 layers, batch 4 keeps the ~40x kernel replay short).  Run only after the same command exited 0 without ncu.
set -x
TAG=${1:-r01}
TCMD="python bench.py --workload cfg3_train_step_1024 --batch 4 --layers 2 --steps 1 --warmup 1 --fused-layers --no-e2e"
$TCMD > gpurun_out/plain2_train_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'add_cast_kernel|add_layernorm|colsum_kernel|msda_fwd_vec|msda_bwd_vec' -c 26 -f -o gpurun_out/prof_train_${TAG} $TCMD > gpurun_out/ncu_full_train_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -5
