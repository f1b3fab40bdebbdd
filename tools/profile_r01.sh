#!/bin/bash
# Round-1 profiling recipe (B200_PROFILING.md): plain run first, then launch list, then one full capture.
set -x
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --layers 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 8 -c 5 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -20
