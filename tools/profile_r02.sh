#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md): each command runs plain first, then under ncu, in ONE gpurun call.
#   gpurun --timeout 1500 -- 'bash tools/profile_r02.sh'
# Outputs land in gpurun_out/; summaries for profiles/: python tools/ncu_summary.py gpurun_out/<x>.ncu-rep profiles/<y>.json
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-secondary"
# launch list of the bench command (compare SHARES, not absolutes)
$BENCH > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r02.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
# the three tiled kernels (opt-in) at batch 4 of the cfg3 geometry
MSDA_B200_TILED=1 python tests/dev/profile_tiled.py 4 > gpurun_out/plain_tiled.log 2>&1 &&
MSDA_B200_TILED=1 ncu --set full --clock-control none --import-source on -k regex:tiled -s 3 -c 3 -f -o gpurun_out/prof_tiled_v3 \
    python tests/dev/profile_tiled.py 4 > gpurun_out/ncu_tiled.log 2>&1
# the direct kernels (default) on the same inputs
python tests/dev/profile_tiled.py 4 > gpurun_out/plain_direct.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 6 -f -o gpurun_out/prof_direct_r02 \
    python tests/dev/profile_tiled.py 4 > gpurun_out/ncu_direct.log 2>&1
ls -la gpurun_out/ | tail -20
