#!/bin/bash
# Profiling recipe (B200_PROFILING.md), round 2: launch list of one short steady-state bench run + one full-set capture of k_substeps and
# of the three policy kernels.  Run under gpurun from the repo root; results land in gpurun_out/ (copy summaries into profiles/).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --preroll-episodes 1 --no-sublines --no-cpu-baseline"
# launch list: skip the 501 pre-roll steps (5 launches each + reset), list the launches of ~12 timed steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 2560 -c 80 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_substeps -s 516 -c 1 -f -o gpurun_out/prof_substeps $B > gpurun_out/ncu_substeps.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_policy -s 1545 -c 4 -f -o gpurun_out/prof_policy $B > gpurun_out/ncu_policy.log 2>&1
ls -la gpurun_out
