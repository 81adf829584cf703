#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list of one short bench run + one full-set capture of the two top kernels.
# Run under gpurun from the repo root; results land in gpurun_out/.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_substeps -s 5 -c 2 -f -o gpurun_out/prof_substeps \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_substeps.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_policy_l0_tc -s 5 -c 1 -f -o gpurun_out/prof_policy_tc \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_policy.log 2>&1
ls -la gpurun_out
