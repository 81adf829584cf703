#!/bin/bash
# A/B of two builds of libmqe_b200.so on ONE box: alternates the libraries so that box-to-box and thermal drift cancel.
#   tools/ab_bench.sh <libA> <libB> [rounds] [extra bench.py args]
A=$1; B=$2; R=${3:-3}; shift 3
for r in $(seq 1 $R); do
  for L in "$A" "$B"; do
    MQE_B200_LIB=$L python bench.py --steps 300 --no-sublines --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1])
print('$L'.split('/')[-3] if '$L'.count('/')>2 else '$L', 'ms/step %.4f  value %.3f M  e2e %.3f M  k_substeps %.4f  launches %.1f' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['kernel_ms'], d['launches_per_step']))"
  done
done
