# CTA granularity of k_substeps on go1sheep-hard (one env per warp): warps per CTA x phase-alignment barrier.  Measured r2: 8 warps + barrier 1.26 ms,
# 8 warps without 1.86 ms, 4 / 2 / 1 warps per CTA 1.75 / 1.83 / 2.28 ms -- the 270 KB instruction stream of the substep loop is fetched once per
# CTA when its warps walk it together, and once per warp when they do not (L1.5 I-cache = 32 KB).
run() { python bench.py --task go1sheep-hard --num-envs 4096 --steps 60 --preroll-episodes 1 --no-sublines --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('$1', 'ms/step %.4f  value %.3f M  k_substeps %.4f' % (d['ms_per_step'], d['value']/1e6, d['roofline']['kernel_ms']))"; }
run default
for w in 1 2 4; do for c in 0 1; do MQE_SUBSTEP_WARPS=$w MQE_CTA_SYNC=$c run "warps=$w cta_sync=$c"; done; done
MQE_CTA_SYNC=0 run "warps=default cta_sync=0"
