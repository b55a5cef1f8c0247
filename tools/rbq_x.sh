#!/bin/bash
# project-phase time of the default bench with parts of k_rbq_fused switched off (FLUIDB200_RBQ_X bits:
# 1 skip the sweeps, 2 skip the writer's I/O, 4 skip the TMA copies) -- results are wrong, timing only
for x in ${@:-0 1 2 4 3 5 6 7}; do
  FLUIDB200_RBQ_X=$x python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('X=$x project ms', round(d['roofline']['phases_ms_per_step']['project'],4))"
done
