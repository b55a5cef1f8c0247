#!/bin/bash
# usage: tools/rbq_x_variants.sh "x flags" variant...   -- project-phase time per variant library and flag set
xs=$1; shift
for v in "$@"; do
  for x in $xs; do
    FLUIDB200_LIB=$PWD/tools/variants/lib_$v.so FLUIDB200_RBQ_X=$x python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v X=$x project ms', round(d['roofline']['phases_ms_per_step']['project'],4))"
  done
done
