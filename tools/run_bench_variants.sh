#!/bin/bash
# usage: tools/run_bench_variants.sh name...  -- phase times of the default bench with each variant library
for v in "$@"; do
  FLUIDB200_LIB=$PWD/tools/variants/lib_$v.so python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['phases_ms_per_step'].items()})"
done
