#!/bin/bash
# round 2, GPU run 28: the sweep pair loop unrolled by 2 / 3 / 4 (register rotation of the carried vectors)
for v in pu2 pu3 pu4; do
  export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so
  echo -n "$v: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
done
