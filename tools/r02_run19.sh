#!/bin/bash
# round 2, GPU run 19: pair-wait hand-offs (iterations 1..7 wait once per two steps; 32 / 34 / 36 ring slots)
O=gpurun_out/r02_run19; mkdir -p $O
for v in pw32 pw34 pw36; do
  export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so
  echo -n "$v: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
  echo -n "$v X=31: "; FLUIDB200_RBQ_X=31 timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-100
done
