#!/bin/bash
# round 2, GPU run 26: 30 / 31 / 32 ring slots with the FULL writer staging ring (the free shared memory allows 32)
for v in nl30 nl31 nl32; do
  export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so
  echo -n "$v: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
done
