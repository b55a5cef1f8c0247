#!/bin/bash
# round 2, GPU run 23: one warp per sweep stage (no named barrier between the halves of a line), timing with and without the work
for v in split1 split1w8; do
  export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so
  echo -n "$v: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
  for x in 1 31; do echo -n "$v X=$x: "; FLUIDB200_RBQ_X=$x timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-100; done
done
