#!/bin/bash
# round 2, GPU run 25: the writer hands the ring slots back before it computes and stores (RQ_WEARLY): timing, determinism, timeline
for v in wearly wearly4; do
  export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so
  echo -n "$v: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
  timeout 120 python tools/rbq_race_hunt.py 60 2>&1 | tail -2 | sed "s/^/$v: /"
done
export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_wearlyt.so
export FLUIDB200_RBQ_TRACE=/tmp/rbq_trace.bin
timeout 120 python tools/rbq_trace.py 8 2>&1 | tail -11
