#!/bin/bash
# round 2, GPU run 24: timeline of one CTA of the solve kernel (RQ_TRACE build), with and without the work
export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_trace.so
export FLUIDB200_RBQ_TRACE=/tmp/rbq_trace.bin
for x in 0 31 1; do
  for it in 8 1; do FLUIDB200_RBQ_X=$x timeout 120 python tools/rbq_trace.py $it 2>&1 | tail -12; done
done
