#!/bin/bash
# round 2, GPU run 15: is the many-chunks mismatch tied to the per-CTA debug pointer build?  (5 repetitions per build)
O=gpurun_out/r02_run15; mkdir -p $O
for v in default olddebug default olddebug; do
  if [ $v = default ]; then unset FLUIDB200_LIB; else export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so; fi
  for k in 1 2 3; do
    timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "many_chunks or pressure_form" 2>&1 | tail -1 | sed "s/^/$v $k: /"
  done
done
