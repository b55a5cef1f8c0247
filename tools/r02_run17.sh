#!/bin/bash
# round 2, GPU run 17: race hunt on the many-chunks solve (default build, round-1 debug pointer build, elected-lane hand-offs) + timing of the latter
O=gpurun_out/r02_run17; mkdir -p $O
for v in default olddebug elect; do
  if [ $v = default ]; then unset FLUIDB200_LIB; else export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so; fi
  timeout 600 python tools/rbq_race_hunt.py 400 2>&1 | tail -8 | sed "s/^/$v: /"
  HUNT_STATS=0 timeout 600 python tools/rbq_race_hunt.py 200 2>&1 | tail -4 | sed "s/^/$v nostats: /"
  echo -n "$v: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
done
export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_elect.so
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "many_chunks or pressure_form or slab or fused_path or step_local or projection" 2>&1 | tail -2 | sed "s/^/elect pytest: /"
