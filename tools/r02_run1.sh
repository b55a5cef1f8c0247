#!/bin/bash
# round 2, first GPU run: parity of the new stream kernel + timings of its variants
set -x
O=gpurun_out/r02_run1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest.txt; cat $O/pytest.txt
for v in default rsf0; do
  if [ $v = default ]; then unset FLUIDB200_LIB; else export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so; fi
  timeout 300 python bench.py --workload project4096 --steps 20 --warmup 3 --no-cpu-baseline > $O/project4096_$v.json 2> $O/project4096_$v.err
  tail -c 1500 $O/project4096_$v.json
done
unset FLUIDB200_LIB
FLUIDB200_RBQ_RING=1 timeout 300 python bench.py --workload project4096 --steps 20 --warmup 3 --no-cpu-baseline > $O/project4096_ring.json 2> $O/project4096_ring.err
timeout 600 python bench.py --no-cpu-baseline --no-secondary > $O/karman4096.json 2> $O/karman4096.err
tail -c 3000 $O/karman4096.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_rbq_stream" -s 3 -c 1 -o $O/rbq_stream_full -f \
    python bench.py --workload project4096 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
tail -3 $O/*.err | cut -c1-300
