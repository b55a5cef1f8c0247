#!/bin/bash
# Collect the round's evidence on a GPU box (run through gpurun); outputs land in gpurun_out/final/.
#   1 launch list of the default bench      2 one full ncu capture per kernel
#   3 bench lines (not under a profiler): default, jet4096, jet16384, --impl reference, config 5 (projection) at 4096^2 and 32768^2
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
O=gpurun_out/final; mkdir -p $O
python bench.py > $O/bench_karman4096.json 2> $O/bench_karman4096.err
python bench.py --workload jet4096 > $O/bench_jet4096.json 2> $O/bench_jet4096.err
python bench.py --workload jet16384 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_jet16384.json 2> $O/bench_jet16384.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload project4096 --steps 20 --warmup 3 > $O/bench_project4096.json 2> $O/bench_project4096.err
python bench.py --workload project32768 --steps 10 --warmup 3 > $O/bench_project32768_1gpu.json 2> $O/bench_project32768_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/launches_karman4096.csv \
    python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $O/launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_rbq|k_confine|k_advect|k_bfecc" -s 24 -c 9 -o $O/step_full -f \
    python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline > $O/step_full.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_gs_wavefront|k_rb_fused" -c 2 -o $O/solvers_full -f \
    python bench.py --workload jet4096 --solver exact --steps 1 --warmup 3 --no-secondary --no-cpu-baseline > $O/solvers_full.log 2>&1
tail -2 $O/*.err | cut -c1-200
