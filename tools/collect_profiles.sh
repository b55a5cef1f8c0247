#!/bin/bash
# Collect the round's evidence on a GPU box (run through gpurun); outputs land in gpurun_out/final/.
#   1 the parity suite     2 bench lines (not under a profiler): default (= BASELINE config[2], with the config[3] leg, the
#   residual triple and the CPU baseline inside), --impl reference, jet4096, config 5 (projection) at 4096^2 and 32768^2
#   3 launch list of the default bench      4 one full ncu capture per kernel of the same command
set -x
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench_karman4096.json 2> $O/bench_karman4096.err
timeout 1200 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload jet4096 --no-secondary > $O/bench_jet4096.json 2> $O/bench_jet4096.err
timeout 600 python bench.py --workload project4096 --steps 20 --warmup 3 > $O/bench_project4096.json 2> $O/bench_project4096.err
timeout 600 python bench.py --workload project32768 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_project32768_1gpu.json 2> $O/bench_project32768_1gpu.err
# profiler passes over a short developed-flow run (3 warm-up + 3 + 300 pre-roll + 3 + 3 steps): every launch for the list
# (make_profiles.py keeps the last three steps), one full capture per kernel near the end (8 matching launches per step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_karman4096.csv \
    python bench.py --steps 3 --warmup 3 --preroll 300 --min-timed-steps 3 --no-secondary --no-cpu-baseline > $O/launches.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_rbq|k_confine|k_advect|k_bfecc" -s 2440 -c 9 -o $O/step_full -f \
    python bench.py --steps 3 --warmup 3 --preroll 300 --min-timed-steps 3 --no-secondary --no-cpu-baseline > $O/step_full.log 2>&1
python tools/sass_summary.py final > /dev/null 2>&1; mv profiles/final_sass_summary.md $O/sass_summary.md 2>/dev/null
for f in $O/*.err; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
