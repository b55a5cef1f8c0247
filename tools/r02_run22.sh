#!/bin/bash
# round 2, GPU run 22: the committed default bench line (traffic stamped to the current sources) + the GPU suite once more
O=gpurun_out/final; mkdir -p $O
timeout 900 python bench.py > $O/bench_karman4096.json 2> $O/bench_karman4096.err; tail -2 $O/bench_karman4096.err
python -c "
import json; d=json.load(open('$O/bench_karman4096.json')); r=d['roofline']; print('ms', d['ms_per_step'], 'traffic', r['traffic'], r['traffic_source'], 'frac', r['frac'])"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee $O/pytest_gpu.txt
