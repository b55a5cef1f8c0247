#!/bin/bash
# round 2, GPU run 4: k_rbq_stream v3 (one warp, eight iterations, lag 3): parity + timing + ncu
set -x
O=gpurun_out/r02_run4; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.txt
K="pressure_form or many_chunks or projection or slab or step_local"
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "$K" > $O/pytest_quick.txt 2>&1; rc=$?; tail -12 $O/pytest_quick.txt
timeout 300 python bench.py --workload project4096 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > $O/project4096.json 2> $O/project4096.err
python -c "import json; d=json.load(open('$O/project4096.json')); print('project4096', d['ms_per_step'], d['roofline']['frac'])"
timeout 600 python bench.py --no-cpu-baseline --no-secondary > $O/karman4096.json 2> $O/karman4096.err
python -c "import json; d=json.load(open('$O/karman4096.json')); print('ms/step', d['ms_per_step']); [print(' ', k, round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()]"
timeout 600 python bench.py --workload jet16384 --no-cpu-baseline --no-secondary --steps 10 > $O/jet16384.json 2> $O/jet16384.err
python -c "import json; d=json.load(open('$O/jet16384.json')); print('jet16384 ms/step', d['ms_per_step']); [print(' ', k, round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()]"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_rbq_stream" -s 3 -c 1 -o $O/r02_rbq -f \
    python bench.py --workload project4096 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1
tail -2 $O/ncu.log
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
