#!/bin/bash
# round 2, GPU run 3: parity of k_rbq_stream and the tile advection kernels, timings of variants, ncu
set -x
O=gpurun_out/r02_run3; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.txt; tail -3 $O/smoke.txt
K="pressure_form or redblack or fused_path or single_phase or presets_exact or slab or many_chunks or projection"
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "$K" > $O/pytest_quick.txt 2>&1; rc=$?; tail -15 $O/pytest_quick.txt
if [ $rc -ne 0 ]; then
  FLUIDB200_ADV_FULL=1 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "$K" > $O/pytest_quick_advfull.txt 2>&1; tail -15 $O/pytest_quick_advfull.txt
fi
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -15 $O/pytest.txt
for v in default; do
  unset FLUIDB200_LIB
  timeout 300 python bench.py --workload project4096 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > $O/project4096_$v.json 2> $O/project4096_$v.err
  python -c "import json; d=json.load(open('$O/project4096_$v.json')); print('$v', d['ms_per_step'], d['roofline']['frac'])"
done
unset FLUIDB200_LIB
timeout 900 python bench.py > $O/karman4096.json 2> $O/karman4096.err
tail -c 300 $O/karman4096.err
python - <<PY
import json
d=json.load(open('$O/karman4096.json'))
print('ms/step', d['ms_per_step'], 'quiescent', d['quiescent']['ms_per_step'], 'step frac', d['roofline']['step']['frac'])
for k,v in d['roofline']['kernels'].items(): print(' ', k, round(v['ms_per_launch'],4), round(v['frac'],3), v['launches_per_step'])
print(d['config']['residual']); print(d['e2e'] and d['e2e']['value'], d['cpu_baseline'])
PY
FLUIDB200_ADV_FULL=1 timeout 600 python bench.py --no-cpu-baseline --no-secondary > $O/karman4096_advfull.json 2> $O/karman4096_advfull.err
python -c "import json; d=json.load(open('$O/karman4096_advfull.json')); print('advfull ms/step', d['ms_per_step']); [print(' ', k, round(v['ms_per_launch'],4), round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()]"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_rbq_stream|k_advect_velocity_tile|k_bfecc_velocity_tile" -s 30 -c 4 -o $O/r02_full -f \
    python bench.py --steps 3 --warmup 3 --preroll 300 --min-timed-steps 3 --no-cpu-baseline --no-secondary > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
