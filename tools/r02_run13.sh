#!/bin/bash
# round 2, GPU run 13: k_confine_fast with scaled-exact fall-backs and quad halo rows: parity, self-test, bench
set -x
O=gpurun_out/r02_run13; mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fastmath or extremes" -s > $O/pytest_fast.txt 2>&1; grep -E "FASTMATH|passed|failed" $O/pytest_fast.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "single_phase or reference_suite or golden or presets_exact or slab or fused_path or step_local or u8 or view" > $O/pytest.txt 2>&1; tail -4 $O/pytest.txt
timeout 400 python bench.py --no-cpu-baseline --no-secondary --min-timed-steps 60 > $O/default.json 2> $O/default.err
python - <<PY
import json
d=json.load(open('$O/default.json'))
print('default', 'ms/step', round(d['ms_per_step'],4), 'quiescent', round(d['quiescent']['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items()))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_confine_fast" -s 1240 -c 1 -o $O/r02_confine_fast -f \
    python bench.py --steps 3 --warmup 3 --min-timed-steps 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1
tail -2 $O/ncu.log | cut -c1-200
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
