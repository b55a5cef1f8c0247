#!/bin/bash
# round 2, GPU run 10: smoke + confinement tile kernels (tensor-map TMA): full parity suite, bench, ncu
set -x
O=gpurun_out/r02_run10; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "single_phase or reference_suite or golden or presets_exact or slab or fused_path or step_local" > $O/pytest.txt 2>&1; tail -12 $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline --no-secondary --preroll 600 --min-timed-steps 60 > $O/default.json 2> $O/default.err
python - <<PY
import json
d=json.load(open('$O/default.json'))
print('default', 'ms/step', round(d['ms_per_step'],4), 'quiescent', round(d['quiescent']['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items()))
PY
timeout 300 python bench.py --workload jet4096 --no-cpu-baseline --no-secondary --preroll 600 --min-timed-steps 60 > $O/jet4096.json 2> $O/jet4096.err
python - <<PY
import json
d=json.load(open('$O/jet4096.json'))
print('jet4096', 'ms/step', round(d['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items()))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_confine_tile" -s 2 -c 1 -o $O/r02_tiles2 -f \
    python bench.py --steps 3 --warmup 3 --preroll 300 --min-timed-steps 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1
tail -2 $O/ncu.log
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
