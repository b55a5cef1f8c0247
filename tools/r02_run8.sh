#!/bin/bash
# round 2, GPU run 8: branch-free tile sampling (ILP), ring solve without hint
set -x
O=gpurun_out/r02_run8; mkdir -p $O
K="fused_path or single_phase or presets_exact or slab or step_local or ghost or quirk"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -x -q -k "$K or golden" > $O/pytest_quick.txt 2>&1; rc=$?; tail -6 $O/pytest_quick.txt
timeout 300 python bench.py --no-cpu-baseline --no-secondary --preroll 600 --min-timed-steps 60 > $O/default.json 2> $O/default.err
python - <<PY
import json
d=json.load(open('$O/default.json'))
print('default', 'ms/step', round(d['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items()))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_advect_velocity_tile|k_bfecc_velocity_tile|k_confine" -s 20 -c 3 -o $O/r02_tiles -f \
    python bench.py --steps 3 --warmup 3 --preroll 300 --min-timed-steps 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1
tail -2 $O/ncu.log
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
