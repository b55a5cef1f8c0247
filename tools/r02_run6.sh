#!/bin/bash
# round 2, GPU run 6: ring solve with try_wait hint + scalar-shift update; single-buffer tile kernels, tile-size variants
set -x
O=gpurun_out/r02_run6; mkdir -p $O
K="pressure_form or redblack or fused_path or single_phase or presets_exact or slab or many_chunks or projection or step_local or ghost"
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "$K" > $O/pytest_quick.txt 2>&1; rc=$?; tail -6 $O/pytest_quick.txt
run() { # name
  python - <<PY
import json
try:
    d=json.load(open('$O/$1.json'))
    print('$1', 'ms/step', round(d['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items()))
except Exception as e: print('$1 failed', e)
PY
}
for v in default hint0 hint1000 tile16; do
  if [ $v = default ]; then unset FLUIDB200_LIB; else export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-secondary --preroll 600 --min-timed-steps 60 > $O/$v.json 2> $O/$v.err
  run $v
done
unset FLUIDB200_LIB
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -6 $O/pytest.txt
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
