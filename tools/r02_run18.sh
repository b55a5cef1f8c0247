#!/bin/bash
# round 2, GPU run 18: elected-lane hand-offs (one arrive / one poller per warp): determinism, timing, parity
O=gpurun_out/r02_run18; mkdir -p $O
export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_elect.so
timeout 120 python tools/rbq_race_hunt.py 100 2>&1 | tail -4 | sed "s/^/elect: /"
echo -n "elect: "; timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-130
for x in 1 31; do echo -n "elect X=$x: "; FLUIDB200_RBQ_X=$x timeout 60 python tools/rbq_iters.py 1 8 2>&1 | tail -1 | cut -c1-100; done
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "many_chunks or pressure_form or slab or fused_path or step_local or projection" 2>&1 | tail -2 | sed "s/^/elect pytest: /"
timeout 300 python bench.py --no-cpu-baseline --no-secondary --min-timed-steps 60 > $O/elect.json 2> $O/elect.err
python - <<PY
import json
d=json.load(open('$O/elect.json'))
print('elect bench', 'ms/step', round(d['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items()))
PY
