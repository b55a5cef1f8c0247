#!/bin/bash
# round 2, GPU run 21: tile shapes of the advection kernels after the 2-D TMA change (threads per CTA, lines per tile, CTAs per SM)
O=gpurun_out/r02_run21; mkdir -p $O
run() { python - <<PY
import json
d=json.load(open('$O/$1.json'))
print('$1', 'ms/step', round(d['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items() if 'tile' in k))
PY
}
for v in a128 a128m6 a128m8 a128t16; do
  if [ $v = default ]; then unset FLUIDB200_LIB; else export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-secondary --min-timed-steps 60 > $O/$v.json 2> $O/$v.err || tail -2 $O/$v.err
  run $v
done
