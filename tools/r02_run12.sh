#!/bin/bash
# round 2, GPU run 12: k_confine_fast variants (CTA height, CTAs per SM, early noise loads) + ncu of the default
set -x
O=gpurun_out/r02_run12; mkdir -p $O
run() { # name
  python - <<PY
import json
d=json.load(open('$O/$1.json'))
print('$1', 'ms/step', round(d['ms_per_step'],4), 'quiescent', round(d['quiescent']['ms_per_step'],4), ' '.join('%s=%.4f(%.2f)'%(k.replace('k_',''),v['ms_per_launch'],v['frac']) for k,v in d['roofline']['kernels'].items() if 'confine' in k or 'rbq' in k))
PY
}
for v in default cf_pre cf_ct8 cf_ct8pre cf_minb4 cf_pre4; do
  if [ $v = default ]; then unset FLUIDB200_LIB; else export FLUIDB200_LIB=$PWD/fluid_b200/variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-secondary --min-timed-steps 60 > $O/$v.json 2> $O/$v.err
  run $v
done
unset FLUIDB200_LIB
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_confine_fast" -s 308 -c 1 -o $O/r02_confine_fast -f \
    python bench.py --steps 3 --warmup 3 --preroll 300 --min-timed-steps 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1
tail -2 $O/ncu.log | cut -c1-200
for f in $O/*.err; do echo "== $f"; tail -n 3 $f | cut -c1-300; done
