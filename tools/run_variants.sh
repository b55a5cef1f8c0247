#!/bin/bash
# usage: XS="0 7" tools/run_variants.sh name...   -- ablation timing of each variant library on the GPU box
for v in "$@"; do
  for x in ${XS:-0 7}; do
    echo -n "variant $v: "
    FLUIDB200_LIB=$PWD/tools/variants/lib_$v.so FLUIDB200_RBQ_X=$x timeout 120 python tools/dbg_rbq.py 2>&1 | tail -1
  done
done
