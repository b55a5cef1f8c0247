#!/bin/bash
# Build experiment variants of libfluidb200.so with extra -D flags into fluid_b200/variants/
# (git-ignored, travels to the GPU box).  usage: fluid_b200/variants.sh name "-DFOO=1" [name "-D..."]...
set -e
cd "$(dirname "$0")/.."
mkdir -p fluid_b200/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC $flags \
       -shared -o fluid_b200/variants/lib_$name.so fluid_b200/csrc/fluidb200.cu &
done
wait
ls -la fluid_b200/variants
