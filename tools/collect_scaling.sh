#!/bin/bash
# Strong scaling of BASELINE config 5 (projection on the fixed 32768^2 grid) on N GPUs of one box:
#   gpurun --gpus N -- tools/collect_scaling.sh N      -> gpurun_out/final/bench_project32768_Ngpu.json
# (N = 1 runs the plain single-process bench)
N=${1:-2}
O=gpurun_out/final; mkdir -p $O
if [ "$N" = "1" ]; then
  python bench.py --workload project32768 --steps 10 --warmup 3 > $O/bench_project32768_1gpu.json 2> $O/bench_project32768_1gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --workload project32768 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_project32768_${N}gpu.json 2> $O/bench_project32768_${N}gpu.err
fi
tail -1 $O/bench_project32768_${N}gpu.json | cut -c1-400
tail -2 $O/bench_project32768_${N}gpu.err | cut -c1-300
