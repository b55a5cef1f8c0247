#!/bin/bash
N=${1:-2}
O=gpurun_out/r02_scale; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > $O/multi_gpu_check_${N}gpu.log 2>&1; tail -3 $O/multi_gpu_check_${N}gpu.log; grep -c "mismatches=0" $O/multi_gpu_check_${N}gpu.log; grep "mismatches=[1-9]" $O/multi_gpu_check_${N}gpu.log | head -5
bash tools/r02_mg_probe.sh $N
