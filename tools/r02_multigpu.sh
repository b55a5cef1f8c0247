#!/bin/bash
# round 2: the driver's SCALE invocation on N GPUs of one box (weak scaling, Karman 4096^2 per GPU + config[3] jet 16384^2 per
# GPU inside the same line), the N-rank vs 1-GPU parity check it carries, and the multi-process slab check.
#   gpurun --gpus N -- tools/r02_multigpu.sh N      -> gpurun_out/r02_scale/
N=${1:-2}
O=gpurun_out/r02_scale; mkdir -p $O
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > $O/multi_gpu_check_${N}gpu.log 2>&1; tail -5 $O/multi_gpu_check_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_karman4096_${N}gpu.json 2> $O/bench_karman4096_${N}gpu.err
tail -3 $O/bench_karman4096_${N}gpu.err | cut -c1-400
python - <<PY
import json
d=json.loads(open('$O/bench_karman4096_${N}gpu.json').read().strip().splitlines()[-1])
print('N=$N ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'full', d['e2e'] and d['e2e']['full_field_loop']['value'])
print('parity_check', d.get('parity_check'))
c3=d.get('config3_jet16384'); print('jet16384', c3 and (c3['ms_per_step'], c3['value']))
PY
# A/B: the same run with the exchange in front of every step (round 1's placement)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-overlap --no-secondary > $O/bench_karman4096_${N}gpu_nooverlap.json 2> $O/bench_karman4096_${N}gpu_nooverlap.err
python -c "import json; d=json.loads(open('$O/bench_karman4096_${N}gpu_nooverlap.json').read().strip().splitlines()[-1]); print('N=$N no-overlap ms/step', d['ms_per_step'])"
if [ "$N" = "2" ]; then
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_karman4096_1gpu.json 2> $O/bench_karman4096_1gpu.err
  python -c "import json; d=json.load(open('$O/bench_karman4096_1gpu.json')); print('N=1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['full_field_loop']['value'], 'jet16384', d['config3_jet16384']['ms_per_step'])"
fi
