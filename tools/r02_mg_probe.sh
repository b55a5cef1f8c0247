#!/bin/bash
# round 2: where do the ~0.05 ms per step of the 2-GPU run go?  (halo profiling slot, host enqueue time, overlap A/B)
N=${1:-2}
O=gpurun_out/r02_scale; mkdir -p $O
for mode in overlap nooverlap; do
  flag=""; [ $mode = nooverlap ] && flag="--no-overlap"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-secondary $flag > $O/probe_${N}gpu_$mode.json 2> $O/probe_${N}gpu_$mode.err
  python - <<PY
import json
d=json.loads(open('$O/probe_${N}gpu_$mode.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$mode N=$N ms/step', round(d['ms_per_step'],4), 'quiescent', round(d['quiescent']['ms_per_step'],4), 'host enqueue', d.get('host_enqueue_ms_per_step'), 'prof block', round(r['profiled_block_ms_per_step'],4))
print('   phases', {k:round(v,4) for k,v in r['phases_ms_per_step'].items()}, 'sum', round(sum(r['phases_ms_per_step'].values()),4))
PY
  tail -2 $O/probe_${N}gpu_$mode.err | cut -c1-300
done
