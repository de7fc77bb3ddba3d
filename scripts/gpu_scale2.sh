#!/bin/bash
mkdir -p gpurun_out
for w in cartpole_b4096_n50 connect4_b1024_n200; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale2_$w.err > gpurun_out/scale2_$w.json
  tail -2 gpurun_out/scale2_$w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/scale2_$w.json').read().strip().splitlines()[-1])
print('$w n_gpus', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
done
