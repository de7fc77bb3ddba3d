#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests -m gpu -q -x -k "rank or world or multi" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 600 gpurun_out/bench_2gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'loop', d.get('loop',{}).get('value'), 'n_gpus', d['n_gpus'])
for k,w in d.get('workloads',{}).items(): print(k, round(w['value']), round(w['e2e']['value']), w['ms_per_search']['median'], (w.get('loop') or {}).get('value'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m muzero_general_b200.parallel --game connect4 --games 2048 --reports 2 --moves-per-report 4 2>&1 | tail -4
