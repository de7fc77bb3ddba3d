#!/bin/bash
# 2 GPUs: bench under torchrun (NCCL), the multi-GPU self-play entry point, the world-size-invariance test with NCCL
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-saturation \
    > gpurun_out/r2_9_bench_2gpu.json 2> gpurun_out/r2_9_bench_2gpu.err; tail -5 gpurun_out/r2_9_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_9_bench_2gpu.json'))
print('2gpu cartpole value',d['value'],'e2e',d['e2e']['value'],'loop',(d.get('loop') or {}).get('value'))
for k,w in d.get('workloads',{}).items():
    print(k, 'value', w.get('value'), 'loop', (w.get('loop') or {}).get('value'), w.get('error'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m muzero_general_b200.parallel --game connect4 --games 2048 --reports 3 --moves-per-report 8 2>&1 | tail -6 | cut -c1-400
timeout 600 python -m pytest tests/test_device_selfplay_gpu.py -m gpu -q -k "multi_rank" --timeout 600 2>&1 | tail -5
