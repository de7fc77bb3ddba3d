#!/bin/bash
# 2-GPU call: NCCL path of bench.py (weak scaling), plus an ncu capture of the heads kernel on GPU 0
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench6_cartpole_2gpu.json 2> gpurun_out/bench6_cartpole_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload connect4_b1024_n200 --steps 3 --warmup 3 > gpurun_out/bench6_connect4_2gpu.json 2> gpurun_out/bench6_connect4_2gpu.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_cartpole_1gpu.json 2> gpurun_out/bench6_cartpole_1gpu.err
CUDA_VISIBLE_DEVICES=0 ncu --set full --clock-control none --import-source on -k regex:heads_kernel -s 40 -c 2 -f -o gpurun_out/prof_heads \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_heads.log 2>&1
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_tree_parity_gpu.py -q -k dirichlet 2>&1 | tail -3
for f in gpurun_out/bench6_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['n_gpus'], d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'))" 2>/dev/null || tail -5 ${f%.json}.err; done
