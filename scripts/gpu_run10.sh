#!/bin/bash
mkdir -p gpurun_out
MZ_BENCH_WATCHDOG=240 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-saturation \
    > gpurun_out/r2_10_bench_2gpu.json 2> gpurun_out/r2_10_bench_2gpu.err; grep -v "^\*\|OMP_NUM" gpurun_out/r2_10_bench_2gpu.err | tail -60 | cut -c1-220
cut -c1-600 gpurun_out/r2_10_bench_2gpu.json
