#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
MZ_BENCH_WATCHDOG=200 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 20 --warmup 5 --no-saturation --extras tictactoe_b8192_n50 \
    > gpurun_out/r2_11_bench_2gpu_$i.json 2> gpurun_out/r2_11_bench_2gpu_$i.err; echo "run $i rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_11_bench_2gpu_$i.err | tail -40 | cut -c1-200
cut -c1-200 gpurun_out/r2_11_bench_2gpu_$i.json
done
