#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q -k "graph_replay" 2>&1 | tail -3
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:tree_step -s 150 -c 1 -f -o gpurun_out/prof_tree_step_warm \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tree_w.log 2>&1
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:heads -s 300 -c 2 -f -o gpurun_out/prof_heads_warm \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_heads_w.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tower -s 12 -c 2 -f -o gpurun_out/prof_tower_resident \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tower_r.log 2>&1
ls -la gpurun_out/*.ncu-rep
