#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tree_step -s 150 -c 1 -f -o gpurun_out/prof_tree_step_c4 \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tree.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads -s 300 -c 2 -f -o gpurun_out/prof_heads_c4 \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_heads.log 2>&1
ls -la gpurun_out/*.ncu-rep
