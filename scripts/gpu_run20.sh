#!/bin/bash
mkdir -p gpurun_out
./scripts/mma_rate 40 > gpurun_out/r2_20_mma_rate.md 2>&1; cat gpurun_out/r2_20_mma_rate.md
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:heads_kernel|tree_step_kernel' -s 60 -c 5 -o gpurun_out/r02_connect4_heads_tree \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation > gpurun_out/r2_20_ncu_c4.log 2>&1; tail -2 gpurun_out/r2_20_ncu_c4.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:small_tower|heads_kernel|tree_step_kernel' -s 100 -c 5 -o gpurun_out/r02_tictactoe_step \
    python bench.py --workload tictactoe_b8192_n50 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation > gpurun_out/r2_20_ncu_ttt.log 2>&1; tail -2 gpurun_out/r2_20_ncu_ttt.log
ls -la gpurun_out/*.ncu-rep
