#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_search -s 3 -c 1 -f -o gpurun_out/prof_fc_search_r2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fc2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tower -s 12 -c 2 -f -o gpurun_out/prof_conv_tower \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tower.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 100 --csv --log-file gpurun_out/launches10_connect4.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch10.log 2>&1
ls -la gpurun_out/*.ncu-rep
