#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 100 --csv --log-file gpurun_out/launches21_connect4.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch21.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tower -s 12 -c 2 -f -o gpurun_out/prof_conv_tower_final \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tower2.log 2>&1
ls -la gpurun_out/launches21_connect4.csv gpurun_out/prof_conv_tower_final.ncu-rep
