#!/bin/bash
# final-build evidence: launch lists of the bench command (default workload and Connect4), full capture of the fused FC kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches39_cartpole.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch39_cartpole.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 100 --csv --log-file gpurun_out/launches39_connect4.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch39_connect4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_search -s 4 -c 1 -f -o gpurun_out/prof_fc_search_final \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fc39.log 2>&1
ls -la gpurun_out/launches39_*.csv gpurun_out/prof_fc_search_final.ncu-rep
