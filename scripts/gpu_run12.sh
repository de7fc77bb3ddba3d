#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 520 --csv --log-file gpurun_out/r02_launches_ttt_loop.csv \
    python -m muzero_general_b200.parallel --game tictactoe --games 8192 --reports 1 --moves-per-report 16 > gpurun_out/r2_12_ncu_ttt.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_launches_ttt_loop.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 520 --csv --log-file gpurun_out/r02_launches_ttt_search.csv \
    python bench.py --workload tictactoe_b8192_n50 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop > gpurun_out/r2_12_ncu_ttt2.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_launches_ttt_search.csv
