#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:selfplay_step -s 5 -c 40 --csv --log-file gpurun_out/r02_step_ttt.csv \
    python -m muzero_general_b200.parallel --game tictactoe --games 8192 --reports 2 --moves-per-report 30 > gpurun_out/r2_7_ncu_ttt.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_step_ttt.csv
grep -o '"[0-9.]*","[a-z]*"$' gpurun_out/r02_step_ttt.csv | head -45 | tr '\n' ' '
timeout 900 python -m pytest tests/test_selfplay_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15
