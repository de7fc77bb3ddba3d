#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -5
for w in connect4_b1024_n200 tictactoe_b8192_n50 breakout_b128_n50; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench22_$w.json | cut -c1-400
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/launches22_ttt.csv \
    python bench.py --workload tictactoe_b8192_n50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch22_ttt.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/launches22_breakout.csv \
    python bench.py --workload breakout_b128_n50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch22_breakout.log 2>&1
ls -la gpurun_out/launches22_*.csv
