#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -15
for w in tictactoe_b8192_n50 breakout_b128_n50 connect4_b1024_n200; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench23_$w.err | tee gpurun_out/bench23_$w.json | cut -c1-300
  tail -3 gpurun_out/bench23_$w.err
done
