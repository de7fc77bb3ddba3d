#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tree_parity_gpu.py tests/test_resnet_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -8
for w in connect4_b1024_n200 tictactoe_b8192_n50 breakout_b128_n50; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench25_$w.err | tee gpurun_out/bench25_$w.json | cut -c1-300
  tail -3 gpurun_out/bench25_$w.err
done
