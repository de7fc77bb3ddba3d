#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench26.err | tee gpurun_out/bench26_connect4.json | cut -c1-300
tail -3 gpurun_out/bench26.err
timeout 300 python scripts/conv_bench2.py 2>&1 | tail -8
