#!/bin/bash
mkdir -p gpurun_out
for bo in 1 0; do
echo "== base_offset=$bo"
MZ_TC_BASE_OFFSET=$bo timeout 300 python -m pytest tests/test_conv_gpu.py -q -x -k "tensor_core" 2>&1 | tail -4
done
timeout 300 python scripts/conv_bench.py 2>&1 | grep tcgen05
