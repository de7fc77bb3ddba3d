#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tree_parity_gpu.py tests/test_selfplay_gpu.py tests/test_device_selfplay_gpu.py -m gpu -q -x 2>&1 | tail -5
for g in 0 1; do
MZ_FC_GENERIC=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-saturation 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic=$g cartpole:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],4), 'ms kernel', round(d['kernel_ms_per_search'],4), 'e2e', round(d['e2e']['value']), 'loop', round(d['loop']['value']))"
done
