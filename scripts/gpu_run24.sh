#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "partitioned or graph_replay" 2>&1 | tail -5
for parts in 1 2 3 4; do
for w in connect4_b1024_n200 tictactoe_b8192_n50 breakout_b128_n50; do
MZ_PARTS=$parts timeout 300 python bench.py --workload $w --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parts=$parts $w:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']))"
done
done
