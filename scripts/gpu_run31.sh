#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-saturation 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cartpole:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],4), 'ms kernel', round(d['kernel_ms_per_search'],4), 'e2e', round(d['e2e']['value']), 'loop', round(d['loop']['value']))"
for w in tictactoe_b8192_n50 connect4_b1024_n200; do
timeout 300 python bench.py --workload $w --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']))"
done
