#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "fused_small_search or closed_loop or partitioned" 2>&1 | tail -15
run() { # env, workload
env $1 timeout 300 python bench.py --workload $2 --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']), d['roofline'].get('kernel_split'))"
}
run MZ_SMALL_SEARCH=1 tictactoe_b8192_n50
run MZ_SMALL_SEARCH=0 tictactoe_b8192_n50
run MZ_SMALL_SEARCH=1 breakout_b128_n50
run MZ_SMALL_SEARCH=0 breakout_b128_n50
