#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
run() { # env, workload
env $1 timeout 300 python bench.py --workload $2 --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['roofline'].get('kernel_split',{}); print('$1 $2:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']), {k:round(v['ms'],3) for k,v in ks.items()})"
}
run MZ_PARTS=0 tictactoe_b8192_n50
run MZ_PARTS=0 breakout_b128_n50
