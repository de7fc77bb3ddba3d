#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "fused_small_search" 2>&1 | tail -5
run() { # env, workload
env $1 timeout 300 python bench.py --workload $2 --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ks=d['roofline'].get('kernel_split',{}); print('$1 $2:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']), {k:round(v['ms'],3) for k,v in ks.items()})"
}
run MZ_SMALL_SEARCH=1 tictactoe_b8192_n50
run MZ_SMALL_SEARCH=1 breakout_b128_n50
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_search -c 1 -o gpurun_out/r02_small_search_ttt \
    env MZ_SMALL_SEARCH=1 python bench.py --workload tictactoe_b8192_n50 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation > gpurun_out/r2_27_ncu.log 2>&1; tail -2 gpurun_out/r2_27_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
