#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py -m gpu -q -x -k "partitioned or graph_replay or closed_loop" 2>&1 | tail -3
run() { # env, workload
env $1 timeout 300 python bench.py --workload $2 --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']))"
}
run MZ_PARTS=0 connect4_b1024_n200
run MZ_PARTS=4 connect4_b1024_n200
run MZ_PARTS=1 connect4_b1024_n200
run MZ_PARTS=0 tictactoe_b8192_n50
run MZ_NO_GRAPH=1 tictactoe_b8192_n50
run MZ_PARTS=0 breakout_b128_n50
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_search -s 4 -c 1 -o gpurun_out/r02_fc_search \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation > gpurun_out/r2_25_ncu_fc.log 2>&1; tail -2 gpurun_out/r2_25_ncu_fc.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
