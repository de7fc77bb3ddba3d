#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_search -c 1 -o gpurun_out/r02_small_search_ttt2 \
    python bench.py --workload tictactoe_b8192_n50 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation > gpurun_out/r2_29_ncu.log 2>&1; tail -2 gpurun_out/r2_29_ncu.log | cut -c1-200
timeout 300 python bench.py --workload tictactoe_b8192_n50 --no-extras --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ttt:', round(d['value']), 'env-steps/s', round(d['ms_per_search']['median'],3), 'ms e2e', round(d['e2e']['value']), 'loop', d.get('loop'))"
