#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_search -s 4 -c 1 -o gpurun_out/r02_fc_search_fixed \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation > gpurun_out/r2_33_ncu_fc.log 2>&1; tail -1 gpurun_out/r2_33_ncu_fc.log | cut -c1-120
timeout 300 python bench.py --workload connect4_b1024_n200 --no-extras --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4:', round(d['value']), 'env-steps/s', d['ms_per_search'], 'e2e', round(d['e2e']['value']), 'loop', (d.get('loop') or {}).get('value'))"
