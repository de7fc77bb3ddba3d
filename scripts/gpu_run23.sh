#!/bin/bash
mkdir -p gpurun_out
./scripts/mma_rate 40 > gpurun_out/r2_23_mma_rate.md 2>&1; cat gpurun_out/r2_23_mma_rate.md
rm -f gpurun_out/x3_timeline.txt
MZ_NO_GRAPH=1 MZ_X3_TIMELINE=gpurun_out/x3_timeline.txt timeout 600 python scripts/x3_timeline.py > gpurun_out/r2_23_timeline.md 2>&1; cat gpurun_out/r2_23_timeline.md
for w in connect4_b1024_n200 tictactoe_b8192_n50; do
timeout 300 python bench.py --workload $w --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w:', round(d['value']), 'env-steps/s', d['ms_per_search']['median'], 'ms', d['roofline'].get('kernel_split'))"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-loop --no-saturation 2>/dev/null | cut -c1-300
