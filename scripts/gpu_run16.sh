#!/bin/bash
timeout 900 python -m pytest tests/test_selfplay_gpu.py tests/test_tree_parity_gpu.py -m gpu -q --timeout 600 2>&1 | tail -4
for v in 0 1; do
MZ_TREE_LATENCY=$v timeout 300 python bench.py --workload tictactoe_b8192_n50 --no-cpu-baseline --no-extras > gpurun_out/r2_16_ttt_$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_16_ttt_$v.json').read().strip().splitlines()[-1])
print('MZ_TREE_LATENCY=$v', d['value'], d['ms_per_search']['median'], {k:round(x['ms'],2) for k,x in d['roofline']['kernel_split'].items()}, 'loop', d['loop']['value'], d['loop'].get('parked_events'))
PY
done
