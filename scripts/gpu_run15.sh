#!/bin/bash
for v in 0 1; do
MZ_TREE_LATENCY=$v timeout 300 python bench.py --workload tictactoe_b8192_n50 --no-cpu-baseline --no-extras --no-loop > gpurun_out/r2_15_ttt_$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_15_ttt_$v.json').read().strip().splitlines()[-1])
print('MZ_TREE_LATENCY=$v', d['value'], d['ms_per_search']['median'], {k:round(x['ms'],2) for k,x in d['roofline']['kernel_split'].items()})
PY
done
for v in 0 1; do
MZ_TREE_LATENCY=$v timeout 300 python bench.py --workload connect4_b1024_n200 --no-cpu-baseline --no-extras --no-loop > gpurun_out/r2_15_c4_$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_15_c4_$v.json').read().strip().splitlines()[-1])
print('c4 MZ_TREE_LATENCY=$v', d['value'], d['ms_per_search']['median'], {k:round(x['ms'],2) for k,x in d['roofline']['kernel_split'].items()})
PY
done
