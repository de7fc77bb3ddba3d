#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tree_parity_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in connect4_b1024_n200 breakout_b128_n50; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench42_$w.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench42_$w.json').read().strip().splitlines()[-1])
print('$w', round(d['ms_per_step'],3), round(d['value']), 'kernel_ms', round(d['kernel_ms_per_step'],4))
for k,v in d['roofline'].get('kernel_split',{}).items(): print('     ',k, round(1000*v['ms']/v['launches'],2),'us x', v['launches'])
PY
done
