#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -4
for w in tictactoe_b8192_n50 breakout_b128_n50 connect4_b1024_n200; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench35_$w.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench35_$w.json').read().strip().splitlines()[-1])
print('$w', round(d['ms_per_step'],3), round(d['value']), 'kernel_ms', round(d['kernel_ms_per_step'],4), 'e2e', round(d['e2e']['value']))
ks=d['roofline'].get('kernel_split',{})
for k,v in ks.items(): print('     ',k, round(1000*v['ms']/v['launches'],2),'us x', v['launches'])
PY
done
