#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload connect4_b1024_n200 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench36.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench36.json').read().strip().splitlines()[-1])
print('connect4', round(d['ms_per_step'],3), round(d['value']), 'kernel_ms', round(d['kernel_ms_per_step'],4))
for k,v in d['roofline']['kernel_split'].items(): print('     ',k, round(1000*v['ms']/v['launches'],2),'us x', v['launches'])
PY
timeout 300 python scripts/conv_bench.py 2>&1 | tail -4
