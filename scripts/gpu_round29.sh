#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -3
for skip in 0 4 1 15; do
  MZ_TC_DEBUG_SKIP=$skip timeout 200 python bench.py --workload connect4_b1024_n200 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null > gpurun_out/bench29_skip$skip.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench29_skip$skip.json').read().strip().splitlines()[-1])
v=d['roofline']['kernel_split']['conv_tower_tc_kernel']
print('skip=$skip tower avg us', round(1000*v['ms']/v['launches'],2), 'step ms', round(d['ms_per_step'],2))
PY
done
