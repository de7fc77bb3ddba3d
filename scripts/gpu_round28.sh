#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 400 python -m pytest tests/test_resnet_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench28.err | tee gpurun_out/bench28_connect4.json | cut -c1-200
tail -3 gpurun_out/bench28.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench28_connect4.json').read().strip().splitlines()[-1])
for k,v in d['roofline']['kernel_split'].items(): print('   ',k, round(v['ms'],3), v['launches'], round(1000*v['ms']/v['launches'],2),'us')
PY
