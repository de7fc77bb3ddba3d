#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_device_selfplay_gpu.py tests/test_selfplay_gpu.py -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/r2_8_tests.log; tail -15 gpurun_out/r2_8_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-saturation --extras tictactoe_b8192_n50 > gpurun_out/r2_8_bench.json 2> gpurun_out/r2_8_bench.err; tail -5 gpurun_out/r2_8_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_8_bench.json'))
print('cartpole value',d['value'],'loop',d.get('loop'))
for k,w in d.get('workloads',{}).items():
    print(k, 'value', w.get('value'), 'loop', w.get('loop'))
PY
