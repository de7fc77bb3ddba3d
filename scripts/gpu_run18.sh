#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/r2_18_tests.log; tail -12 gpurun_out/r2_18_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
SECONDS=0
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_18_bench.json 2> gpurun_out/r2_18_bench.err; echo "bench default: ${SECONDS}s"; grep "^\[bench" gpurun_out/r2_18_bench.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_18_bench.json').read().strip().splitlines()[-1])
print('cartpole value',d['value'],'e2e',d['e2e']['value'],'loop',(d.get('loop') or {}).get('value'),'cpu',(d.get('cpu_baseline') or {}).get('value'))
for k,w in d.get('workloads',{}).items():
    print(k, 'value', w.get('value'), 'loop', (w.get('loop') or {}).get('value'), w.get('error'))
PY
