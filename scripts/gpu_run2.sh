#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/r2_2_tests.log; cat gpurun_out/r2_2_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r2_2_smoke.log
timeout 900 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err; tail -5 gpurun_out/r2_2_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_2_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'loop',d.get('loop'))
PY
