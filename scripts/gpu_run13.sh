#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/r2_13_tests.log; tail -12 gpurun_out/r2_13_tests.log
SECONDS=0
timeout 1500 python bench.py > gpurun_out/r2_13_bench.json 2> gpurun_out/r2_13_bench.err; echo "bench default: ${SECONDS}s"; grep "^\[bench" gpurun_out/r2_13_bench.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_13_bench.json'))
print('cartpole value',d['value'],'e2e',d['e2e']['value'],'loop',(d.get('loop') or {}).get('value'),'cpu',d.get('cpu_baseline'))
for k,w in d.get('workloads',{}).items():
    print(k, 'value', w.get('value'), 'loop', (w.get('loop') or {}).get('value'), w.get('error'))
PY
SECONDS=0
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_13_bench_ref.json 2> gpurun_out/r2_13_bench_ref.err; echo "bench reference arm: ${SECONDS}s"; cut -c1-300 gpurun_out/r2_13_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 520 --csv --log-file gpurun_out/r02_launches_ttt_loop.csv \
    python -m muzero_general_b200.parallel --game tictactoe --games 8192 --simulations 50 --reports 1 --moves-per-report 12 > gpurun_out/r2_13_ncu_ttt.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_launches_ttt_loop.csv
