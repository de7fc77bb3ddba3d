#!/bin/bash
# full check + profiles of the x3 tower
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -80 > gpurun_out/r2_4_tests.log; tail -40 gpurun_out/r2_4_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r2_4_smoke.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2_4_bench.json 2> gpurun_out/r2_4_bench.err; tail -5 gpurun_out/r2_4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_4_bench.json'))
print('cartpole value',d['value'],'e2e',d['e2e']['value'],'loop',d.get('loop',{}).get('value'), d.get('loop',{}).get('device_seconds'), d.get('loop',{}).get('seconds'))
for k,w in d.get('workloads',{}).items():
    print(k, 'value', w.get('value'), 'ms', (w.get('ms_per_search') or {}).get('median'), 'loop', (w.get('loop') or {}).get('value'), 'frac', (w.get('roofline') or {}).get('frac'), w.get('error'))
PY
# launch list of one Connect4 search (x3) and a full capture of the tower kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1300 --csv --log-file gpurun_out/r02_launches_connect4_x3.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop > gpurun_out/r2_4_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tower_x3 -s 40 -c 2 -o gpurun_out/r02_conv_tower_x3 \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-loop > gpurun_out/r2_4_ncu_full.log 2>&1
ls -la gpurun_out/r02_*
