#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_resnet_gpu.py tests/test_conv_gpu.py tests/test_device_selfplay_gpu.py -m gpu -q --timeout 900 2>&1 | tail -70 > gpurun_out/r2_6_tests.log; tail -40 gpurun_out/r2_6_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-saturation --extras tictactoe_b8192_n50 > gpurun_out/r2_6_bench.json 2> gpurun_out/r2_6_bench.err; tail -5 gpurun_out/r2_6_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_6_bench.json'))
print('cartpole value',d['value'],'loop',d.get('loop'))
for k,w in d.get('workloads',{}).items():
    print(k, 'value', w.get('value'), 'loop', w.get('loop'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/r02_launches_cartpole_loop.csv \
    python -m muzero_general_b200.parallel --game cartpole --games 4096 --reports 2 --moves-per-report 60 > gpurun_out/r2_6_ncu_loop.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_launches_cartpole_loop.csv
