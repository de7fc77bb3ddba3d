#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py -q -s 2>&1 | grep -E "err|passed|failed|FAILED|Error|assert|TV" | tail -60 > gpurun_out/tests4_resnet.log
cat gpurun_out/tests4_resnet.log
python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench4_connect4_tc.json 2> gpurun_out/bench4_connect4_tc.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 400 --csv --log-file gpurun_out/launches_connect4.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 100 -c 2 -f -o gpurun_out/prof_conv_tc \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1
python bench.py --workload connect4_b1024_n200 --impl reference --steps 1 --warmup 0 > gpurun_out/bench4_connect4_ref.json 2> gpurun_out/bench4_connect4_ref.err
for f in gpurun_out/bench4_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'))" 2>/dev/null || tail -3 ${f%.json}.err; done
