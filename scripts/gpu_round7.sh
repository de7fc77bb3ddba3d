#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_bench.py 2> gpurun_out/conv_bench7.log; cat gpurun_out/conv_bench7.log
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q 2>&1 | tail -8 > gpurun_out/tests7.log; cat gpurun_out/tests7.log
for w in connect4_b1024_n200 tictactoe_b8192_n50 breakout_b128_n50; do
python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench7_$w.json 2> gpurun_out/bench7_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 160 --csv --log-file gpurun_out/launches7_connect4.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch7.log 2>&1
for f in gpurun_out/bench7_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'))" 2>/dev/null || tail -3 ${f%.json}.err; done
