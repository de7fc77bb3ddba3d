#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/tests3_conv.log
cat gpurun_out/tests3_conv.log
timeout 600 python -m pytest tests/test_resnet_gpu.py -q 2>&1 | tail -30 > gpurun_out/tests3_resnet.log
cat gpurun_out/tests3_resnet.log
python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_connect4_tc.json 2> gpurun_out/bench3_connect4_tc.err
MZ_NO_TC=1 python bench.py --workload connect4_b1024_n200 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench3_connect4_simt.json 2> gpurun_out/bench3_connect4_simt.err
for f in gpurun_out/bench3_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['sims_per_sec'], d['kernel_ms_per_step'], d['e2e']['value'], d['gpu_launches'])" 2>/dev/null || tail -3 ${f%.json}.err; done
