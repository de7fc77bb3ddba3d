#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q -x 2>&1 | tail -12 > gpurun_out/tests9.log; cat gpurun_out/tests9.log
timeout 300 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench9_connect4.json 2> gpurun_out/bench9_connect4.err
python -c "
import json
d=json.load(open('gpurun_out/bench9_connect4.json'))
print(d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'), d['roofline'])" || tail -5 gpurun_out/bench9_connect4.err
