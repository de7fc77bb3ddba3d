#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/conv_bench.py 2>&1 | grep tcgen05
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py tests/test_selfplay_gpu.py -q 2>&1 | tail -4
timeout 300 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench20_connect4.json 2> gpurun_out/bench20_connect4.err
python -c "
import json
d=json.load(open('gpurun_out/bench20_connect4.json'))
print(d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'), d['roofline']['frac'])" || tail -5 gpurun_out/bench20_connect4.err
