#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -s -x --timeout 300 2>&1 | tail -80 > gpurun_out/r2_3_conv.log; tail -40 gpurun_out/r2_3_conv.log
grep -q "failed" gpurun_out/r2_3_conv.log && { echo "conv tests failed: stopping"; exit 1; }
timeout 1800 python -m pytest tests/test_resnet_gpu.py tests/test_selfplay_gpu.py tests/test_device_selfplay_gpu.py -m gpu -q --timeout 600 2>&1 | tail -150 > gpurun_out/r2_3_tests.log; tail -70 gpurun_out/r2_3_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-loop --workload connect4_b1024_n200 --extras connect4_b1024_n200@fp16 > gpurun_out/r2_3_bench_c4.json 2> gpurun_out/r2_3_bench_c4.err; tail -3 gpurun_out/r2_3_bench_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_3_bench_c4.json'))
print('x3 value',d['value'],'ms/search',d['ms_per_search']['median'],d['dtype'][:40]); print(d['roofline']['kernel_split'])
w=d['workloads']['connect4_b1024_n200@fp16']; print('fp16 value',w['value'],w['ms_per_search']['median'])
PY
