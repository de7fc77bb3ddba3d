#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_selfplay_gpu.py -q 2>&1 | tail -25
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
python -c "
import json
d=json.load(open('gpurun_out/bench12.json'))
print(d['value'], d['e2e']['value'], d.get('selfplay_loop'))" || tail -5 gpurun_out/bench12.err
timeout 600 python bench.py --workload connect4_b1024_n200 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench12_c4.json 2> gpurun_out/bench12_c4.err
python -c "
import json
d=json.load(open('gpurun_out/bench12_c4.json'))
print(d['value'], d['e2e']['value'], d.get('selfplay_loop'))" || tail -5 gpurun_out/bench12_c4.err
