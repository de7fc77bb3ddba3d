#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
for w in connect4_b1024_n200 tictactoe_b8192_n50 breakout_b128_n50; do
timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_$w.json 2> gpurun_out/bench16_$w.err
python -c "
import json
d=json.load(open('gpurun_out/bench16_$w.json'))
print('$w', d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'), d['roofline']['frac'], d.get('selfplay_loop',{}).get('value'))" || tail -5 gpurun_out/bench16_$w.err
done
