#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tree_parity_gpu.py -q 2>&1 | tail -5
for g in 16 8; do
MZ_FC_GROUP=$g timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench17_g$g.json 2> gpurun_out/bench17_g$g.err
python -c "
import json
d=json.load(open('gpurun_out/bench17_g$g.json'))
print($g, d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('selfplay_loop',{}).get('value'))" || tail -5 gpurun_out/bench17_g$g.err
done
