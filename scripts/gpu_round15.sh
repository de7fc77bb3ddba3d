#!/bin/bash
mkdir -p gpurun_out
for c in 2 1; do
MZ_TC_VERBOSE=1 MZ_TC_CTAS=$c timeout 300 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench15_connect4_$c.json 2> gpurun_out/bench15_connect4_$c.err
grep conv_tower gpurun_out/bench15_connect4_$c.err | head -2
python -c "
import json
d=json.load(open('gpurun_out/bench15_connect4_$c.json'))
print($c, d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'), d['roofline']['frac'])" || tail -5 gpurun_out/bench15_connect4_$c.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"heads_kernel|tree_step" -s 10 -c 4 -f -o gpurun_out/prof_heads_tree \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_heads_tree.log 2>&1
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q 2>&1 | tail -3
