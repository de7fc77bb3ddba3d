#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_bench.py 2> gpurun_out/conv_bench.log; cat gpurun_out/conv_bench.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/tests5.log; cat gpurun_out/tests5.log
for w in connect4_b1024_n200 tictactoe_b8192_n50 breakout_b128_n50; do
python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_$w.json 2> gpurun_out/bench5_$w.err
done
MZ_NO_GRAPH=1 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_connect4_nograph.json 2> gpurun_out/bench5_connect4_nograph.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_cartpole.json 2> gpurun_out/bench5_cartpole.err
for f in gpurun_out/bench5_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['sims_per_sec'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('gpu_launches'))" 2>/dev/null || tail -3 ${f%.json}.err; done
