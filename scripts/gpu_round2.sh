#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/tests2.log
for cfg in "16 64" "16 128" "8 32" "8 64" "16 32" "32 64"; do
  set -- $cfg
  MZ_FC_GROUP=$1 MZ_FC_THREADS=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_g$1_t$2.json 2> gpurun_out/bench2_g$1_t$2.err
done
for w in tictactoe_b8192_n50 connect4_b1024_n200 breakout_b128_n50; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$w.json 2> gpurun_out/bench2_$w.err
done
tail -25 gpurun_out/tests2.log
for f in gpurun_out/bench2_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d['value'], d['sims_per_sec'], d['kernel_ms_per_step'], d['e2e']['value'], d['gpu_launches'])" 2>/dev/null || tail -3 ${f%.json}.err; done
