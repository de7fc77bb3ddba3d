#!/bin/bash
mkdir -p gpurun_out
for skip in 0 1 2 4 8 6 12 14 15; do
  MZ_TC_DEBUG_SKIP=$skip timeout 200 python bench.py --workload connect4_b1024_n200 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null > gpurun_out/bench27_skip$skip.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench27_skip$skip.json').read().strip().splitlines()[-1])
v=d['roofline']['kernel_split']['conv_tower_tc_kernel']
print('skip=$skip tower avg us', round(1000*v['ms']/v['launches'],2), 'step ms', round(d['ms_per_step'],2))
PY
done
for g in 4 8 16; do for t in 32 64 128; do
  echo "cartpole G=$g T=$t: $(MZ_FC_GROUP=$g MZ_FC_THREADS=$t timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['value']))")"
done; done
