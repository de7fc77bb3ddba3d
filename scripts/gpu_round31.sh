#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do
  echo "== MZ_NO_PDL=$pdl"
  MZ_NO_PDL=$pdl timeout 200 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench31_$pdl.err > gpurun_out/bench31_nopdl$pdl.json
  tail -2 gpurun_out/bench31_$pdl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench31_nopdl$pdl.json').read().strip().splitlines()[-1])
v=d['roofline']['kernel_split']['conv_tower_tc_kernel']
print('no_pdl=$pdl tower avg us', round(1000*v['ms']/v['launches'],2), 'step ms', round(d['ms_per_step'],2), 'value', round(d['value']))
PY
done
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py tests/test_selfplay_gpu.py tests/test_tree_parity_gpu.py -m gpu -x -q 2>&1 | tail -8
for w in tictactoe_b8192_n50 breakout_b128_n50 cartpole_b4096_n50; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],3), round(d['value']), 'kernel_ms', round(d['kernel_ms_per_step'],3))"
done
