#!/bin/bash
mkdir -p gpurun_out
./scripts/mma_rate 40 > gpurun_out/r2_19_mma_rate.md 2>&1; cat gpurun_out/r2_19_mma_rate.md
rm -f gpurun_out/x3_timeline.txt
MZ_NO_GRAPH=1 MZ_X3_TIMELINE=gpurun_out/x3_timeline.txt timeout 600 python scripts/x3_timeline.py > gpurun_out/r2_19_timeline.md 2>&1; cat gpurun_out/r2_19_timeline.md
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -q -k gomoku 2>&1 | tail -3
for g in 8 16 32; do for t in 64 128; do
  MZ_FC_GROUP=$g MZ_FC_THREADS=$t timeout 300 python bench.py --no-extras --no-loop --no-saturation --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fc_group $g threads $t:', round(d['value']), 'env-steps/s', d['ms_per_search']['median'], 'ms')"
done; done
