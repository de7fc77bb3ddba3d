#!/bin/bash
# round 2, run 1: the new fixtures / device self-play tests on the existing kernels, smoke, the restructured bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/r2_1_tests.log; cat gpurun_out/r2_1_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r2_1_smoke.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; tail -5 gpurun_out/r2_1_bench.err; cut -c1-1500 gpurun_out/r2_1_bench.json
