#!/bin/bash
# Round-end style check: GPU tests, smoke, default bench (both arms), launch list of the default command.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/check_tests.log; cat gpurun_out/check_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; cut -c1-300 gpurun_out/check_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/check_bench_ref.json 2> gpurun_out/check_bench_ref.err; cut -c1-300 gpurun_out/check_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/check_launches_default.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-saturation > gpurun_out/check_ncu_default.log 2>&1
ls -la gpurun_out/*.csv
