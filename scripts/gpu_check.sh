#!/bin/bash
# Full round-end style check: GPU tests, smoke, default bench (both arms), Connect4 bench, Connect4 launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/check_tests.log; cat gpurun_out/check_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; cat gpurun_out/check_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/check_bench_ref.json 2> gpurun_out/check_bench_ref.err; cat gpurun_out/check_bench_ref.json | cut -c1-300
timeout 300 python bench.py --workload connect4_b1024_n200 --steps 3 --warmup 3 > gpurun_out/check_bench_c4.json 2> gpurun_out/check_bench_c4.err; cat gpurun_out/check_bench_c4.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 100 --csv --log-file gpurun_out/check_launches_connect4.csv \
    python bench.py --workload connect4_b1024_n200 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/check_ncu_launch_connect4.log 2>&1
ls -la gpurun_out/check_launches_connect4.csv
