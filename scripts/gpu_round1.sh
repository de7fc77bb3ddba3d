#!/bin/bash
# One gpurun call: GPU tests, smoke, bench sweep over the lane-group width, ncu launch list + one full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/tests.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
for g in 8 16 32; do
  MZ_FC_GROUP=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fc_search -s 3 -c 2 -f -o gpurun_out/prof_fc_search \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
nproc > gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/nproc.txt
tail -5 gpurun_out/tests.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench*.json
