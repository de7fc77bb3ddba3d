#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_search -s 4 -c 1 -f -o gpurun_out/prof_fc_search_r34 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fc34.log 2>&1
ls -la gpurun_out/prof_fc_search_r34.ncu-rep
