#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
