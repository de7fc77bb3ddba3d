#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/conv_bench2.py > gpurun_out/conv_bench18.log 2>&1; cat gpurun_out/conv_bench18.log
