#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -q -k "connect4 or resident or tower_modes" 2>&1 | grep -E "passed|failed|FAILED|Error|assert|Mismatch|Max abs|mismatch" | head -40
