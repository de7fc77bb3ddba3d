#!/bin/bash
for flag in 32 96 0 64; do
echo "== MZ_TC_DEBUG_SKIP=$flag (32 = solo issuer, 64 = issuer 0 is warp 2)"
MZ_NO_TC=0 MZ_TC_DEBUG_SKIP=$flag timeout 300 python -m pytest tests/test_resnet_gpu.py -m gpu -q -k "resident" 2>&1 | grep -E "passed|failed" | head -3
done
