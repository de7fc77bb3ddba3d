#!/bin/bash
echo "== solo issuer (MZ_TC_DEBUG_SKIP=32)"
MZ_NO_TC=0 MZ_TC_DEBUG_SKIP=32 timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -q -k "resident or tower_modes or graph_replay" 2>&1 | grep -E "passed|failed|FAILED" | head -12
echo "== two issuers"
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -q -k "resident or tower_modes or graph_replay" 2>&1 | grep -E "passed|failed|FAILED" | head -12
