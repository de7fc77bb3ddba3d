#!/bin/bash
python scripts/debug_ttt_loop.py 2>&1 | tail -15
timeout 900 python -m pytest tests/test_device_selfplay_gpu.py -m gpu -q -k "selfplay_api" --timeout 600 2>&1 | tail -5
