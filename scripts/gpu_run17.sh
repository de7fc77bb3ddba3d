#!/bin/bash
timeout 1800 python -m pytest tests/test_tree_parity_gpu.py tests/test_resnet_gpu.py -m gpu -q --timeout 900 -k "gomoku" 2>&1 | tail -40
