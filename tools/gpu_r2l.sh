#!/bin/bash
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "layer_packs" 2>&1 | grep -E "^E  |passed|failed" | head -30
