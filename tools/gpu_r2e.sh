#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_reference_gpu.py -m gpu -q -rf -s -k "run_control_loop" > $OUT/pytest_r02e.log 2>&1; echo "pytest rc=$?"; grep -v "^GPUassert\|^$" $OUT/pytest_r02e.log | grep -v "^E  " | tail -30
