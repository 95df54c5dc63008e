#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_reference_gpu.py tests/test_host_cpp.py tests/test_control_loop.py -m gpu -q -rf -s -k "feedback or run_control_loop or drop_in or bf or hot_swap or laps" > $OUT/pytest_r02d.log 2>&1; echo "pytest rc=$?"; grep -v "^GPUassert\|^$" $OUT/pytest_r02d.log | tail -60
timeout 300 python tools/exp_fused.py 2>&1 | grep "^bf" 
