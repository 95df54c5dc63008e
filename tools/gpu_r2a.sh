#!/bin/bash
# round 2, GPU call A: the whole GPU suite (all failures, not -x) + a quick look at fused vs sampler at 1M
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -rf --durations=15 > $OUT/pytest_r02a.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_r02a.log
timeout 300 python tools/exp_fused.py > $OUT/exp_fused_r02a.txt 2>&1; echo "exp rc=$?"; cat $OUT/exp_fused_r02a.txt
