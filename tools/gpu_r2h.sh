#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
(python tools/exp_overlap.py; MPPI_NO_SPLIT_FINALIZE=1 python tools/exp_overlap.py) > $OUT/exp_overlap_r02h.txt 2>&1; sort -k3,3 -k4,4 -s $OUT/exp_overlap_r02h.txt
python tools/profile_step.py --rollouts 131072 --steps 3
python tools/profile_step.py --rollouts 1920 --steps 20 --variant 11
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "layer or generic or fused" 2>&1 | tail -3
