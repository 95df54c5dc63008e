#!/bin/bash
# Latency-regime sweep: rollout-kernel variants across rollout counts.
for n in 1920 2560 4096 8192 16384 32768 65536 131072; do for v in 2 9; do
  echo -n "N=$n v=$v: "; python tools/profile_step.py --rollouts $n --steps 30 --variant $v
done; done
python -m pytest tests/test_parity_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3
