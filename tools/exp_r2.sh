#!/bin/bash
# Experiment: CTA shape / register cap of the two-rollouts-per-thread kernel at 1M rollouts.
for cfg in 0 1 2 3 4 5; do
  echo -n "R2 cfg $cfg: "; MPPI_R2_CONFIG=$cfg python tools/profile_step.py --rollouts 1048576 --steps 3 --variant 2
done
echo -n "R1: "; python tools/profile_step.py --rollouts 1048576 --steps 3 --variant 1
python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
