#!/bin/bash
MPPI_HALF_MODE=1 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "variant or ragged or spread" 2>&1 | tail -2
for n in 1920 4096 16384; do for m in 0 1 2; do
  echo -n "N=$n mode $m: "; MPPI_HALF_MODE=$m python tools/profile_step.py --rollouts $n --steps 50 --variant 9
done; done
