#!/bin/bash
for n in 1048576; do for cfg in 13 20 21 22; do
  echo -n "N=$n R2 cfg $cfg: "; MPPI_R2_CONFIG=$cfg python tools/profile_step.py --rollouts $n --steps 3 --variant 2
done; done
