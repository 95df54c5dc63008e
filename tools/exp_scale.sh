#!/bin/bash
# Latency kernel: time vs horizon length for a nearly empty machine (64 rollouts = 32 warps).
for t in 16 32 64 100 200 400; do
  echo -n "N=64 T=$t v=9: "; python tools/profile_step.py --rollouts 64 --steps 30 --variant 9 --timesteps $t
done
