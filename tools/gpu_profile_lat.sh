#!/bin/bash
# Full ncu capture of one 1920 x 100 step for a given rollout variant (default: AUTO).
TAG=${1:-r01c}; V=${2:-0}
OUT=gpurun_out
P1="python tools/profile_step.py --rollouts 1920 --steps 6 --variant $V"
$P1 > $OUT/plain_1920_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 8 -c 4 -f -o $OUT/prof_1920_$TAG $P1 > $OUT/ncu_1920_full_$TAG.log 2>&1
echo "full 1920 ($TAG, variant $V) rc=$?"; cat $OUT/plain_1920_$TAG.log
