#!/bin/bash
TAG=${1:-r01_bf}
OUT=gpurun_out
P="python tools/profile_step.py --dynamics bf --rollouts 2560 --steps 6"
$P > $OUT/plain_bf_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 8 -c 4 -f -o $OUT/prof_bf_$TAG $P > $OUT/ncu_bf_$TAG.log 2>&1
echo "bf rc=$?"; cat $OUT/plain_bf_$TAG.log
