#!/bin/bash
# Full ncu capture of the filled-GPU rollout kernel (1M rollouts), after a plain run of the same command.
TAG=${1:-r01b}
OUT=gpurun_out
P2="python tools/profile_step.py --rollouts 1048576 --steps 2"
$P2 > $OUT/plain_1m_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout -s 1 -c 1 -f -o $OUT/prof_1m_$TAG $P2 > $OUT/ncu_1m_full_$TAG.log 2>&1
echo "full 1m rc=$?"; cat $OUT/plain_1m_$TAG.log
