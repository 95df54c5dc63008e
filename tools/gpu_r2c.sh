#!/bin/bash
# diagnostic captures: the 131072-rollout tensor-core launch (one GPU's share of 1M over 8), 1M fused, BF 1M
set -u
OUT=gpurun_out; mkdir -p $OUT
P="python tools/profile_step.py --rollouts 131072 --steps 2"
$P > $OUT/plain_131k_r02c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -f -o $OUT/prof_131k_r02c $P > $OUT/ncu_131k_r02c.log 2>&1
echo "131k rc=$?"; cat $OUT/plain_131k_r02c.log
P="python tools/profile_step.py --rollouts 1048576 --steps 2"
$P > $OUT/plain_1m_r02c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_tc -s 1 -c 1 -f -o $OUT/prof_1m_r02c $P > $OUT/ncu_1m_r02c.log 2>&1
echo "1m rc=$?"; cat $OUT/plain_1m_r02c.log
P="python tools/profile_step.py --rollouts 1048576 --steps 2 --dynamics bf"
$P > $OUT/plain_bf1m_r02c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o $OUT/prof_bf1m_r02c $P > $OUT/ncu_bf1m_r02c.log 2>&1
echo "bf rc=$?"; cat $OUT/plain_bf1m_r02c.log
ls -la $OUT/*.ncu-rep
