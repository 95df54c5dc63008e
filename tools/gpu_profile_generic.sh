#!/bin/bash
# Full ncu capture of the run-time layer kernel on the wider_deeper network at controller size (1920 x 100).
TAG=${1:-r02}
OUT=gpurun_out
P1="python tools/profile_step.py --rollouts 1920 --steps 4 --variant 11 --tag wider_deeper"
timeout 120 python tools/exp_generic.py gen
timeout 120 $P1 > $OUT/plain_gen_$TAG.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout_generic -s 2 -c 1 -f -o $OUT/prof_gen_$TAG $P1 > $OUT/ncu_gen_full_$TAG.log 2>&1
echo "full generic ($TAG) rc=$?"; cat $OUT/plain_gen_$TAG.log
