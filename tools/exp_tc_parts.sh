#!/bin/bash
# Times the 1M-rollout tensor-core kernel with one part removed at a time (libraries from tools/build_tc_exp.sh).
L=autorally_b200/lib
cp $L/libmppi_b200.so /tmp/libmppi_b200.keep
echo -n "product: "; python tools/profile_step.py --rollouts 1048576 --steps 3 --variant 10
for n in ${@:-1 2 3 4 5}; do
  cp $L/exp/libmppi_b200_exp$n.so $L/libmppi_b200.so
  echo -n "exp $n: "; timeout 120 python tools/profile_step.py --rollouts 1048576 --steps 3 --variant 10
done
cp /tmp/libmppi_b200.keep $L/libmppi_b200.so
