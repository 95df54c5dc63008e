#!/bin/bash
# Builds libmppi_b200 variants of the tensor-core kernel's occupancy trade-off: name:flags pairs, flags from
# -DTC_MINCTAS=n (resident tiles per SM) and -DTC_PAD=1 (pad shared memory to exactly that residency).  The two-chunks-in-flight
# variant (TC_DUAL) measured in profiles/exp_tc_cfg_r01.txt was removed after the measurement.
set -e
cd "$(dirname "$0")/.."
L=autorally_b200/lib; mkdir -p $L/exp
for cfg in "$@"; do
  name=${cfg%%:*}; flags=${cfg#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas=-v -I include $flags -c autorally_b200/csrc/rollout_tc.cu -o $L/exp/rollout_tc_$name.o 2>&1 | grep -A2 ILi32 | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
  objs=$(ls $L/*.o | grep -v rollout_tc.o)
  nvcc -shared -o $L/exp/libmppi_b200_exp$name.so $objs $L/exp/rollout_tc_$name.o -gencode arch=compute_100a,code=sm_100a -ldl
done
