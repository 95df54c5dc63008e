#!/bin/bash
# Builds libmppi_b200 variants with one part of the tensor-core rollout kernel removed (timing experiments, wrong results):
#   1 no running cost   2 no MUFU.EX2   3 no MMA issue / wait   4 no MMA and no CTA barriers   5 fast sincos
#   9 cycle stamps of one timestep (tools/exp_tc_stamps.py)
#   pipe9: the layer-pipeline kernel with per-role work cycles per tick (-DPIPE_EXP=9, tools/exp_pipe_stamps.sh)

set -e
cd "$(dirname "$0")/.."
L=autorally_b200/lib; mkdir -p $L/exp
for n in ${@:-1 2 3 4 5}; do
  if [ "$n" = pipe9 ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -DPIPE_EXP=9 -c autorally_b200/csrc/rollout_pipe64.cu -o $L/exp/rollout_pipe64_9.o
    nvcc -shared -o $L/exp/libmppi_b200_pipe9.so $(ls $L/*.o | grep -v rollout_pipe64.o) $L/exp/rollout_pipe64_9.o -gencode arch=compute_100a,code=sm_100a -ldl
    continue
  fi
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -DTC_EXP=$n -c autorally_b200/csrc/rollout_tc.cu -o $L/exp/rollout_tc_$n.o
  objs=$(ls $L/*.o | grep -v rollout_tc.o)
  nvcc -shared -o $L/exp/libmppi_b200_exp$n.so $objs $L/exp/rollout_tc_$n.o -gencode arch=compute_100a,code=sm_100a -ldl
done
