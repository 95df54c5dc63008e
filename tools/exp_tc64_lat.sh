#!/bin/bash
# the tensor-core kernels at controller size: step times, then cycle stamps of one timestep (experiment library 9)
L=autorally_b200/lib
timeout 120 python tools/exp_generic.py tc64
MPPI_TC_NO_SLICES=1 timeout 120 python tools/exp_generic.py tc64
timeout 120 python tools/exp_generic.py tc32
cp $L/libmppi_b200.so /tmp/keep.so
cp $L/exp/libmppi_b200_exp9.so $L/libmppi_b200.so
echo "== wider_deeper, sliced"; timeout 120 python tools/exp_tc_stamps.py wider_deeper 1920
echo "== 6-32-32-4"; timeout 120 python tools/exp_tc_stamps.py autorally_nnet 1920
cp /tmp/keep.so $L/libmppi_b200.so
