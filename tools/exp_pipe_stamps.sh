#!/bin/bash
# work cycles per tick of the roles of the layer-pipeline kernel (experiment library: tools/build_tc_exp.sh pipe9)
L=autorally_b200/lib
cp $L/libmppi_b200.so /tmp/keep.so
cp $L/exp/libmppi_b200_pipe9.so $L/libmppi_b200.so
timeout 100 python tools/exp_pipe_stamps.py 256
timeout 100 python tools/exp_pipe_stamps.py 1920
cp /tmp/keep.so $L/libmppi_b200.so
