#!/bin/bash
# Builds the micro-benchmark executables of tools/*.cu (sm_100a).  The binaries are git-ignored.
set -e
cd "$(dirname "$0")"
for t in latbench microbench mmabench texprobe; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $t.cu -o $t
done
