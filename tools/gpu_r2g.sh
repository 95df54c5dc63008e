#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $OUT/pytest_r02g.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_r02g.log
(python tools/exp_overlap.py; MPPI_NO_SPLIT_FINALIZE=1 python tools/exp_overlap.py; MPPI_NO_PDL=1 MPPI_NO_SPLIT_FINALIZE=1 python tools/exp_overlap.py; MPPI_NO_PDL=1 python tools/exp_overlap.py) > $OUT/exp_overlap_r02g.txt 2>&1; sort -k3,3 -k4,4 -s $OUT/exp_overlap_r02g.txt
