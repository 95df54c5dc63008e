#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $OUT/pytest_r02b.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_r02b.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_r02b.json 2> $OUT/bench_ref_r02b.err; echo "ref rc=$?"; cat $OUT/bench_ref_r02b.json; tail -3 $OUT/bench_ref_r02b.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02b.json 2> $OUT/bench_r02b.err; echo "bench rc=$?"; cat $OUT/bench_r02b.json; tail -3 $OUT/bench_r02b.err
