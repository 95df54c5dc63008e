#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, the bench line, the per-launch time list and the
# full ncu captures of the dominant kernels.  Outputs go to gpurun_out/; summaries are copied into
# profiles/ by tools/summarize_ncu.py on the CPU side.
#   tools/gpu_profile.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
# (1) launch list of the benchmark configuration (1920 x 100): shares of the step per kernel
P1="python tools/profile_step.py --rollouts 1920 --steps 6"
$P1 > $OUT/plain_1920_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_1920_$TAG.csv $P1 > $OUT/ncu_1920_list.log 2>&1
echo "launch list 1920 rc=$?"
# (2) full capture of one complete step at 1920 x 100 (noise, rollout, weighting, finalize)
$P1 > $OUT/plain_1920b_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 8 -c 4 -f -o $OUT/prof_1920_$TAG $P1 > $OUT/ncu_1920_full.log 2>&1
echo "full 1920 rc=$?"
# (3) the filled-GPU configuration (1M rollouts): launch list + full capture of one step
P2="python tools/profile_step.py --rollouts 1048576 --steps 2"
$P2 > $OUT/plain_1m_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $OUT/launches_1m_$TAG.csv $P2 > $OUT/ncu_1m_list.log 2>&1
echo "launch list 1m rc=$?"
$P2 > $OUT/plain_1mb_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 4 -c 4 -f -o $OUT/prof_1m_$TAG $P2 > $OUT/ncu_1m_full.log 2>&1
echo "full 1m rc=$?"
cat $OUT/plain_1920_$TAG.log $OUT/plain_1m_$TAG.log
