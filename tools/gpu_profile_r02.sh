#!/bin/bash
# Round-2 evidence run on the GPU box (under gpurun): GPU tests, both bench arms, ncu launch lists (ours and the reference's
# kernels) and full captures.  Every ncu command runs only after the same command has exited 0 without ncu.  Outputs go to
# gpurun_out/; tools/collect_profiles_r02.py turns them into the summaries under profiles/ on the CPU side.
#   tools/gpu_profile_r02.sh [tag]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
python -m pytest tests -m gpu -q -rf > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err; echo "bench rc=$?"; cut -c1-600 $OUT/bench_${TAG}_n1.json
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
FULL="ncu --set full --clock-control none --import-source on -f"
run_list() {  # name, count, command...
  local name=$1 cnt=$2; shift 2
  "$@" > $OUT/plain_${name}_$TAG.log 2>&1 && $LIST -c $cnt --log-file $OUT/launches_${name}_$TAG.csv "$@" > $OUT/nculist_${name}_$TAG.log 2>&1
  echo "launch list $name rc=$?"
}
run_full() {  # name, ncu selection args (quoted), command...
  local name=$1 sel=$2; shift 2
  "$@" > $OUT/plainf_${name}_$TAG.log 2>&1 && $FULL $sel -o $OUT/prof_${name}_$TAG "$@" > $OUT/ncufull_${name}_$TAG.log 2>&1
  echo "full capture $name rc=$?"
}
P1920="python tools/profile_step.py --rollouts 1920 --steps 6"
P1M="python tools/profile_step.py --rollouts 1048576 --steps 2"
P131K="python tools/profile_step.py --rollouts 131072 --steps 3"
PBF1M="python tools/profile_step.py --rollouts 1048576 --steps 2 --dynamics bf"
PBF="python tools/profile_step.py --rollouts 2560 --steps 6 --dynamics bf"
PWD64="python tools/profile_step.py --rollouts 1920 --steps 6 --tag wider_deeper"
run_list 1920 40 $P1920
run_list 1m 12 $P1M
run_list 131k 12 $P131K
run_list bf_1m 12 $PBF1M
run_list wd64_1920 40 $PWD64
run_list ref 80 python tools/ref_latency.py 3
run_full 1920 "-s 8 -c 4" $P1920
run_full 1m "-s 3 -c 3" $P1M
run_full 131k "-k regex:rollout_tc -s 1 -c 1" $P131K
run_full bf_1m "-k regex:rollout_kernel -s 1 -c 1" $PBF1M
run_full bf "-k regex:rollout_bf -s 2 -c 1" $PBF
run_full wd64_1920 "-k regex:rollout_pipe64 -s 2 -c 1" $PWD64
cat $OUT/plain_wd64_1920_$TAG.log $OUT/plain_1920_$TAG.log $OUT/plain_1m_$TAG.log $OUT/plain_131k_$TAG.log $OUT/plain_bf_1m_$TAG.log $OUT/plain_ref_$TAG.log
ls -la $OUT/*_$TAG.ncu-rep
