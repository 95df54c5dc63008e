#!/bin/bash
# 2 GPUs: multi-GPU tests + the N = 2 bench line
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -rf > $OUT/pytest_n2_r02f.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_n2_r02f.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_r02f.json 2> $OUT/bench_n2_r02f.err; echo "bench rc=$?"; cat $OUT/bench_n2_r02f.json; tail -5 $OUT/bench_n2_r02f.err
