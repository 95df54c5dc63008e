#!/bin/bash
# bench.py at N GPUs of one box (driver-style launch) + the multi-GPU tests.   tools/gpu_scale_r02.sh <N> [tag]
set -u
N=$1; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multi_gpu_nccl.py -m gpu -q -rf > $OUT/pytest_n${N}_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_n${N}_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "bench rc=$?"
grep "^{" $OUT/bench_${TAG}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','sharded_parity_max_rel','sharded_large_strong_efficiency','weak_large_efficiency','batched_efficiency')})
for k in ('sharded_large','weak_large','batched_4096x256x100','sharded_1920'): print(k, {a:b for a,b in d[k].items() if isinstance(b,(int,float))})
print(d['e2e'])"
tail -3 $OUT/bench_${TAG}_n$N.err
