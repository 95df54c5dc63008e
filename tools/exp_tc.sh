#!/bin/bash
# crossover between the latency kernel (9), the FFMA2 kernel (2) and the tensor-core kernel (10)
for n in 8192 16384 32768 65536 131072 262144 1048576; do for v in 9 2 10; do
  echo -n "N=$n: "; python tools/profile_step.py --rollouts $n --steps 5 --variant $v
done; done
