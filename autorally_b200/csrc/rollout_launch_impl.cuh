// rollout_launch_impl.cuh -- shared body of the rollout launchers.
#pragma once
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

template <class DYN, int BLOCK, int MINB = 1>
cudaError_t launch_rollout_t(const RolloutParams &p, cudaStream_t st) {
  const long long total = (long long)p.B * p.n_local;
  const long long per_block = (long long)BLOCK * DYN::R;
  const unsigned grid = (unsigned)((total + per_block - 1) / per_block);
  const size_t smem = ((size_t)((DYN::SMEM_FLOATS + 3) & ~3) + (size_t)DYN::THREAD_SMEM_FLOATS * BLOCK) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rollout_kernel<DYN, BLOCK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  rollout_kernel<DYN, BLOCK, MINB><<<grid, BLOCK, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mppi
