// rollout_launch_impl.cuh -- shared body of the rollout launchers.
#pragma once
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

template <class DYN, int BLOCK, int MINB, bool FUSED>
cudaError_t launch_rollout_tf(const RolloutParams &p, cudaStream_t st) {
  const long long total = (long long)p.B * p.n_local;
  const long long per_block = (long long)BLOCK * DYN::R;
  const unsigned grid = (unsigned)((total + per_block - 1) / per_block);
  const size_t smem = ((size_t)((DYN::SMEM_FLOATS + 3) & ~3) + (size_t)DYN::THREAD_SMEM_FLOATS * BLOCK) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rollout_kernel<DYN, BLOCK, MINB, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  rollout_kernel<DYN, BLOCK, MINB, FUSED><<<grid, BLOCK, smem, st>>>(p);
  return cudaGetLastError();
}

// CAN_FUSE: also instantiate the variant that draws its Philox noise in place (chosen by p.fused_noise)
template <class DYN, int BLOCK, int MINB = 1, bool CAN_FUSE = false>
cudaError_t launch_rollout_t(const RolloutParams &p, cudaStream_t st) {
  if constexpr (CAN_FUSE) {
    if (p.fused_noise) return launch_rollout_tf<DYN, BLOCK, MINB, true>(p, st);
  } else {
    if (p.fused_noise) return cudaErrorInvalidValue;  // the host never asks for it (supports_fused_noise)
  }
  return launch_rollout_tf<DYN, BLOCK, MINB, false>(p, st);
}

}  // namespace mppi
