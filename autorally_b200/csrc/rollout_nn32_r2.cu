// NeuralNetModel<7,2,3,6,32,32,4>, two rollouts per thread packed in f32x2 registers (FFMA2).
//
// CTA shape (measured at 1M rollouts x 100 steps on B200, profiles/exp_r2_r01.txt): one-warp CTAs capped at 170
// registers (12 resident warps per SM) ran the rollout kernel in 7.72 ms, 64-thread CTAs x 6 in 7.90 ms, the
// uncapped 128-thread shape (180 registers, 8 warps per SM) in 8.20 ms, and a 128-register cap (spills) in 11.2 ms.
#include <cstdlib>
#include "rollout_launch_impl.cuh"
namespace mppi {
cudaError_t launch_rollout_nn32_r2(const RolloutParams &p, cudaStream_t st, bool small) {
  using D = NeuralNetDynP2<0, 6, 32, 32, 4>;
  if (small) return launch_rollout_t<D, 32>(p, st);
  // MPPI_R2_CONFIG selects the other measured shapes (experiments only)
  static const int cfg = std::getenv("MPPI_R2_CONFIG") ? std::atoi(std::getenv("MPPI_R2_CONFIG")) : 5;
  switch (cfg) {
    case 0: return launch_rollout_t<D, 128>(p, st);
    case 2: return launch_rollout_t<D, 64, 6>(p, st);
    case 3: return launch_rollout_t<D, 128, 3>(p, st);
    default: return launch_rollout_t<D, 32, 12>(p, st);
  }
}
}  // namespace mppi
