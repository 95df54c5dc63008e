// NeuralNetModel<7,2,3,6,32,32,4>, two rollouts per thread packed in f32x2 registers (FFMA2).
//
// Measured at 1M rollouts x 100 steps on B200 (profiles/exp_r2_r01.txt), rollout kernel only:
//   fully unrolled MLP (NeuralNetDynP2):   128-thread CTAs, 180 registers, 8 warps/SM     8.20 ms
//                                          one-warp CTAs capped at 168 registers, 12/SM     7.72 ms
//                                          capped at 128 registers (spills)                11.2  ms
//   layer 2 as a rolled loop (NeuralNetDynP2Compact, 128 registers, 16 warps/SM, no spills):
//                                          one-warp CTAs 7.61 ms, 64-thread 7.33 ms, 128-thread CTAs 7.22 ms  <- default
// Two further rewrites were measured and rejected: prefetching the next step's noise (7.97 ms) and finishing the cost
// after the MLP to hide the costmap fetch (7.26 ms); together they spill and run at 25 ms.
#include <cstdlib>
#include "rollout_launch_impl.cuh"
namespace mppi {
cudaError_t launch_rollout_nn32_r2(const RolloutParams &p, cudaStream_t st, bool small) {
  using D = NeuralNetDynP2<0, 6, 32, 32, 4>;
  if (small) return launch_rollout_t<D, 32>(p, st);
  // MPPI_R2_CONFIG selects the other measured shapes (experiments only)
  static const int cfg = std::getenv("MPPI_R2_CONFIG") ? std::atoi(std::getenv("MPPI_R2_CONFIG")) : 13;
  switch (cfg) {
    case 0: return launch_rollout_t<D, 128>(p, st);
    case 5: return launch_rollout_t<D, 32, 12>(p, st);
    case 11: return launch_rollout_t<NeuralNetDynP2Compact<32>, 32, 16>(p, st);
    case 12: return launch_rollout_t<NeuralNetDynP2Compact<64>, 64, 8>(p, st);
    default: return launch_rollout_t<NeuralNetDynP2Compact<128>, 128, 4>(p, st);
  }
}
}  // namespace mppi
