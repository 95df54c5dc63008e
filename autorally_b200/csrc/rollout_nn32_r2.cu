// NeuralNetModel<7,2,3,6,32,32,4>, two rollouts per thread packed in f32x2 registers (FFMA2).
#include "rollout_launch_impl.cuh"
namespace mppi {
cudaError_t launch_rollout_nn32_r2(const RolloutParams &p, cudaStream_t st, bool small) {
  using D = NeuralNetDynP2<0, 6, 32, 32, 4>;
  return small ? launch_rollout_t<D, 32>(p, st) : launch_rollout_t<D, 128>(p, st);
}
}  // namespace mppi
