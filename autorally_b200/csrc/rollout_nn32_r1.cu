// NeuralNetModel<7,2,3,6,32,32,4>, one rollout per thread.
#include "rollout_launch_impl.cuh"
namespace mppi {
cudaError_t launch_rollout_nn32_r1(const RolloutParams &p, cudaStream_t st, bool small) {
  using D = NeuralNetDyn<1, 6, 32, 32, 4>;
  return small ? launch_rollout_t<D, 32, 1, true>(p, st) : launch_rollout_t<D, 128, 1, true>(p, st);
}
}  // namespace mppi
