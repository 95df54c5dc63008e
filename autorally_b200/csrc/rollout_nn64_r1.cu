// NeuralNetModel<7,2,3,6,64,64,64,64,4> (wider_deeper_network_08_20_2020.npz), one rollout per thread.
#include "rollout_launch_impl.cuh"
namespace mppi {
cudaError_t launch_rollout_nn64_r1(const RolloutParams &p, cudaStream_t st, bool small) {
  using D = NeuralNetDyn<1, 6, 64, 64, 64, 64, 4>;
  (void)small;
  return launch_rollout_t<D, 64, 1, true>(p, st);
}
}  // namespace mppi
