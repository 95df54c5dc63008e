// GeneralizedLinear<CarBasisFuncs,7,2,25,CarKinematics,3>, one rollout per thread.
#include "rollout_launch_impl.cuh"
namespace mppi {
cudaError_t launch_rollout_bf(const RolloutParams &p, cudaStream_t st, bool small) {
  return small ? launch_rollout_t<CarBasisDyn, 32>(p, st) : launch_rollout_t<CarBasisDyn, 128>(p, st);
}
}  // namespace mppi
