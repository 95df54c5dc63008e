// car_kinematics.cuh -- CarKinematics, the K_FUNC argument of GeneralizedLinear<CarBasisFuncs,7,2,25,CarKinematics,3>
// (SRC/path_integral_main.cu:74).  Host-side twin only: on the device the kinematics are part of the fused rollout
// kernels (autorally_b200/csrc/rollout.cuh).  Semantics of PI/car_kinematics.cuh:17-29.
#pragma once
#include <cmath>

#include "car_bfs.cuh"
#include "managed.cuh"

namespace autorally_control {

namespace kinematics_detail {
/// World-frame rates of (x, y, yaw) from the heading s[2] and the body-frame velocities s[4], s[5]; the state
/// estimator reports the yaw rate s[6] with the opposite sign, hence the minus.
inline void world_rates(const float *s, float *rates) {
  const float heading = s[2], forward = s[4], lateral = s[5];
  const float ch = cosf(heading), sh = sinf(heading);
  rates[0] = ch * forward - sh * lateral;
  rates[1] = sh * forward + ch * lateral;
  rates[2] = -s[6];
}
}  // namespace kinematics_detail

class CarKinematics : public Managed {
 public:
  void computeKinematics(float *state, float *state_der) { kinematics_detail::world_rates(state, state_der); }
};

}  // namespace autorally_control
