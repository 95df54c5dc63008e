// car_kinematics.cuh -- planar kinematics of the AutoRally state (API of PI/car_kinematics.cuh:17-29).
#ifndef CAR_GL_CUH_
#define CAR_GL_CUH_
#include <cmath>

#include "car_bfs.cuh"
#include "managed.cuh"

namespace autorally_control {

class CarKinematics : public Managed {
 public:
  void computeKinematics(float *state, float *state_der) {
    const float c = cosf(state[2]), s = sinf(state[2]);
    state_der[0] = c * state[4] - s * state[5];
    state_der[1] = s * state[4] + c * state[5];
    state_der[2] = -state[6];  // the pose estimate reports the negative yaw rate
  }
};

}  // namespace autorally_control
#endif
