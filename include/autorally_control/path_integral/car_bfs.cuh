// car_bfs.cuh -- host twin of the AutoRally basis functions (API of PI/car_bfs.cuh:41-121).
// The device evaluation lives in libmppi_b200.so (autorally_b200/csrc/dynamics.cuh: CarBasisDyn).
#ifndef CAR_BFS_CUH_
#define CAR_BFS_CUH_
#include <cmath>

#include "managed.cuh"

namespace autorally_control {

class CarBasisFuncs : public Managed {
 public:
  // phi_idx(s, u): s = [x, y, yaw, roll, u_x, u_y, yaw_rate], u = [steering, throttle].
  float basisFuncX(int idx, float *s, float *u) {
    const float roll = s[3], vx = s[4], vy = s[5], wz = s[6], steer = u[0], thr = u[1];
    const bool moving = vx > .1;
    // tangent of the front-axle slip angle; the rear term uses vy/vx - .35 wz/vx
    const float tf = moving ? tanf(atanf(vy / vx + .45 * wz / vx) - steer) : tanf(-steer);
    const double rear = moving ? (double)(vy / vx) - .35 * wz / vx : 0.0;
    switch (idx) {
      case 0: return thr;
      case 1: return vx / 10.0;
      case 2: return sinf(steer) * tf / 1200.0;
      case 3: return sinf(steer) * tf * fabsf(tf) / 1440000.0;
      case 4: return sinf(steer) * powf(tf, 3) / 1728000000.0;
      case 5: return wz * vy / 25.0;
      case 6: return wz / 10.0;
      case 7: return vy / 10.0;
      case 8: return sinf(steer);
      case 9: return moving ? vy / vx / 40.0 : 0;
      case 10: return tf / 1400.0;
      case 11: return tf * fabsf(tf) / 1960000;
      case 12: return powf(tf, 3) / 2744000000;
      case 13: return moving ? rear / 40.0 : 0;
      case 14: return moving ? rear * std::fabs(rear) / 1600.0 : 0;
      case 15: return moving ? powf((float)rear, 3) / 64000.0 : 0;
      case 16: return wz * vx / 50.0;
      case 17: return roll;
      case 18: return roll * wz;
      case 19: return roll * vx / 3.0;
      case 20: return roll * vx * wz / 5.0;
      case 21: return powf(vx, 2) / 100.0;
      case 22: return powf(vx, 3) / 1000.0;
      case 23: return powf(thr, 2);
      case 24: return powf(thr, 3);
    }
    return 0;
  }
};

}  // namespace autorally_control
#endif
