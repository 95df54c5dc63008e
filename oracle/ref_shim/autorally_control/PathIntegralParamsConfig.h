// Stand-in for the dynamic_reconfigure-generated header (cfg/PathIntegralParams.cfg:12-21): same field
// names and types, no ROS.  TEST INFRASTRUCTURE (oracle/refbuild.py).
#ifndef REF_SHIM_PATH_INTEGRAL_PARAMS_CONFIG_H_
#define REF_SHIM_PATH_INTEGRAL_PARAMS_CONFIG_H_
namespace autorally_control {
struct PathIntegralParamsConfig {
  double desired_speed = 6.0, max_throttle = 0.65, speed_coefficient = 4.25, track_coefficient = 200.0;
  double max_slip_angle = 1.25, slip_penalty = 10.0, crash_coefficient = 10000.0, track_slop = 0.0;
  double steering_coeff = 0.0, throttle_coeff = 0.0;
};
}  // namespace autorally_control
#endif
