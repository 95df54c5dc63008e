// Stand-in for the dynamic_reconfigure generated config of cfg/PathIntegralParams.cfg:12-21 (no ROS).
#ifndef MPPI_COMPAT_PATH_INTEGRAL_PARAMS_CONFIG_H_
#define MPPI_COMPAT_PATH_INTEGRAL_PARAMS_CONFIG_H_
namespace autorally_control {
struct PathIntegralParamsConfig {
  double max_throttle = 0.65;
  double desired_speed = 6.0;
  double speed_coefficient = 4.25;
  double track_coefficient = 200.0;
  double max_slip_angle = 1.25;
  double slip_penalty = 10.0;
  double crash_coefficient = 10000;
  double track_slop = 0;
  double steering_coeff = 0.0;
  double throttle_coeff = 0.0;
};
}  // namespace autorally_control
#endif
