// sim_plant.h -- SimPlant: an in-process stand-in for AutorallyPlant (PI/autorally_plant.h:165-256) with the methods
// runControlLoop calls.  The reference's plant is the ROS I/O node (pose / servo subscribers, chassisCommand
// publisher); here the "robot" is whatever the loop itself integrates in debug mode
// (PI/run_control_loop.cuh:296-302: the host dynamics model is the plant), and SimPlant only keeps the bookkeeping:
// the last solution handed over, the executed state / control log, timing averages, pending parameter updates.
#ifndef MPPI_SIM_PLANT_H_
#define MPPI_SIM_PLANT_H_
#include <vector>

#include "costs.cuh"

namespace autorally_control {

enum class ControllerType { NONE, ACTUAL_STATE, PREDICTED_STATE };

class SimPlant {
 public:
  struct FullState {
    float x_pos = 0, y_pos = 0, z_pos = 0, roll = 0, pitch = 0, yaw = 0, q0 = 1, q1 = 0, q2 = 0, q3 = 0;
    float x_vel = 0, y_vel = 0, z_vel = 0, u_x = 0, u_y = 0, yaw_mder = 0, steering = 0, throttle = 0;
  };

  SimPlant(float x, float y, float yaw) { full_state_.x_pos = x; full_state_.y_pos = y; full_state_.yaw = yaw; }

  // ---- what runControlLoop uses ----
  double getLastPoseTime() const { return last_pose_time_; }
  FullState getState() const { return full_state_; }
  void setTimingInfo(double avg_loop, double avg_tick, double avg_sleep) { avg_loop_ = avg_loop; avg_tick_ = avg_tick; avg_sleep_ = avg_sleep; }
  bool hasNewDynRcfg() const { return has_dcfg_; }
  PathIntegralParamsConfig getDynRcfgParams() { has_dcfg_ = false; return dcfg_; }
  bool hasNewObstacles() const { return false; }
  void getObstacles(std::vector<int> &, std::vector<float> &) {}
  bool hasNewCostmap() const { return false; }
  void getCostmap(std::vector<int> &, std::vector<float> &) {}
  bool hasNewModel() const { return has_model_ && (int)controller_used_.size() >= model_at_iteration_; }
  void getModel(std::vector<int> &description, std::vector<float> &data) { description = model_description_; data = model_data_; has_model_ = false; }
  template <class GAINS>
  void setSolution(const std::vector<float> &state_seq, const std::vector<float> &control_seq, const GAINS &, double ts,
                   double loop_speed, ControllerType used) {
    state_seq_ = state_seq; control_seq_ = control_seq; solution_ts_ = ts; (void)loop_speed;
    executed_states_.insert(executed_states_.end(), state_seq.begin(), state_seq.begin() + 7);
    executed_controls_.insert(executed_controls_.end(), control_seq.begin(), control_seq.begin() + 2);
    controller_used_.push_back(used == ControllerType::ACTUAL_STATE ? 0 : 1);
  }
  /// Test hooks (no counterpart in the reference): per-iteration N(0,1) draws to inject into BOTH controllers -- the
  /// reference's two controllers each seed cuRAND with 1234 and therefore draw the same sequence -- and a log of what the
  /// arbitration saw and handed over.
  bool hasInjectedNoise(int iteration) const { return noise_per_iter_ > 0 && (size_t)(iteration + 1) * noise_per_iter_ <= noise_.size(); }
  const float *injectedNoise(int iteration) const { return noise_.data() + (size_t)iteration * noise_per_iter_; }
  size_t injectedNoiseCount() const { return noise_per_iter_; }
  void setInjectedNoise(std::vector<float> noise, size_t per_iteration) { noise_ = std::move(noise); noise_per_iter_ = per_iteration; }
  template <class GAINS>
  void logArbitration(float cost_actual, float cost_predicted, const std::vector<float> &u_actual, const std::vector<float> &u_predicted,
                      const GAINS &gains, int timesteps) {
    trajectory_costs_.push_back(cost_actual); trajectory_costs_.push_back(cost_predicted);
    u_actual_.insert(u_actual_.end(), u_actual.begin(), u_actual.end());
    u_predicted_.insert(u_predicted_.end(), u_predicted.begin(), u_predicted.end());
    for (int k = 0; k < timesteps; k++)
      for (int a = 0; a < 2; a++)
        for (int b = 0; b < 7; b++) gains_.push_back(k < (int)gains.size() ? gains[k](a, b) : 0.0f);
  }
  const std::vector<float> &trajectoryCosts() const { return trajectory_costs_; }  // [iterations][2]: actual, predicted
  const std::vector<float> &loggedGains() const { return gains_; }                // [iterations][T][2][7]
  const std::vector<float> &loggedUActual() const { return u_actual_; }
  const std::vector<float> &loggedUPredicted() const { return u_predicted_; }
  /// 1 = "no pose updates: integrate the model" -- the reference's status for debug mode (PI/run_control_loop.cuh:296)
  int checkStatus() const { return 1; }
  template <class IMG> void setDebugImage(const IMG &) {}

  // ---- test / driver side ----
  void pushDynRcfg(const PathIntegralParamsConfig &c) { dcfg_ = c; has_dcfg_ = true; }
  /// Queue a model for hot swap (the /model_updater/model topic of the reference, SRC/autorally_plant.cpp:262-301);
  /// it becomes visible to the loop once `at_iteration` solutions have been handed over.
  void pushModel(const std::vector<int> &description, const std::vector<float> &data, int at_iteration = 0) {
    model_description_ = description; model_data_ = data; has_model_ = true; model_at_iteration_ = at_iteration;
  }
  const std::vector<float> &executedStates() const { return executed_states_; }      // [iterations][7]
  const std::vector<float> &executedControls() const { return executed_controls_; }  // [iterations][2]
  const std::vector<int> &controllerUsed() const { return controller_used_; }
  double avgTickMs() const { return avg_tick_; }

 private:
  FullState full_state_;
  double last_pose_time_ = 0.0, solution_ts_ = 0.0, avg_loop_ = 0, avg_tick_ = 0, avg_sleep_ = 0;
  bool has_dcfg_ = false, has_model_ = false;
  int model_at_iteration_ = 0;
  PathIntegralParamsConfig dcfg_;
  std::vector<int> model_description_, controller_used_;
  std::vector<float> model_data_, state_seq_, control_seq_, executed_states_, executed_controls_;
  std::vector<float> noise_, trajectory_costs_, gains_, u_actual_, u_predicted_;
  size_t noise_per_iter_ = 0;
};

}  // namespace autorally_control
#endif
