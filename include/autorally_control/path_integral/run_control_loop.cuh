// run_control_loop.cuh -- runControlLoop<CONTROLLER_T>: the caller of the hot path (PI/run_control_loop.cuh:84-321),
// without ROS.  Same tube scheme as the reference: two controllers share one model and one cost object; one plans from
// the measured state, the other from its own predicted state; each iteration the sequences are slid by the number of
// controls executed, both plan, and the plan with the SMALLER getComputedTrajectoryCost() is handed to the plant (when
// the measured-state plan wins, the predicted-state controller adopts its sequences).  In debug mode the loop
// integrates the host dynamics model as the plant (:296-302).
//
// Differences from the reference: the plant is a template parameter (the reference hard-wires the ROS AutorallyPlant;
// any class with the methods of sim_plant.h works, AutorallyPlant included, its ros::Time mapping to seconds); the
// optional boolean parameter "sleep_to_rate" (default true) lets simulations run faster than real time; the debug
// costmap window (OpenCV) is not drawn.
#ifndef MPPI_RUN_CONTROL_LOOP_CUH_
#define MPPI_RUN_CONTROL_LOOP_CUH_
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <Eigen/Dense>

#include "param_getter.h"
#include "sim_plant.h"

namespace autorally_control {

// Optional plant hooks (sim_plant.h has them, a plant modelled on AutorallyPlant need not): injected noise and an arbitration log.
template <class P, class C>
auto loop_inject_noise(P *robot, C *a, C *b, int iteration, int) -> decltype(robot->hasInjectedNoise(0), void()) {
  if (!robot->hasInjectedNoise(iteration)) return;
  a->setNoise(robot->injectedNoise(iteration), robot->injectedNoiseCount());
  b->setNoise(robot->injectedNoise(iteration), robot->injectedNoiseCount());
}
template <class P, class C>
void loop_inject_noise(P *, C *, C *, int, long) {}
template <class P, class C, class G>
auto loop_log_arbitration(P *robot, C *a, C *b, const G &gains, int T, int) -> decltype(robot->trajectoryCosts(), void()) {
  robot->logArbitration(a->getComputedTrajectoryCost(), b->getComputedTrajectoryCost(), a->getControlSequenceU(), b->getControlSequenceU(), gains, T);
}
template <class P, class C, class G>
void loop_log_arbitration(P *, C *, C *, const G &, int, long) {}

template <class CONTROLLER_T, class PLANT_T>
void runControlLoop(CONTROLLER_T *predicted_state_controller, CONTROLLER_T *actual_state_controller, PLANT_T *robot,
                    std::map<std::string, XmlRpc::XmlRpcValue> *params, std::atomic<bool> *is_alive) {
  auto flag = [&](const char *key, bool dflt) { return params->count(key) ? (bool)(*params)[key] : dflt; };
  const int hz = (int)(*params)["hz"];
  const int optimization_stride = (int)(*params)["optimization_stride"];
  const int num_timesteps = (int)(*params)["num_timesteps"];
  const bool use_feedback_gains = flag("use_feedback_gains", false);
  const bool debug_mode = flag("debug_mode", false);
  const bool only_actual = flag("use_only_actual_state_controller", false);
  const bool only_predicted = flag("use_only_predicted_state_controller", false);
  const bool sleep_to_rate = flag("sleep_to_rate", true);
  const bool debug_double_step = flag("reference_debug_double_step", false);
  const int max_iter = params->count("profiler_max_iter") ? (int)(*params)["profiler_max_iter"] : INT_MAX;  // :105-106

  Eigen::MatrixXf state(7, 1), u(2, 1);
  std::vector<float> control_solution, state_solution;
  std::vector<int> description;
  std::vector<float> data;
  int num_iter = 0, status = 1;
  double avg_loop_ms = 0, avg_tick_ms = 0, avg_sleep_ms = 0;
  double last_pose_update = robot->getLastPoseTime();
  double loop_time_s = optimization_stride / (1.0 * hz);
  const std::chrono::duration<double, std::milli> period(optimization_stride * 1000.0 / hz);

  if (!debug_mode)  // wait for the first pose estimate (:138-142)
    while (last_pose_update == robot->getLastPoseTime() && is_alive->load()) std::this_thread::sleep_for(std::chrono::microseconds(50));

  auto read_state = [&]() {
    const typename PLANT_T::FullState fs = robot->getState();
    state(0) = fs.x_pos; state(1) = fs.y_pos; state(2) = fs.yaw; state(3) = fs.roll; state(4) = fs.u_x; state(5) = fs.u_y; state(6) = fs.yaw_mder;
  };
  read_state();
  Eigen::Matrix<float, 7, 1> s7;
  auto fixed = [&]() { for (int i = 0; i < 7; i++) s7(i) = state(i); return s7; };
  actual_state_controller->setState(fixed());
  predicted_state_controller->setState(fixed());
  actual_state_controller->resetControls();
  actual_state_controller->computeFeedbackGains(state);
  predicted_state_controller->resetControls();
  predicted_state_controller->computeFeedbackGains(state);

  while (is_alive->load() && num_iter < max_iter) {
    const auto loop_start = std::chrono::steady_clock::now();
    robot->setTimingInfo(avg_loop_ms, avg_tick_ms, avg_sleep_ms);
    num_iter++;
    if (last_pose_update != robot->getLastPoseTime()) {  // new state estimate (:173-178)
      loop_time_s = robot->getLastPoseTime() - last_pose_update;
      last_pose_update = robot->getLastPoseTime();
      read_state();
    }
    if (robot->hasNewDynRcfg()) {  // both controllers share one cost object; the reference updates through both (:183-186)
      const PathIntegralParamsConfig cfg = robot->getDynRcfgParams();
      actual_state_controller->costs_->updateParams_dcfg(cfg);
      predicted_state_controller->costs_->updateParams_dcfg(cfg);
    }
    if (robot->hasNewObstacles()) {
      robot->getObstacles(description, data);
      actual_state_controller->costs_->updateObstacles(description, data);
      predicted_state_controller->costs_->updateObstacles(description, data);
    }
    if (robot->hasNewCostmap()) {
      robot->getCostmap(description, data);
      actual_state_controller->costs_->updateCostmap(description, data);
      predicted_state_controller->costs_->updateCostmap(description, data);
    }
    if (robot->hasNewModel()) {  // model hot swap (:200-204)
      robot->getModel(description, data);
      actual_state_controller->model_->updateModel(description, data);
      predicted_state_controller->model_->updateModel(description, data);
    }
    // slide by the number of controls executed since the last iteration (:208-215)
    int stride = (int)std::lround(loop_time_s * hz);
    if (status != 0) stride = optimization_stride;
    if (stride >= 0 && stride < num_timesteps) {
      actual_state_controller->slideControlAndStateSeq(stride);
      predicted_state_controller->slideControlAndStateSeq(stride);
    }
    loop_inject_noise(robot, actual_state_controller, predicted_state_controller, num_iter - 1, 0);
    // the hot path, twice (:218-219).  The two plans are independent, so both are enqueued before either is awaited:
    // each fills a few percent of a B200 and they overlap on the device.
    actual_state_controller->computeControlAsync(fixed());
    predicted_state_controller->computeControlAsync();
    actual_state_controller->waitControl();
    predicted_state_controller->waitControl();
    if (use_feedback_gains) {
      actual_state_controller->computeFeedbackGains(state);
      predicted_state_controller->computeFeedbackGains(state);
    }
    auto feedback_gain = predicted_state_controller->getFeedbackGains().feedback_gain;
    // arbitration (:229-285)
    ControllerType to_use = ControllerType::NONE;
    if (only_actual && !only_predicted) to_use = ControllerType::ACTUAL_STATE;
    else if (!only_actual && only_predicted) to_use = ControllerType::PREDICTED_STATE;
    if (to_use == ControllerType::NONE)
      to_use = actual_state_controller->getComputedTrajectoryCost() < predicted_state_controller->getComputedTrajectoryCost()
                   ? ControllerType::ACTUAL_STATE : ControllerType::PREDICTED_STATE;
    const bool arbitrated = !(only_actual != only_predicted);
    if (to_use == ControllerType::ACTUAL_STATE) {
      control_solution = actual_state_controller->getControlSeq();
      state_solution = actual_state_controller->getStateSeq();
      feedback_gain = actual_state_controller->getFeedbackGains().feedback_gain;
      if (arbitrated) {  // the predicted-state controller adopts the winning sequences (:259-260)
        predicted_state_controller->setStateSequence(state_solution);
        predicted_state_controller->setControlSequence(control_solution);
      }
    } else {
      control_solution = predicted_state_controller->getControlSeq();
      state_solution = predicted_state_controller->getStateSeq();
      feedback_gain = predicted_state_controller->getFeedbackGains().feedback_gain;
    }
    loop_log_arbitration(robot, actual_state_controller, predicted_state_controller, feedback_gain, num_timesteps, 0);
    robot->setSolution(state_solution, control_solution, feedback_gain, last_pose_update, avg_loop_ms, to_use);
    status = robot->checkStatus();
    if (status != 0 && debug_mode) {  // the host model is the plant (:296-302)
      for (int t = 0; t < optimization_stride; t++) {
        u(0) = control_solution[2 * t]; u(1) = control_solution[2 * t + 1];
        actual_state_controller->model_->updateState(state, u);
        // The reference then calls updateState through the second controller too (:299-300); main() gives both
        // controllers the SAME model object, so its simulated car advances two model steps per control.  That is a
        // quirk of its debug plant, not of the controller; it is reproduced only on request.
        if (debug_double_step) predicted_state_controller->model_->updateState(state, u);
      }
    }
    std::chrono::duration<double, std::milli> elapsed = std::chrono::steady_clock::now() - loop_start;
    const double tick_ms = elapsed.count();
    while (sleep_to_rate && is_alive->load() &&
           (elapsed < period || ((robot->getLastPoseTime() - last_pose_update) < (1.0 / hz - 0.0025) && status == 0))) {
      std::this_thread::sleep_for(std::chrono::microseconds(50));
      elapsed = std::chrono::steady_clock::now() - loop_start;
    }
    const double sleep_ms = elapsed.count() - tick_ms;
    avg_loop_ms = (num_iter - 1.0) / num_iter * avg_loop_ms + 1000.0 * loop_time_s / num_iter;
    avg_tick_ms = (num_iter - 1.0) / num_iter * avg_tick_ms + tick_ms / num_iter;
    avg_sleep_ms = (num_iter - 1.0) / num_iter * avg_sleep_ms + sleep_ms / num_iter;
  }
  robot->setTimingInfo(avg_loop_ms, avg_tick_ms, avg_sleep_ms);
}

}  // namespace autorally_control
#endif
