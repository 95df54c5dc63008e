// host_api_driver.cu -- exercises the drop-in C++ template API the way the reference's main() does
// (SRC/path_integral_main.cu:80-153): launch file -> params map -> MPPICosts -> dynamics model ->
// MPPIController -> computeControl, then the slide / computeControl() / feedback-gain sequence of
// runControlLoop (PI/run_control_loop.cuh:208-225).  Results go to an .npz the pytest side compares
// with the CPU oracle.
//
// usage: host_api_driver <nn|bf> <launch_file> <noise.bin> <out.npz> x y yaw roll ux uy yawrate [swap.npz]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include <autorally_control/path_integral/meta_math.h>
#include <autorally_control/path_integral/param_getter.h>
#include <autorally_control/path_integral/costs.cuh>
#include <autorally_control/path_integral/neural_net_model.cuh>
#include <autorally_control/path_integral/car_bfs.cuh>
#include <autorally_control/path_integral/car_kinematics.cuh>
#include <autorally_control/path_integral/generalized_linear.cuh>
#include <autorally_control/path_integral/mppi_controller.cuh>

using namespace autorally_control;

template <class M>
void swap_model(M *model, const std::string &path) {
  npz::Archive a = npz::load(path);
  const npz::Array &d = a.at("description"), &v = a.at("data");
  std::vector<int> description(d.num_vals());
  std::vector<float> data(v.num_vals());
  for (size_t i = 0; i < description.size(); i++) description[i] = (int)d.at(i);
  for (size_t i = 0; i < data.size(); i++) data[i] = (float)v.at(i);
  model->updateModel(description, data);
}
template <>
void swap_model(GeneralizedLinear<CarBasisFuncs, 7, 2, 25, CarKinematics, 3> *, const std::string &) {}

template <class Controller, class DynamicsModel>
int run(const std::string &launch, const std::string &noise_path, const std::string &out_path, const float *st, const std::string &swap_path) {
  std::map<std::string, XmlRpc::XmlRpcValue> params;
  loadParams(&params, launch);
  MPPICosts *costs = new MPPICosts(&params);
  float2 control_constraints[2] = {make_float2(-.99, .99), make_float2(-.99, (double)params["max_throttle"])};
  DynamicsModel *model = new DynamicsModel(1.0 / (int)params["hz"], control_constraints);
  model->loadParams((std::string)params["model_path"]);
  if (params.count("negate_yaw_der")) model->negate_yaw_der = params["negate_yaw_der"];
  float exploration_std[2] = {(float)(double)params["steering_std"], (float)(double)params["throttle_std"]};
  float init_u[2] = {(float)(double)params["init_steering"], (float)(double)params["init_throttle"]};
  const int hz = (int)params["hz"], T = (int)params["num_timesteps"], stride = (int)params["optimization_stride"];
  const float gamma = (float)(double)params["gamma"];
  const int num_iters = (int)params["num_iters"];
  Controller *ctl = new Controller(model, costs, exploration_std, init_u, hz, T, stride, gamma, num_iters);

  std::ifstream nf(noise_path.c_str(), std::ios::binary);
  std::vector<float> eps((size_t)2 * Controller::NUM_ROLLOUTS * T * 2);  // two calls' worth
  nf.read(reinterpret_cast<char *>(eps.data()), (std::streamsize)(eps.size() * sizeof(float)));
  if (!nf) { fprintf(stderr, "noise file too short\n"); return 2; }
  const size_t per_call = (size_t)Controller::NUM_ROLLOUTS * T * 2;

  Eigen::Matrix<float, 7, 1> state;
  for (int i = 0; i < 7; i++) state(i) = st[i];
  npz::Writer w;
  // ---- call 1: computeControl(state) from reset controls ----
  ctl->setState(state);
  ctl->resetControls();
  ctl->setNoise(eps.data(), per_call);
  ctl->computeControl(state);
  std::vector<float> U1 = ctl->getControlSequenceU(), cs1 = ctl->getControlSeq(), ss1 = ctl->getStateSeq(), rc1 = ctl->getRolloutCosts();
  const float tc1 = ctl->getComputedTrajectoryCost(), b1 = ctl->getBaseline(), z1 = ctl->getNormalizer();
  w.add("U1", U1.data(), {(size_t)T, 2}); w.add("control_solution1", cs1.data(), {(size_t)T, 2});
  w.add("state_solution1", ss1.data(), {(size_t)T, 7}); w.add("rollout_costs1", rc1.data(), {rc1.size()});
  const float stats1[3] = {b1, z1, tc1};
  w.add("stats1", stats1, {3});
  // ---- feedback gains around the solution ----
  Eigen::MatrixXf sx(7, 1);
  for (int i = 0; i < 7; i++) sx(i) = state(i);
  ctl->computeFeedbackGains(sx);
  auto res = ctl->getFeedbackGains();
  std::vector<float> gains((size_t)T * 2 * 7, 0.0f);
  for (int k = 0; k < (int)res.feedback_gain.size() && k < T; k++)
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 7; c++) gains[((size_t)k * 2 + r) * 7 + c] = res.feedback_gain[k](r, c);
  w.add("feedback_gain", gains.data(), {(size_t)T, 2, 7});
  // ---- call 2: slide by the optimisation stride, plan from the predicted state ----
  ctl->slideControlAndStateSeq(stride);
  ctl->setNoise(eps.data() + per_call, per_call);
  ctl->computeControl();
  std::vector<float> U2 = ctl->getControlSequenceU(), ss2 = ctl->getStateSeq();
  w.add("U2", U2.data(), {(size_t)T, 2}); w.add("state_solution2", ss2.data(), {(size_t)T, 7});
  const float tc2 = ctl->getComputedTrajectoryCost();
  w.add("trajectory_cost2", &tc2, {1});
  // ---- model hot swap (updateModel: the flattened /model_updater/model message) must reach the DEVICE: the rollouts and the
  //      nominal trajectory of the next computeControl run on the new weights ----
  if (!swap_path.empty()) {
    swap_model(model, swap_path);
    std::vector<float> Uin = ctl->getControlSequenceU();
    w.add("U_in_swap", Uin.data(), {(size_t)T, 2});
    ctl->setNoise(eps.data(), per_call);
    ctl->computeControl(state);
    std::vector<float> Us = ctl->getControlSequenceU(), sss = ctl->getStateSeq(), rcs = ctl->getRolloutCosts();
    w.add("U_swap", Us.data(), {(size_t)T, 2}); w.add("state_solution_swap", sss.data(), {(size_t)T, 7});
    w.add("rollout_costs_swap", rcs.data(), {rcs.size()});
  }
  // ---- updateControlNoise + cutThrottle + the Philox sampler path ----
  const float wide[2] = {0.4f, 0.5f};
  ctl->updateControlNoise(wide);
  ctl->useSampler();
  ctl->cutThrottle();
  ctl->computeControl(state);
  std::vector<float> cs3 = ctl->getControlSeq();
  w.add("control_solution3", cs3.data(), {(size_t)T, 2});
  w.save(out_path);
  ctl->deallocateCudaMem();
  ctl->deallocateCudaMem();  // idempotent
  delete ctl;
  delete costs;
  delete model;
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 12) { fprintf(stderr, "usage: %s <nn|bf> launch noise.bin out.npz x y yaw roll ux uy yawrate\n", argv[0]); return 1; }
  float st[7];
  for (int i = 0; i < 7; i++) st[i] = (float)atof(argv[5 + i]);
  const std::string kind = argv[1];
  if (kind == "nn") {
    typedef NeuralNetModel<7, 2, 3, 6, 32, 32, 4> DynamicsModel;                      // SRC/path_integral_main.cu:66-69
    typedef MPPIController<DynamicsModel, MPPICosts, 1920, 8, 16> Controller;
    return run<Controller, DynamicsModel>(argv[2], argv[3], argv[4], st, argc > 12 ? argv[12] : "");
  }
  typedef GeneralizedLinear<CarBasisFuncs, 7, 2, 25, CarKinematics, 3> DynamicsModel;  // :71-74
  typedef MPPIController<DynamicsModel, MPPICosts, 2560, 16, 4> Controller;
  return run<Controller, DynamicsModel>(argv[2], argv[3], argv[4], st, argc > 12 ? argv[12] : "");
}
