// control_loop_driver.cu -- the reference's main() (SRC/path_integral_main.cu:80-153) without ROS: launch file ->
// params map -> MPPICosts -> dynamics model -> TWO controllers sharing model and costs -> SimPlant ->
// runControlLoop in debug mode (the host model is the plant) for `profiler_max_iter` iterations.  Writes the executed
// state / control log for the closed-loop test (tests/test_control_loop.py).
//
// usage: control_loop_driver <nn|bf> <launch_file> <out.npz> <iterations> <x> <y> <heading> [double_step [swap.npz swap_iteration
//        [noise.bin use_feedback_gains]]]   (swap.npz may be "-"; noise.bin holds [iterations][NUM_ROLLOUTS][T][2] float32 draws that are
//        injected into BOTH controllers, as the reference's two cuRAND generators, both seeded 1234, draw the same sequence)
//   swap.npz: arrays `description` (int32 layer widths) and `data` (float32, all weights then all biases): the flattened
//   /model_updater/model message, handed to the loop after `swap_iteration` iterations (model hot swap).
#include <atomic>
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
#include <autorally_control/path_integral/run_control_loop.cuh>

using namespace autorally_control;

template <class Controller, class DynamicsModel>
int run(int argc, char **argv) {
  std::map<std::string, XmlRpc::XmlRpcValue> params;
  loadParams(&params, argv[2]);
  params["profiler_max_iter"] = XmlRpc::XmlRpcValue(atoi(argv[4]));
  params["x_pos"] = XmlRpc::XmlRpcValue(atof(argv[5]));
  params["y_pos"] = XmlRpc::XmlRpcValue(atof(argv[6]));
  params["heading"] = XmlRpc::XmlRpcValue(atof(argv[7]));
  params["debug_mode"] = XmlRpc::XmlRpcValue(true);
  params["sleep_to_rate"] = XmlRpc::XmlRpcValue(false);
  params["use_feedback_gains"] = XmlRpc::XmlRpcValue(argc > 12 && atoi(argv[12]) != 0);
  params["reference_debug_double_step"] = XmlRpc::XmlRpcValue(argc > 8 && atoi(argv[8]) != 0);
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
  // two controllers, one model, one cost object (SRC/path_integral_main.cu:119-122)
  Controller *actual = new Controller(model, costs, exploration_std, init_u, hz, T, stride, gamma, num_iters);
  Controller *predicted = new Controller(model, costs, exploration_std, init_u, hz, T, stride, gamma, num_iters);
  SimPlant robot((float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]));
  if (argc > 11) {
    std::ifstream nf(argv[11], std::ios::binary | std::ios::ate);
    const size_t bytes = (size_t)nf.tellg();
    nf.seekg(0);
    std::vector<float> noise(bytes / sizeof(float));
    nf.read(reinterpret_cast<char *>(noise.data()), (std::streamsize)bytes);
    robot.setInjectedNoise(std::move(noise), (size_t)Controller::NUM_ROLLOUTS * T * 2);
  }
  if (argc > 10 && std::string(argv[9]) != "-") {
    npz::Archive swap = npz::load(argv[9]);
    const npz::Array &d = swap.at("description"), &v = swap.at("data");
    std::vector<int> description(d.num_vals());
    std::vector<float> data(v.num_vals());
    for (size_t i = 0; i < description.size(); i++) description[i] = (int)d.at(i);
    for (size_t i = 0; i < data.size(); i++) data[i] = (float)v.at(i);
    robot.pushModel(description, data, atoi(argv[10]));
  }
  std::atomic<bool> is_alive(true);
  runControlLoop<Controller, SimPlant>(predicted, actual, &robot, &params, &is_alive);

  npz::Writer w;
  const std::vector<float> &st = robot.executedStates(), &ct = robot.executedControls();
  const size_t n = st.size() / 7;
  w.add("states", st.data(), {n, 7});
  w.add("controls", ct.data(), {n, 2});
  std::vector<float> used(robot.controllerUsed().begin(), robot.controllerUsed().end());
  w.add("controller_used", used.data(), {used.size()});
  if (!robot.trajectoryCosts().empty()) {
    w.add("trajectory_costs", robot.trajectoryCosts().data(), {used.size(), 2});
    w.add("gains", robot.loggedGains().data(), {used.size(), (size_t)T, 2, 7});
    w.add("U_actual", robot.loggedUActual().data(), {used.size(), (size_t)T, 2});
    w.add("U_predicted", robot.loggedUPredicted().data(), {used.size(), (size_t)T, 2});
  }
  const float tick = (float)robot.avgTickMs();
  w.add("avg_tick_ms", &tick, {1});
  w.save(argv[3]);
  actual->deallocateCudaMem();
  predicted->deallocateCudaMem();
  delete actual; delete predicted; delete costs; delete model;
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 8) { fprintf(stderr, "usage: %s <nn|bf> launch out.npz iterations x y heading [double_step]\n", argv[0]); return 1; }
  const std::string kind = argv[1];
  if (kind == "nn") {
    typedef NeuralNetModel<7, 2, 3, 6, 32, 32, 4> DynamicsModel;                      // SRC/path_integral_main.cu:66-69
    typedef MPPIController<DynamicsModel, MPPICosts, 1920, 8, 16> Controller;
    return run<Controller, DynamicsModel>(argc, argv);
  }
  typedef GeneralizedLinear<CarBasisFuncs, 7, 2, 25, CarKinematics, 3> DynamicsModel;  // :71-74
  typedef MPPIController<DynamicsModel, MPPICosts, 2560, 16, 4> Controller;
  return run<Controller, DynamicsModel>(argc, argv);
}
