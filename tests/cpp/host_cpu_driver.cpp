// host_cpu_driver.cpp -- CPU-only checks of the host layer (no GPU calls): npz reader, launch-file
// parser, NeuralNetModel / GeneralizedLinear host twins, analytic Jacobian, and the DDP feedback-gain
// pass on a toy linear system.  Prints plain numbers; tests/test_host_cpp.py compares them with numpy /
// the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include <autorally_control/path_integral/param_getter.h>
#include <autorally_control/path_integral/costs.cuh>
#include <autorally_control/path_integral/neural_net_model.cuh>
#include <autorally_control/path_integral/car_bfs.cuh>
#include <autorally_control/path_integral/car_kinematics.cuh>
#include <autorally_control/path_integral/generalized_linear.cuh>
#include <autorally_control/ddp/ddp_feedback.h>
#define private public  // the reference's own test idiom (autorally_core/test/serialSensorInterfaceTest.cpp:43-61): U_, control_hist_
#include <autorally_control/path_integral/mppi_controller.cuh>
#undef private

using namespace autorally_control;

// toy linear system x' = Ac x + Bc u for the DDP check (no computeGrad -> central differences)
struct ToyLinear {
  static const int STATE_DIM = 2, CONTROL_DIM = 1;
  Eigen::Matrix<float, 2, 1> state_der_;
  void computeKinematics(Eigen::MatrixXf &) {}
  void computeDynamics(Eigen::MatrixXf &x, Eigen::MatrixXf &u) {
    state_der_(0) = x(1);
    state_der_(1) = -0.5f * x(0) - 0.1f * x(1) + 2.0f * u(0);
  }
};

int main(int argc, char **argv) {
  if (argc < 2) return 1;
  const std::string cmd = argv[1];
  if (cmd == "params") {
    std::map<std::string, XmlRpc::XmlRpcValue> p;
    loadParams(&p, argv[2]);
    printf("%d %d %.9g %.9g %d %s %s\n", (int)p["hz"], (int)p["num_timesteps"], (double)p["gamma"], (double)p["max_throttle"],
           (int)(bool)p["l1_cost"], ((std::string)p["model_path"]).c_str(), ((std::string)p["map_path"]).c_str());
    return 0;
  }
  if (cmd == "nn_step" || cmd == "bf_step") {
    float2 rng[2] = {make_float2(-.99, .99), make_float2(-.99, .65)};
    Eigen::MatrixXf s(7, 1), u(2, 1);
    for (int i = 0; i < 7; i++) s(i) = (float)atof(argv[3 + i]);
    for (int i = 0; i < 2; i++) u(i) = (float)atof(argv[10 + i]);
    if (cmd == "nn_step") {
      NeuralNetModel<7, 2, 3, 6, 32, 32, 4> m(0.02f, rng);
      m.loadParams(argv[2]);
      Eigen::MatrixXf s0 = s, u0 = u;
      m.updateState(s, u);
      for (int i = 0; i < 7; i++) printf("%.9g ", s(i));
      printf("%.9g %.9g\n", u(0), u(1));
      m.enforceConstraints(s0, u0);
      m.computeGrad(s0, u0);
      for (int r = 0; r < 7; r++) { for (int c = 0; c < 9; c++) printf("%.9g ", m.jac_(r, c)); printf("\n"); }
      // finite-difference Jacobian through the same host twin
      for (int c = 0; c < 9; c++) {
        Eigen::MatrixXf sp = s0, sm = s0, up = u0, um = u0;
        const float h = 1e-3f;
        if (c < 7) { sp(c) += h; sm(c) -= h; } else { up(c - 7) += h; um(c - 7) -= h; }
        m.computeKinematics(sp); m.computeDynamics(sp, up);
        Eigen::Matrix<float, 7, 1> fp = m.state_der_;
        m.computeKinematics(sm); m.computeDynamics(sm, um);
        for (int r = 0; r < 7; r++) printf("%.9g ", (fp(r) - m.state_der_(r)) / (2 * h));
        printf("\n");
      }
    } else {
      GeneralizedLinear<CarBasisFuncs, 7, 2, 25, CarKinematics, 3> m(0.02f, rng);
      m.loadParams(argv[2]);
      m.updateState(s, u);
      for (int i = 0; i < 7; i++) printf("%.9g ", s(i));
      printf("%.9g %.9g\n", u(0), u(1));
    }
    return 0;
  }
  if (cmd == "costmap") {
    std::map<std::string, XmlRpc::XmlRpcValue> p;
    loadParams(&p, argv[2]);
    MPPICosts c(&p);
    printf("%d %d %.9g %.9g %.9g %.9g %.9g %.9g %lu %lu\n", c.width(), c.height(), c.params_.r_c1.x, c.params_.r_c2.y, c.params_.trs.x,
           c.params_.trs.y, c.params_.desired_speed, c.params_.boundary_threshold, c.paramsVersion(), c.mapVersion());
    return 0;
  }
  if (cmd == "slide") {
    // slideControlAndStateSeq(stride) of the drop-in MPPIController template (host logic only; without a GPU the context
    // creation fails, which the class reports and survives): slide stride init_u0 init_u1 hist[4] U[2T] -> hist[4] U[2T]
    const int stride = atoi(argv[2]), T = (argc - 9) / 2;
    float init_u[2] = {(float)atof(argv[3]), (float)atof(argv[4])}, nu[2] = {0.275f, 0.3f};
    float2 rng[2] = {make_float2(-.99, .99), make_float2(-.99, .65)};
    typedef NeuralNetModel<7, 2, 3, 6, 32, 32, 4> Model;
    Model model(0.02f, rng);
    MPPICosts costs(4, 4);
    MPPIController<Model, MPPICosts, 1920, 8, 16> ctl(&model, &costs, nu, init_u, 50, T, stride, 0.15f, 1);
    for (int i = 0; i < 4; i++) ctl.control_hist_[i] = (float)atof(argv[5 + i]);
    for (int i = 0; i < 2 * T; i++) ctl.U_[i] = (float)atof(argv[9 + i]);
    ctl.slideControlAndStateSeq(stride);
    for (int i = 0; i < 4; i++) printf("%.9g ", ctl.control_hist_[i]);
    for (int i = 0; i < 2 * T; i++) printf("%.9g ", ctl.U_[i]);
    printf("\n");
    return 0;
  }
  if (cmd == "ddp_toy") {
    ToyLinear toy;
    ModelWrapperDDP<ToyLinear> dyn(&toy);
    const int H = 30;
    Eigen::Matrix<float, 2, 2> Q, Qf; Eigen::Matrix<float, 1, 1> R; Eigen::Matrix<float, 1, 1> lo, hi;
    Q.setZero(); Q(0, 0) = 1.0f; Q(1, 1) = 0.1f; Qf.setZero(); Qf(0, 0) = 2.0f; R(0, 0) = 0.5f; lo(0) = -100; hi(0) = 100;
    std::vector<float> tx(2 * H, 0.0f), tu(H, 0.0f);
    Eigen::MatrixXf x0(2, 1); x0(0) = 1.0f; x0(1) = -0.5f;
    auto res = ddp_feedback_gains(dyn, x0, tx, tu, H, 0.05f, Q, Qf, R, lo, hi);
    for (int k = 0; k < H; k++) printf("%.9g %.9g %.9g\n", res.feedback_gain[k](0, 0), res.feedback_gain[k](0, 1), res.feedforward_gain(0, k));
    return 0;
  }
  return 1;
}
