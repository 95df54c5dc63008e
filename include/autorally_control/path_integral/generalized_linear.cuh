// generalized_linear.cuh -- GeneralizedLinear<BF, S_DIM, C_DIM, BF_DIM, K_FUNC, K_DIM>: host side of the
// basis-function dynamics model, state_der[dyn] = theta (DYN x BF_DIM) * phi(s, u)
// (API of PI/generalized_linear.cuh:46-112).  Device evaluation: libmppi_b200.so (CarBasisDyn).
#ifndef GENERALIZED_LINEAR_CUH_
#define GENERALIZED_LINEAR_CUH_
#include <cfloat>
#include <cstdio>
#include <string>
#include <vector>

#include <Eigen/Dense>

#include "../../mppi_b200.h"
#include "gpu_err_chk.h"
#include "managed.cuh"
#include "meta_math.h"
#include "npz_io.h"
#include "param_getter.h"

namespace autorally_control {

template <class BF, int S_DIM, int C_DIM, int BF_DIM, class K_FUNC, int K_DIM>
class GeneralizedLinear : public Managed {
 public:
  float2 *control_rngs_;

  static const int NUM_BFS = BF_DIM;
  static const int STATE_DIM = S_DIM;
  static const int CONTROL_DIM = C_DIM;
  static const int DYNAMICS_DIM = STATE_DIM - K_DIM;
  static const int SHARED_MEM_REQUEST_GRD = DYNAMICS_DIM * BF_DIM;
  static const int SHARED_MEM_REQUEST_BLK = 0;
  static const int MPPI_DYNAMICS_KIND = MPPI_DYNAMICS_BF;

  typedef Eigen::Matrix<float, DYNAMICS_DIM, NUM_BFS, Eigen::RowMajor> ThetaMatrix;

  Eigen::Matrix<float, STATE_DIM, 1> state_der_;
  bool negate_yaw_der = true;  // kept for API parity; this model always negates (PI/generalized_linear.cu:222)

  GeneralizedLinear(ThetaMatrix theta, float delta_t, float2 *control_rngs = NULL) : dt_(delta_t) {
    init_ranges(control_rngs);
    setParams(theta);
  }
  GeneralizedLinear(float delta_t, float2 *control_rngs = NULL) : dt_(delta_t) { init_ranges(control_rngs); }

  void setParams(ThetaMatrix theta) { theta_ = theta; paramsToDevice(); }

  /// npz key "W": DYN x NUM_BFS, float64 (PI/generalized_linear.cu:92-108)
  void loadParams(std::string model_path) {
    if (!fileExists(model_path)) {
      fprintf(stderr, "Could not load generalized linear model at path: %s\n", model_path.c_str());
      return;
    }
    npz::Archive dict = npz::load(model_path);
    const npz::Array &w = dict.at("W");
    if ((int)w.num_vals() != DYNAMICS_DIM * NUM_BFS) {
      fprintf(stderr, "Basis function model %s has the wrong shape\n", model_path.c_str());
      return;
    }
    ThetaMatrix theta;
    for (int i = 0; i < DYNAMICS_DIM; i++)
      for (int j = 0; j < NUM_BFS; j++) theta(i, j) = (float)w.at((size_t)i * NUM_BFS + j);
    setParams(theta);
  }

  void paramsToDevice() { params_version_++; }
  void freeCudaMem() {}
  void updateModel(std::vector<int>, std::vector<float>) {}

  void enforceConstraints(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    (void)state;
    for (int i = 0; i < CONTROL_DIM; i++) {
      if (control(i) < control_rngs_[i].x) control(i) = control_rngs_[i].x;
      else if (control(i) > control_rngs_[i].y) control(i) = control_rngs_[i].y;
    }
  }

  void computeKinematics(Eigen::MatrixXf &state) {
    float der[3];
    kinematics_.computeKinematics(state.data(), der);
    for (int i = 0; i < 3; i++) state_der_(i) = der[i];
  }

  void computeDynamics(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    float phi[NUM_BFS];
    for (int i = 0; i < NUM_BFS; i++) phi[i] = basis_.basisFuncX(i, state.data(), control.data());
    for (int j = 0; j < DYNAMICS_DIM; j++) {
      float t = 0.0f;
      for (int i = 0; i < NUM_BFS; i++) t += theta_(j, i) * phi[i];
      state_der_(j + (STATE_DIM - DYNAMICS_DIM)) = t;
    }
  }

  void updateState(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    enforceConstraints(state, control);
    computeKinematics(state);
    computeDynamics(state, control);
    for (int i = 0; i < STATE_DIM; i++) { state(i) += state_der_(i) * dt_; state_der_(i) = 0; }
  }

  float dt() const { return dt_; }
  unsigned long paramsVersion() const { return params_version_; }
  int uploadTo(mppi_ctx *ctx) const {
    float th[DYNAMICS_DIM * NUM_BFS];
    for (int i = 0; i < DYNAMICS_DIM; i++)
      for (int j = 0; j < NUM_BFS; j++) th[i * NUM_BFS + j] = theta_(i, j);
    return mppi_set_bf_params(ctx, th);
  }

 protected:
  void init_ranges(float2 *control_rngs) {
    if (control_rngs == NULL) {
      control_rngs_ = new float2[CONTROL_DIM];
      for (int i = 0; i < CONTROL_DIM; i++) { control_rngs_[i].x = -FLT_MAX; control_rngs_[i].y = FLT_MAX; }
    } else {
      control_rngs_ = control_rngs;
    }
  }
  float dt_;
  BF basis_;
  K_FUNC kinematics_;
  ThetaMatrix theta_;
  unsigned long params_version_ = 0;
};

}  // namespace autorally_control
#endif
