// neural_net_model.cuh -- NeuralNetModel<S_DIM, C_DIM, K_DIM, layers...>: host side of the MLP dynamics
// model with the reference's API (PI/neural_net_model.cuh:48-132).  The device forward pass is a
// pre-instantiated kernel of libmppi_b200.so; this class owns the weights, loads them from the same
// .npz files, packs them exactly as paramsToDevice does (PI/neural_net_model.cu:120-141) and hands them
// to a controller's context through the C ABI (mppi_set_nn_params).  It also carries the host twin
// (updateState / computeDynamics / computeGrad) the control loop and DDP use.
#ifndef NEURAL_NET_MODEL_CUH_
#define NEURAL_NET_MODEL_CUH_
#include <cfloat>
#include <cmath>
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

template <int S_DIM, int C_DIM, int K_DIM, int... layer_args>
class NeuralNetModel : public Managed {
 public:
  static const int STATE_DIM = S_DIM;
  static const int CONTROL_DIM = C_DIM;
  static const int DYNAMICS_DIM = STATE_DIM - K_DIM;
  static const int NUM_LAYERS = layer_counter(layer_args...);
  static const int PRIME_PADDING = 1;
  static const int LARGEST_LAYER = neuron_counter(layer_args...) + PRIME_PADDING;
  static const int NUM_PARAMS = param_counter(layer_args...);
  static const int SHARED_MEM_REQUEST_GRD = 0;
  static const int SHARED_MEM_REQUEST_BLK = 2 * LARGEST_LAYER;
  static const int MPPI_DYNAMICS_KIND = MPPI_DYNAMICS_NN;

  typedef Eigen::Matrix<float, -1, -1, Eigen::RowMajor> RowMatrix;

  float2 *control_rngs_;                            ///< per-control (min, max)
  Eigen::Matrix<float, STATE_DIM, 1> state_der_;    ///< last state derivative of the host twin
  Eigen::MatrixXf ip_delta_;
  Eigen::Matrix<float, STATE_DIM, STATE_DIM + CONTROL_DIM> jac_;  ///< d(state_der)/d(state, control)
  bool negate_yaw_der = true;

  NeuralNetModel(float delta_t, float2 *control_rngs = NULL) : dt_(delta_t) {
    if (control_rngs == NULL) {
      control_rngs_ = new float2[CONTROL_DIM];
      for (int i = 0; i < CONTROL_DIM; i++) { control_rngs_[i].x = -FLT_MAX; control_rngs_[i].y = FLT_MAX; }
    } else {
      control_rngs_ = control_rngs;
    }
    weights_.resize(NUM_LAYERS - 1);
    biases_.resize(NUM_LAYERS - 1);
    weighted_in_.resize(NUM_LAYERS - 1);
    for (int l = 0; l + 1 < NUM_LAYERS; l++) {
      weights_[l] = RowMatrix::Zero(net_structure_[l + 1], net_structure_[l]);
      biases_[l] = RowMatrix::Zero(net_structure_[l + 1], 1);
      weighted_in_[l] = Eigen::MatrixXf::Zero(net_structure_[l + 1], 1);
    }
    net_params_.assign(NUM_PARAMS, 0.0f);
  }

  ~NeuralNetModel() {}

  /// npz keys dynamics_W{i} (out x in, row-major) and dynamics_b{i}, float64 (PI/neural_net_model.cu:73-106)
  void loadParams(std::string model_path) {
    if (!fileExists(model_path)) {
      fprintf(stderr, "Could not load neural net model at path: %s\n", model_path.c_str());
      return;
    }
    npz::Archive dict = npz::load(model_path);
    for (int l = 1; l < NUM_LAYERS; l++) {
      const npz::Array &w = dict.at("dynamics_W" + std::to_string(l));
      const npz::Array &b = dict.at("dynamics_b" + std::to_string(l));
      const int nout = net_structure_[l], nin = net_structure_[l - 1];
      if ((int)w.num_vals() != nout * nin || (int)b.num_vals() != nout) {
        fprintf(stderr, "Neural net model %s does not match the compiled structure\n", model_path.c_str());
        return;
      }
      for (int j = 0; j < nout; j++) {
        for (int k = 0; k < nin; k++) weights_[l - 1](j, k) = (float)w.at((size_t)j * nin + k);
        biases_[l - 1](j, 0) = (float)b.at(j);
      }
    }
    paramsToDevice();
  }

  void setParams(RowMatrix *weights, RowMatrix *biases) {
    for (int l = 0; l + 1 < NUM_LAYERS; l++) { weights_[l] = weights[l]; biases_[l] = biases[l]; }
    paramsToDevice();
  }

  /// Packs [W1|b1|W2|b2|...] and marks the parameters dirty; every controller bound to this model
  /// re-uploads them (asynchronously, on its own stream) at its next computeControl.
  void paramsToDevice() {
    int stride = 0;
    for (int l = 0; l + 1 < NUM_LAYERS; l++) {
      const int nin = net_structure_[l], nout = net_structure_[l + 1];
      stride_idcs_[2 * l] = stride;
      for (int j = 0; j < nout; j++)
        for (int k = 0; k < nin; k++) net_params_[stride + j * nin + k] = weights_[l](j, k);
      stride += nin * nout;
      stride_idcs_[2 * l + 1] = stride;
      for (int j = 0; j < nout; j++) net_params_[stride + j] = biases_[l](j, 0);
      stride += nout;
    }
    stride_idcs_[2 * NUM_LAYERS] = stride;
    params_version_++;
  }

  /// Hot swap from a flattened message: ALL weights first, then ALL biases (PI/neural_net_model.cu:152-180).
  void updateModel(std::vector<int> description, std::vector<float> data) {
    for (size_t i = 0; i < description.size(); i++)
      if ((int)i >= NUM_LAYERS || description[i] != net_structure_[i]) return;
    size_t need = 0;
    for (int l = 0; l + 1 < NUM_LAYERS; l++) need += (size_t)(net_structure_[l] + 1) * net_structure_[l + 1];
    if (data.size() < need) return;
    size_t off = 0;
    for (int l = 0; l + 1 < NUM_LAYERS; l++) {
      const int nin = net_structure_[l], nout = net_structure_[l + 1];
      for (int j = 0; j < nout; j++)
        for (int k = 0; k < nin; k++) weights_[l](j, k) = data[off + (size_t)j * nin + k];
      off += (size_t)nin * nout;
    }
    for (int l = 0; l + 1 < NUM_LAYERS; l++) {
      for (int j = 0; j < net_structure_[l + 1]; j++) biases_[l](j, 0) = data[off + j];
      off += net_structure_[l + 1];
    }
    // The reference re-uploads on every computeControl (model_->paramsToDevice(), PI/mppi_controller.cu:603-604); here the
    // controllers upload only when the parameter version changed, so the swap has to repack and bump it now.
    paramsToDevice();
  }

  void printParamVec() {
    for (int i = 0; i < NUM_PARAMS; i++) printf("Buffer Idx: %d, Value: %f \n", i, net_params_[i]);
  }

  void freeCudaMem() {}  // device storage belongs to the controllers' contexts

  // ---- host twin (PI/neural_net_model.cu:191-288) ----
  void enforceConstraints(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    (void)state;
    for (int i = 0; i < CONTROL_DIM; i++) {
      if (control(i) < control_rngs_[i].x) control(i) = control_rngs_[i].x;
      else if (control(i) > control_rngs_[i].y) control(i) = control_rngs_[i].y;
    }
  }

  void computeKinematics(Eigen::MatrixXf &state) {
    const float c = cosf(state(2)), s = sinf(state(2));
    state_der_(0) = c * state(4) - s * state(5);
    state_der_(1) = s * state(4) + c * state(5);
    state_der_(2) = negate_yaw_der ? -state(6) : state(6);
  }

  void computeDynamics(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    std::vector<float> acts(net_structure_[0]);
    for (int i = 0; i < DYNAMICS_DIM; i++) acts[i] = state(i + (STATE_DIM - DYNAMICS_DIM));
    for (int i = 0; i < CONTROL_DIM; i++) acts[DYNAMICS_DIM + i] = control(i);
    for (int l = 0; l + 1 < NUM_LAYERS; l++) {
      const int nin = net_structure_[l], nout = net_structure_[l + 1];
      std::vector<float> next(nout);
      for (int j = 0; j < nout; j++) {
        float t = 0.0f;
        for (int k = 0; k < nin; k++) t += weights_[l](j, k) * acts[k];
        t += biases_[l](j, 0);
        weighted_in_[l](j, 0) = t;
        next[j] = (l + 2 < NUM_LAYERS) ? tanhf(t) : t;
      }
      acts.swap(next);
    }
    for (int i = 0; i < DYNAMICS_DIM; i++) state_der_(i + (STATE_DIM - DYNAMICS_DIM)) = acts[i];
  }

  /// Analytic Jacobian of the state derivative w.r.t. (state, control) by back-propagation
  /// (PI/neural_net_model.cu:233-264): kinematic rows + d(MLP)/d(inputs) in the lower-right block.
  void computeGrad(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    jac_.setZero();
    const float c = cosf(state(2)), s = sinf(state(2));
    jac_(0, 2) = -s * state(4) - c * state(5); jac_(0, 4) = c; jac_(0, 5) = -s;
    jac_(1, 2) = c * state(4) - s * state(5);  jac_(1, 4) = s; jac_(1, 5) = c;
    jac_(2, 6) = -1.0f;  // as in the reference, independent of negate_yaw_der
    computeDynamics(state, control);
    // delta: [width of current layer] x DYNAMICS_DIM, starting from the identity at the output
    Eigen::MatrixXf delta = Eigen::MatrixXf::Identity(DYNAMICS_DIM, DYNAMICS_DIM);
    for (int l = NUM_LAYERS - 2; l > 0; l--) {
      Eigen::MatrixXf next = weights_[l].transpose() * delta;  // [width_l x D]
      for (int j = 0; j < net_structure_[l]; j++) {
        const float th = tanhf(weighted_in_[l - 1](j, 0));
        const float dact = 1.0f - th * th;
        for (int d = 0; d < DYNAMICS_DIM; d++) next(j, d) *= dact;
      }
      delta = next;
    }
    ip_delta_ = weights_[0].transpose() * delta;  // [(DYNAMICS_DIM + CONTROL_DIM) x DYNAMICS_DIM]
    for (int d = 0; d < DYNAMICS_DIM; d++)
      for (int i = 0; i < DYNAMICS_DIM + CONTROL_DIM; i++)
        jac_(K_DIM + d, K_DIM + i) += ip_delta_(i, d);
  }

  void updateState(Eigen::MatrixXf &state, Eigen::MatrixXf &control) {
    enforceConstraints(state, control);
    computeKinematics(state);
    computeDynamics(state, control);
    for (int i = 0; i < STATE_DIM; i++) { state(i) += state_der_(i) * dt_; state_der_(i) = 0; }
  }

  // ---- bridge to the C ABI (used by MPPIController) ----
  float dt() const { return dt_; }
  unsigned long paramsVersion() const { return params_version_; }
  const float *packedParams() const { return net_params_.data(); }
  const int *netStructure() const { return net_structure_; }
  int uploadTo(mppi_ctx *ctx) const {
    int rc = mppi_set_nn_params(ctx, net_params_.data(), net_structure_, NUM_LAYERS);
    if (rc) return rc;
    return mppi_set_negate_yaw_der(ctx, negate_yaw_der ? 1 : 0);
  }

 private:
  float dt_;
  int net_structure_[NUM_LAYERS] = {layer_args...};
  int stride_idcs_[NUM_LAYERS * 2 + 1] = {0};
  std::vector<RowMatrix> weights_, biases_;
  std::vector<Eigen::MatrixXf> weighted_in_;
  std::vector<float> net_params_;
  unsigned long params_version_ = 0;
};

template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::STATE_DIM;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::CONTROL_DIM;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::DYNAMICS_DIM;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::NUM_LAYERS;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::LARGEST_LAYER;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::NUM_PARAMS;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::SHARED_MEM_REQUEST_GRD;
template <int S_DIM, int C_DIM, int K_DIM, int... layer_args> const int NeuralNetModel<S_DIM, C_DIM, K_DIM, layer_args...>::SHARED_MEM_REQUEST_BLK;

}  // namespace autorally_control
#endif
