// mppi_controller.cuh -- MPPIController<DYNAMICS_T, COSTS_T, ROLLOUTS, BDIM_X, BDIM_Y>: drop-in host
// class with the reference's public surface (PI/mppi_controller.cuh:52-217).  It owns the host state
// exactly as the reference does (U_, control history, state / control solutions) and forwards the hot
// path -- noise, rollouts, importance weighting, smoothing, nominal trajectory -- to one call of the
// C ABI (mppi_compute_control, include/mppi_b200.h).  BDIM_X / BDIM_Y are accepted for source
// compatibility; the sm_100a kernels choose their own launch geometry.
//
// Extensions (not in the reference): updateControlNoise(), setNoise()/useSampler(), getRolloutCosts(),
// getBaseline()/getNormalizer(), context().
#ifndef MPPI_CONTROLLER_CUH_
#define MPPI_CONTROLLER_CUH_
#include <cstdio>
#include <vector>

#include <Eigen/Dense>
#include <cuda_runtime.h>

#include "../../mppi_b200.h"
#include "../ddp/ddp_feedback.h"
#include "gpu_err_chk.h"
#include "managed.cuh"

namespace autorally_control {

template <class DYNAMICS_T, class COSTS_T, int ROLLOUTS = 2560, int BDIM_X = 64, int BDIM_Y = 1>
class MPPIController {
 public:
  static const int BLOCKSIZE_WRX = 64;
  static const int NUM_ROLLOUTS = (ROLLOUTS / BLOCKSIZE_WRX) * BLOCKSIZE_WRX;  // multiple of 64 (:58-60)
  static const int BLOCKSIZE_X = BDIM_X;
  static const int BLOCKSIZE_Y = BDIM_Y;
  static const int STATE_DIM = DYNAMICS_T::STATE_DIM;
  static const int CONTROL_DIM = DYNAMICS_T::CONTROL_DIM;

  cudaStream_t stream_;
  int numTimesteps_;
  int hz_;
  int optimizationStride_;
  DYNAMICS_T *model_;  ///< not owned
  COSTS_T *costs_;     ///< not owned

  // feedback gains around the MPPI solution (PI/mppi_controller.cuh:75-86)
  ModelWrapperDDP<DYNAMICS_T> *ddp_model_ = nullptr;
  Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, DYNAMICS_T::STATE_DIM> Q_, Qf_;
  Eigen::Matrix<float, DYNAMICS_T::CONTROL_DIM, DYNAMICS_T::CONTROL_DIM> R_;
  Eigen::Matrix<float, DYNAMICS_T::CONTROL_DIM, 1> U_MIN_, U_MAX_;
  OptimizerResult<ModelWrapperDDP<DYNAMICS_T> > result_;

  MPPIController(DYNAMICS_T *model, COSTS_T *costs, float *exploration_var, float *init_control, int hz,
                 int num_timesteps, int optimization_stride, float gamma, int num_iters, cudaStream_t stream = 0)
      : stream_(stream), numTimesteps_(num_timesteps), hz_(hz), optimizationStride_(optimization_stride),
        model_(model), costs_(costs), num_iters_(num_iters), gamma_(gamma) {
    setCudaStream(stream);
    nu_.assign(exploration_var, exploration_var + CONTROL_DIM);
    init_u_.assign(init_control, init_control + CONTROL_DIM);
    control_hist_.assign(2 * CONTROL_DIM, 0);
    state_solution_.assign(numTimesteps_ * STATE_DIM, 0);
    control_solution_.assign(numTimesteps_ * CONTROL_DIM, 0);
    du_.resize(numTimesteps_ * CONTROL_DIM);
    U_.resize(numTimesteps_ * CONTROL_DIM);
    traj_costs_.resize(NUM_ROLLOUTS);
    allocateCudaMem();
    initDDP();
    resetControls();
  }

  ~MPPIController() { deallocateCudaMem(); delete ddp_model_; }

  void setCudaStream(cudaStream_t stream) {
    stream_ = stream;
    model_->bindToStream(stream_);
    costs_->bindToStream(stream_);
  }

  /// Creates the device context (state, U, noise / sampled-control buffer, costs; PI/mppi_controller.cu:379-387).
  void allocateCudaMem() {
    if (ctx_) return;
    mppi_config cfg;
    mppi_config_default(&cfg);
    cfg.dynamics = DYNAMICS_T::MPPI_DYNAMICS_KIND;
    cfg.num_rollouts = NUM_ROLLOUTS;
    cfg.num_timesteps = numTimesteps_;
    cfg.hz = hz_;
    cfg.optimization_stride = optimizationStride_;
    cfg.gamma = gamma_;
    cfg.num_iters = num_iters_;
    cfg.bdim_x = BDIM_X;
    cfg.bdim_y = BDIM_Y;
    cfg.seed = 1234ULL;  // the reference seeds cuRAND with 1234 (:331)
    HANDLE_ERROR(mppi_create(&cfg, &ctx_));
    model_version_ = costs_version_ = map_version_ = ~0ul;
  }

  /// Idempotent; unlike the reference it neither destroys the caller's stream nor frees memory shared
  /// with a second controller (PI/mppi_controller.cu:389-400 does both).
  void deallocateCudaMem() {
    if (ctx_) { mppi_destroy(ctx_); ctx_ = nullptr; }
  }

  void initDDP() {
    delete ddp_model_;
    ddp_model_ = new ModelWrapperDDP<DYNAMICS_T>(model_);
    Q_.setZero(); Qf_.setZero(); R_.setZero();
    const float q[7] = {0.5f, 0.5f, 0.25f, 0.0f, 0.05f, 0.01f, 0.01f};  // :410-417
    for (int i = 0; i < STATE_DIM && i < 7; i++) Q_(i, i) = q[i];
    for (int i = 0; i < CONTROL_DIM; i++) R_(i, i) = 10.0f;
    for (int i = 0; i < CONTROL_DIM; i++) { U_MIN_(i) = model_->control_rngs_[i].x; U_MAX_(i) = model_->control_rngs_[i].y; }
  }

  /// Time-varying feedback gains around the nominal trajectory (PI/mppi_controller.cu:402-445).
  void computeFeedbackGains(Eigen::MatrixXf state) {
    for (int i = 0; i < CONTROL_DIM; i++) { U_MIN_(i) = model_->control_rngs_[i].x; U_MAX_(i) = model_->control_rngs_[i].y; }
    result_ = ddp_feedback_gains(*ddp_model_, state, state_solution_, control_solution_, numTimesteps_, 1.0f / hz_, Q_, Qf_, R_, U_MIN_, U_MAX_);
  }

  OptimizerResult<ModelWrapperDDP<DYNAMICS_T> > getFeedbackGains() { return result_; }

  void resetControls() {
    for (int i = 0; i < numTimesteps_; i++)
      for (int j = 0; j < CONTROL_DIM; j++) U_[i * CONTROL_DIM + j] = init_u_[j];
  }

  void cutThrottle() {
    costs_->params_.desired_speed = 0.0;
    model_->control_rngs_[1].y = 0.0;
    costs_->paramsToDevice();
    model_->paramsToDevice();
  }

  /// Host Savitzky-Golay pass over U_ (PI/mppi_controller.cu:468-499); computeControl runs the same filter on the device.
  void savitskyGolay() {
    const float filt[5] = {-3.0f / 35.0f, 12.0f / 35.0f, 17.0f / 35.0f, 12.0f / 35.0f, -3.0f / 35.0f};
    const int T = numTimesteps_;
    std::vector<float> padded((T + 4) * CONTROL_DIM);
    for (int i = 0; i < T + 4; i++)
      for (int j = 0; j < CONTROL_DIM; j++)
        padded[i * CONTROL_DIM + j] = i < 2 ? control_hist_[CONTROL_DIM * i + j]
                                           : (i < T + 2 ? U_[CONTROL_DIM * (i - 2) + j] : U_[CONTROL_DIM * (T - 1) + j]);
    for (int i = 0; i < T; i++)
      for (int j = 0; j < CONTROL_DIM; j++) {
        float acc = 0.0f;
        for (int k = 0; k < 5; k++) acc += filt[k] * padded[(i + k) * CONTROL_DIM + j];
        U_[CONTROL_DIM * i + j] = acc;
      }
  }

  /// Host nominal rollout with the model's host twin (PI/mppi_controller.cu:501-519).
  void computeNominalTraj(Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> state) {
    Eigen::MatrixXf s(STATE_DIM, 1), u(CONTROL_DIM, 1);
    for (int j = 0; j < STATE_DIM; j++) s(j) = state(j);
    for (int i = 0; i < numTimesteps_; i++) {
      for (int j = 0; j < STATE_DIM; j++) state_solution_[i * STATE_DIM + j] = s(j);
      for (int j = 0; j < CONTROL_DIM; j++) u(j) = U_[CONTROL_DIM * i + j];
      model_->updateState(s, u);
      for (int j = 0; j < CONTROL_DIM; j++) control_solution_[CONTROL_DIM * i + j] = u(j);
    }
  }

  void slideControlAndStateSeq(int stride) {
    slideControlSeq(stride);
    slideStateSeq(stride);
  }

  /// The reference indexes column 1 of a 7x1 vector here (out of bounds, :574); the intended
  /// meaning -- remember `state` as the head of the state sequence -- is implemented.
  void setState(Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> state) {
    for (int i = 0; i < STATE_DIM; i++) state_solution_[i] = state(i);
  }
  void setStateSequence(std::vector<float> state_seq) { state_solution_ = state_seq; }
  void setControlSequence(std::vector<float> control_seq) { control_solution_ = control_seq; }

  /// Plans from the head of the previously computed state sequence (:588-597).
  void computeControl() {
    Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> expected;
    for (int i = 0; i < STATE_DIM; i++) expected(i) = state_solution_[i];
    computeControl(expected);
  }

  /// The hot path (:600-675): one C-ABI call.
  void computeControl(Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> state) {
    if (!ctx_) return;
    syncParams();
    float st[DYNAMICS_T::STATE_DIM];
    for (int i = 0; i < STATE_DIM; i++) st[i] = state(i);
    mppi_result res;
    int rc = mppi_compute_control(ctx_, st, U_.data(), control_hist_.data(), state_solution_.data(), control_solution_.data(), &res);
    HANDLE_ERROR(rc);
    if (rc == 0) {
      normalizer_ = res.normalizer;
      trajectory_cost_ = res.trajectory_cost;
      baseline_ = res.baseline;
    }
  }

  /// Extension: the two halves of computeControl(state).  computeControlAsync enqueues the whole hot path on this
  /// controller's stream and returns; waitControl blocks and stores the results.  Two controllers started back to back
  /// run concurrently on the GPU (run_control_loop.cuh uses this for the actual- / predicted-state pair).
  void computeControlAsync(Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> state) {
    if (!ctx_) return;
    syncParams();
    float st[DYNAMICS_T::STATE_DIM];
    for (int i = 0; i < STATE_DIM; i++) st[i] = state(i);
    pending_rc_ = mppi_compute_control_async(ctx_, st, U_.data(), control_hist_.data());
    HANDLE_ERROR(pending_rc_);
  }
  void computeControlAsync() {
    Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> expected;
    for (int i = 0; i < STATE_DIM; i++) expected(i) = state_solution_[i];
    computeControlAsync(expected);
  }
  void waitControl() {
    if (!ctx_ || pending_rc_ != 0) return;
    mppi_result res;
    const int rc = mppi_compute_control_wait(ctx_, U_.data(), state_solution_.data(), control_solution_.data(), &res);
    HANDLE_ERROR(rc);
    if (rc == 0) { normalizer_ = res.normalizer; trajectory_cost_ = res.trajectory_cost; baseline_ = res.baseline; }
  }

  std::vector<float> getControlSeq() { return control_solution_; }
  std::vector<float> getStateSeq() { return state_solution_; }
  float getComputedTrajectoryCost() { return trajectory_cost_; }

  // ---- extensions ----
  /// north_star's updateControlNoise (no such symbol in the reference): overwrite the exploration
  /// standard deviations nu_ (what the ctor's exploration_var sets, PI/mppi_controller.cu:345,357).
  void updateControlNoise(const float *exploration_std) {
    nu_.assign(exploration_std, exploration_std + CONTROL_DIM);
    if (ctx_) HANDLE_ERROR(mppi_set_exploration_std(ctx_, nu_.data()));
  }
  /// Inject N(0,1) noise [num_iters][NUM_ROLLOUTS][T][2] (parity runs); useSampler() returns to Philox.
  void setNoise(const float *eps, size_t count) { if (ctx_) HANDLE_ERROR(mppi_set_noise(ctx_, eps, count)); }
  void useSampler() { if (ctx_) HANDLE_ERROR(mppi_use_sampler(ctx_)); }
  std::vector<float> getRolloutCosts() {
    if (ctx_) HANDLE_ERROR(mppi_get_rollout_costs(ctx_, traj_costs_.data()));
    return traj_costs_;
  }
  std::vector<float> getControlSequenceU() { return U_; }
  float getBaseline() const { return baseline_; }
  float getNormalizer() const { return normalizer_; }
  mppi_ctx *context() { return ctx_; }

 private:
  void syncParams() {
    if (model_->paramsVersion() != model_version_) {
      HANDLE_ERROR(model_->uploadTo(ctx_));
      model_version_ = model_->paramsVersion();
    }
    // ranges, flags and cost scalars are host-side in the context (kernel parameters): always current
    const float rng[4] = {model_->control_rngs_[0].x, model_->control_rngs_[0].y, model_->control_rngs_[1].x, model_->control_rngs_[1].y};
    HANDLE_ERROR(mppi_set_control_ranges(ctx_, rng));
    HANDLE_ERROR(mppi_set_negate_yaw_der(ctx_, model_->negate_yaw_der ? 1 : 0));
    HANDLE_ERROR(costs_->uploadParamsTo(ctx_));
    if (costs_->mapVersion() != map_version_) {
      HANDLE_ERROR(costs_->uploadMapTo(ctx_));
      map_version_ = costs_->mapVersion();
    }
    HANDLE_ERROR(mppi_set_exploration_std(ctx_, nu_.data()));
    HANDLE_ERROR(mppi_set_gamma(ctx_, gamma_));
  }

  /// PI/mppi_controller.cu:527-554, including the stride != 1 branch that indexes the flat U_.
  void slideControlSeq(int stride) {
    if (stride == 1) {
      control_hist_[0] = control_hist_[2];
      control_hist_[1] = control_hist_[3];
      control_hist_[2] = U_[0];
      control_hist_[3] = U_[1];
    } else {
      const int t = stride - 2;
      for (int i = 0; i < 4; i++) control_hist_[i] = U_[t + i];
    }
    for (int i = 0; i < numTimesteps_ - stride; i++)
      for (int j = 0; j < CONTROL_DIM; j++) U_[i * CONTROL_DIM + j] = U_[(i + stride) * CONTROL_DIM + j];
    for (int j = 1; j <= stride; j++)
      for (int i = 0; i < CONTROL_DIM; i++) U_[(numTimesteps_ - j) * CONTROL_DIM + i] = init_u_[i];
  }

  void slideStateSeq(int stride) {
    for (int i = 0; i < numTimesteps_ - stride; i++)
      for (int j = 0; j < STATE_DIM; j++) state_solution_[i * STATE_DIM + j] = state_solution_[(i + stride) * STATE_DIM + j];
  }

  int num_iters_;
  float gamma_;
  float normalizer_ = 0, trajectory_cost_ = 0, baseline_ = 0;
  int pending_rc_ = 0;
  std::vector<float> traj_costs_, state_solution_, control_solution_, control_hist_, U_, du_, nu_, init_u_;
  mppi_ctx *ctx_ = nullptr;
  unsigned long model_version_ = ~0ul, costs_version_ = ~0ul, map_version_ = ~0ul;
};

}  // namespace autorally_control
#endif
