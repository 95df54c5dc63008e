// ref_harness.cu -- C ABI around the REFERENCE's own MPPI controller, compiled from the sources where they
// lie under /root/reference (never copied into this repo) by oracle/refbuild.py into
// oracle/_ref/libautorally_ref.so.
//
// TEST INFRASTRUCTURE ONLY (checker and "reference" CPU/GPU baseline): nothing in autorally_b200/ or include/
// loads this library.  What is the reference's and what is not:
//   * reference, unmodified: MPPIController<...>::computeControl and everything it calls -- rolloutKernel,
//     normExpKernel, weightedReductionKernel, the NeuralNetModel / GeneralizedLinear / CarBasisFuncs / MPPICosts
//     device members, the host min / normaliser loops, savitskyGolay, computeNominalTraj, slideControlAndStateSeq
//     (PI/mppi_controller.{cuh,cu}, PI/neural_net_model.{cuh,cu}, PI/generalized_linear.{cuh,cu}, PI/car_bfs.cuh,
//     PI/car_kinematics.cuh, PI/costs.{cuh,cu}), built for sm_100a instead of sm_52 and without -maxrregcount=32;
//     cuRAND (XORWOW, seed 1234) is the toolkit's own library.
//   * also the reference's, unmodified: DDP<...>::run with TrackingCostDDP / ModelWrapperDDP (DDP/*.h) behind
//     computeFeedbackGains, and runControlLoop (PI/run_control_loop.cuh) with its two-controller arbitration;
//   * stand-ins (oracle/ref_shim/): Eigen, cnpy, ROS/XmlRpc (+ message headers), OpenCV and Boost, none of which is
//     installed here.  They only carry host-side containers and plain float loops; no device code uses them.  The plant
//     runControlLoop drives is declared by the reference's PI/autorally_plant.h; its ROS implementation
//     (SRC/autorally_plant.cpp) is replaced below by a recording stand-in (debug mode: the loop itself integrates the model).
// The harness reaches private members the way the reference's own tests do (`#define private public`,
// autorally_core/test/serialSensorInterfaceTest.cpp:43-61).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <map>
#include <string>
#include <vector>
#include <unistd.h>

#include <cuda_runtime.h>
#include <curand.h>
#include <vector_types.h>

#include <Eigen/Dense>
#include <cnpy.h>
#include <opencv2/core/core.hpp>
#include <ros/ros.h>
#include <autorally_control/PathIntegralParamsConfig.h>
#include <autorally_control/ddp/ddp.h>

#define private public
#define protected public
#include <autorally_control/path_integral/costs.cuh>
#include <autorally_control/path_integral/neural_net_model.cuh>
#include <autorally_control/path_integral/car_bfs.cuh>
#include <autorally_control/path_integral/car_kinematics.cuh>
#include <autorally_control/path_integral/generalized_linear.cuh>
#include <autorally_control/path_integral/mppi_controller.cuh>
#include <autorally_control/path_integral/run_control_loop.cuh>   // + PI/autorally_plant.h (class declaration only)
#undef private
#undef protected

using namespace autorally_control;

extern "C" {
// MPPICosts::CostParams field order (PI/costs.cuh:67-85) + l1_cost_
typedef struct ref_cost_params {
  float desired_speed, speed_coeff, track_coeff, max_slip_ang, slip_penalty, track_slop, crash_coeff;
  float steering_coeff, throttle_coeff, boundary_threshold, discount;
  int num_timesteps, grid_res;
  float r_c1[3], r_c2[3], trs[3];
  int l1_cost;
} ref_cost_params;
}

// ---- recording stand-in for the member functions of AutorallyPlant that runControlLoop calls (declared by the reference's
//      PI/autorally_plant.h:94-300; the ROS node of SRC/autorally_plant.cpp is not built) ----
namespace {
struct LoopRecorder {
  int T = 0, N = 0;
  size_t noise_per_iter = 0;
  std::vector<float> states, controls, tcost, gains, eps, u_actual, u_predicted;   // per iteration
  std::vector<int> used;
  // filled by RefImpl::run_loop: how to read the two controllers after each iteration
  float (*traj_cost)(void *) = nullptr;
  const std::vector<float> *(*controls_of)(void *) = nullptr;
  void *actual = nullptr, *predicted = nullptr;
  curandGenerator_t twin = nullptr;
  float *eps_d = nullptr;
};
LoopRecorder *g_rec = nullptr;
}  // namespace

namespace autorally_control {
AutorallyPlant::AutorallyPlant(ros::NodeHandle, ros::NodeHandle, std::map<std::string, XmlRpc::XmlRpcValue> *params, bool nodelet)
    : new_model_available_(false), is_nodelet_(nodelet), status_(1), debug_mode_(true), activated_(true) {
  std::memset(&full_state_, 0, sizeof(full_state_));
  full_state_.x_pos = (float)(double)(*params)["x_pos"];
  full_state_.y_pos = (float)(double)(*params)["y_pos"];
  full_state_.yaw = (float)(double)(*params)["heading"];
  full_state_.q0 = 1.0f;
  hz_ = (int)(*params)["hz"];
  numTimesteps_ = (int)(*params)["num_timesteps"];
}
AutorallyPlant::FullState AutorallyPlant::getState() { return full_state_; }
ros::Time AutorallyPlant::getLastPoseTime() { return last_pose_call_; }
void AutorallyPlant::setTimingInfo(double, double, double) {}
bool AutorallyPlant::hasNewDynRcfg() { return false; }
autorally_control::PathIntegralParamsConfig AutorallyPlant::getDynRcfgParams() { return costParams_; }
bool AutorallyPlant::hasNewModel() { return false; }
void AutorallyPlant::getModel(std::vector<int> &, std::vector<float> &) {}
void AutorallyPlant::modelCall(autorally_msgs::neuralNetModel) {}
void AutorallyPlant::setDebugImage(cv::Mat) {}
int AutorallyPlant::checkStatus() { return 1; }  // "no pose updates": with debug_mode the loop integrates the model (PI/run_control_loop.cuh:296)
void AutorallyPlant::displayDebugImage(const ros::TimerEvent &) {}
void AutorallyPlant::shutdown() {}
void AutorallyPlant::setSolution(std::vector<float> traj, std::vector<float> controls, util::EigenAlignedVector<float, 2, 7> gains,
                                 ros::Time, double, ControllerType used) {
  LoopRecorder *r = g_rec;
  if (!r) return;
  r->states.insert(r->states.end(), traj.begin(), traj.begin() + 7);
  r->controls.insert(r->controls.end(), controls.begin(), controls.begin() + 2);
  r->used.push_back(used == ControllerType::ACTUAL_STATE ? 0 : 1);
  r->tcost.push_back(r->traj_cost(r->actual));
  r->tcost.push_back(r->traj_cost(r->predicted));
  const std::vector<float> *ua = r->controls_of(r->actual), *up = r->controls_of(r->predicted);
  r->u_actual.insert(r->u_actual.end(), ua->begin(), ua->end());
  r->u_predicted.insert(r->u_predicted.end(), up->begin(), up->end());
  for (int k = 0; k < r->T; k++)
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 7; b++) r->gains.push_back(k < (int)gains.size() ? gains[k](a, b) : 0.0f);
  // the draws both controllers consumed in this iteration: each owns a cuRAND generator seeded 1234 (PI/mppi_controller.cu:330-331)
  // and makes one curandGenerateNormal call per computeControl, so the two draw the SAME sequence; the twin replays it
  curandGenerateNormal(r->twin, r->eps_d, r->noise_per_iter, 0.0f, 1.0f);
  const size_t off = r->eps.size();
  r->eps.resize(off + r->noise_per_iter);
  cudaMemcpy(r->eps.data() + off, r->eps_d, r->noise_per_iter * sizeof(float), cudaMemcpyDeviceToHost);
}
}  // namespace autorally_control

namespace {

struct RefBase {
  virtual ~RefBase() {}
  virtual int num_rollouts() const = 0;
  virtual int set_controls(const float *U, const float *hist) = 0;
  virtual int get_controls(float *U, float *hist) = 0;
  virtual int slide(int stride) = 0;
  virtual int compute(const float *state, float *eps_out, float *U_out, float *ss, float *cs, float *scalars, float *weights) = 0;
  virtual int rollout_costs(const float *state, const float *U, const float *eps, float *costs, float *V) = 0;
  virtual int time_compute(const float *state, int reps, float *ms_per_call) = 0;
  virtual int time_kernels(const float *state, int reps, float *ms4) = 0;
  virtual int feedback_gains(const float *state, float *gains, float *feedforward) = 0;
  virtual int run_loop(const float *pose3, int iterations, int use_feedback_gains, LoopRecorder *rec) = 0;
};

typedef NeuralNetModel<7, 2, 3, 6, 32, 32, 4> RefNN;
typedef NeuralNetModel<7, 2, 3, 6, 64, 64, 64, 64, 4> RefNN64;  // SRC/params/models/wider_deeper_network_08_20_2020.npz
typedef NeuralNetModel<7, 2, 3, 6, 16, 16, 4> RefNN16;  // layer packs without a dedicated kernel in the product: the
typedef NeuralNetModel<7, 2, 3, 6, 48, 4> RefNN48;      // reference's variadic template takes any (PI/neural_net_model.cuh:48-52)
typedef GeneralizedLinear<CarBasisFuncs, 7, 2, 25, CarKinematics, 3> RefBF;

template <class MODEL, int ROLLOUTS, int BX, int BY>
struct RefImpl : RefBase {
  typedef MPPIController<MODEL, MPPICosts, ROLLOUTS, BX, BY> Controller;
  static const int N = Controller::NUM_ROLLOUTS;
  MODEL *model = nullptr;
  MPPICosts *costs = nullptr;
  Controller *ctrl = nullptr;
  float2 ranges[2];
  curandGenerator_t twin = nullptr;
  float *eps_d = nullptr;
  int T = 0, iters = 1, opt_stride = 1;

  ~RefImpl() override {
    // MPPIController::deallocateCudaMem destroys stream 0 and frees shared memory: free piecewise.
    if (ctrl) {
      cudaFree(ctrl->state_d_); cudaFree(ctrl->nu_d_); cudaFree(ctrl->traj_costs_d_); cudaFree(ctrl->U_d_); cudaFree(ctrl->du_d_);
      curandDestroyGenerator(ctrl->gen_);
      delete ctrl;
    }
    if (model) { model->freeCudaMem(); delete model; }
    if (costs) { costs->freeCudaMem(); delete costs; }
    if (twin) curandDestroyGenerator(twin);
    cudaFree(eps_d);
    cudaGetLastError();
  }

  int init_common(const float *lo_hi, const float *costmap, int w, int h, const ref_cost_params *cp, const float *nu,
                  const float *init_u, int hz, int T_, int opt_stride_, float gamma, int num_iters) {
    T = T_; iters = num_iters; opt_stride = opt_stride_;
    nu_saved[0] = nu[0]; nu_saved[1] = nu[1]; init_u_saved[0] = init_u[0]; init_u_saved[1] = init_u[1]; gamma_saved = gamma; hz_saved = hz;
    costs = new MPPICosts(w, h);
    costs->costmap_tex_ = 0;  // uninitialised in the reference; destroyed by the first costmapToTexture (PI/costs.cu:152)
    costs->l1_cost_ = cp->l1_cost != 0;  // uninitialised by the (w, h) constructor (PI/costs.cu:41-50)
    MPPICosts::CostParams &p = costs->params_;
    p.desired_speed = cp->desired_speed; p.speed_coeff = cp->speed_coeff; p.track_coeff = cp->track_coeff;
    p.max_slip_ang = cp->max_slip_ang; p.slip_penalty = cp->slip_penalty; p.track_slop = cp->track_slop;
    p.crash_coeff = cp->crash_coeff; p.steering_coeff = cp->steering_coeff; p.throttle_coeff = cp->throttle_coeff;
    p.boundary_threshold = cp->boundary_threshold; p.discount = cp->discount; p.num_timesteps = cp->num_timesteps;
    p.grid_res = cp->grid_res;
    Eigen::MatrixXf m = Eigen::MatrixXf::Zero(3, 3);
    Eigen::ArrayXf trs(3);
    for (int i = 0; i < 3; i++) { m(i, 0) = cp->r_c1[i]; m(i, 1) = cp->r_c2[i]; trs(i) = cp->trs[i]; }
    costs->updateTransform(m, trs);  // PI/costs.cu:176-188 (+ paramsToDevice)
    std::vector<float> map(costmap, costmap + (size_t)w * h);
    costs->costmapToTexture(map.data(), 0);  // channel 0 = .x, the only channel computeCost reads (PI/costs.cu:373-380)
    float nu2[2] = {nu[0], nu[1]}, iu[2] = {init_u[0], init_u[1]};
    ctrl = new Controller(model, costs, nu2, iu, hz, T, opt_stride, gamma, num_iters, 0);
    if (curandCreateGenerator(&twin, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) return -2;
    curandSetPseudoRandomGeneratorSeed(twin, 1234ULL);  // PI/mppi_controller.cu:330-331
    curandSetStream(twin, 0);
    if (cudaMalloc((void **)&eps_d, (size_t)iters * N * T * 2 * sizeof(float)) != cudaSuccess) return -3;
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : (int)e;
  }

  int num_rollouts() const override { return N; }

  int set_controls(const float *U, const float *hist) override {
    for (int i = 0; i < 2 * T; i++) ctrl->U_[i] = U[i];
    if (hist) for (int i = 0; i < 4; i++) ctrl->control_hist_[i] = hist[i];
    return 0;
  }
  int get_controls(float *U, float *hist) override {
    for (int i = 0; i < 2 * T; i++) U[i] = ctrl->U_[i];
    if (hist) for (int i = 0; i < 4; i++) hist[i] = ctrl->control_hist_[i];
    return 0;
  }
  int slide(int stride) override { ctrl->slideControlAndStateSeq(stride); return 0; }

  int compute(const float *state, float *eps_out, float *U_out, float *ss, float *cs, float *scalars, float *weights) override {
    const size_t count = (size_t)N * T * 2;
    // the twin generator replays exactly the draws computeControl is about to make (same seed, same call sequence)
    for (int it = 0; it < iters; it++)
      if (curandGenerateNormal(twin, eps_d + it * count, count, 0.0f, 1.0f) != CURAND_STATUS_SUCCESS) return -2;
    Eigen::Matrix<float, 7, 1> s;
    for (int i = 0; i < 7; i++) s(i) = state[i];
    ctrl->computeControl(s);  // PI/mppi_controller.cu:600-675
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    if (eps_out) cudaMemcpy(eps_out, eps_d, iters * count * sizeof(float), cudaMemcpyDeviceToHost);
    if (U_out) for (int i = 0; i < 2 * T; i++) U_out[i] = ctrl->U_[i];
    if (ss) for (int i = 0; i < 7 * T; i++) ss[i] = ctrl->state_solution_[i];
    if (cs) for (int i = 0; i < 2 * T; i++) cs[i] = ctrl->control_solution_[i];
    if (scalars) { scalars[0] = ctrl->normalizer_; scalars[1] = ctrl->trajectory_cost_; }
    if (weights) for (int i = 0; i < N; i++) weights[i] = ctrl->traj_costs_[i];  // exp(-gamma (c - min c)) of the last iteration
    return 0;
  }

  // The reference's own launchRolloutKernel on caller-supplied noise: raw rollout costs and the sampled controls.
  int rollout_costs(const float *state, const float *U, const float *eps, float *costs_out, float *V) override {
    const size_t count = (size_t)N * T * 2;
    costs->paramsToDevice();
    model->paramsToDevice();
    cudaMemcpy(ctrl->state_d_, state, 7 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(ctrl->U_d_, U, 2 * T * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(ctrl->du_d_, eps, count * sizeof(float), cudaMemcpyHostToDevice);
    launchRolloutKernel<MODEL, MPPICosts, ROLLOUTS, BX, BY>(T, ctrl->state_d_, ctrl->U_d_, ctrl->du_d_, ctrl->nu_d_,
                                                           ctrl->traj_costs_d_, model, costs, opt_stride, (cudaStream_t)0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    if (costs_out) cudaMemcpy(costs_out, ctrl->traj_costs_d_, N * sizeof(float), cudaMemcpyDeviceToHost);
    if (V) cudaMemcpy(V, ctrl->du_d_, count * sizeof(float), cudaMemcpyDeviceToHost);
    return 0;
  }

  // Wall-clock of the reference's computeControl on this GPU (its own host syncs and memcpys included).
  int time_kernels(const float *state, int reps, float *ms4) override;

  // computeFeedbackGains (PI/mppi_controller.cu:425-437) -> DDP<ModelWrapperDDP<MODEL>>::run (DDP/ddp.h:49-157) around the
  // controller's current state / control solution; gains [T][2][7], feedforward [T][2]
  int feedback_gains(const float *state, float *gains, float *feedforward) override {
    Eigen::MatrixXf s(7, 1);
    for (int i = 0; i < 7; i++) s(i) = state[i];
    ctrl->computeFeedbackGains(s);
    auto res = ctrl->getFeedbackGains();
    for (int k = 0; k < T; k++)
      for (int a = 0; a < 2; a++) {
        for (int b = 0; b < 7; b++) gains[(k * 2 + a) * 7 + b] = k < (int)res.feedback_gain.size() ? res.feedback_gain[k](a, b) : 0.0f;
        if (feedforward) feedforward[k * 2 + a] = res.feedforward_gain(a, k);
      }
    return 0;
  }

  // The reference's runControlLoop (PI/run_control_loop.cuh:84-321), unmodified, in debug mode on two fresh controllers that
  // share this model and cost object (SRC/path_integral_main.cu:119-122), for `iterations` iterations.
  float nu_saved[2] = {0, 0}, init_u_saved[2] = {0, 0}, gamma_saved = 0;
  int hz_saved = 50;
  static float traj_cost_of(void *c) { return static_cast<Controller *>(c)->getComputedTrajectoryCost(); }
  static const std::vector<float> *controls_of(void *c) { return &static_cast<Controller *>(c)->U_; }
  int run_loop(const float *pose3, int iterations, int use_feedback_gains, LoopRecorder *rec) override {
    std::map<std::string, XmlRpc::XmlRpcValue> params;
    params["x_pos"] = XmlRpc::XmlRpcValue((double)pose3[0]);
    params["y_pos"] = XmlRpc::XmlRpcValue((double)pose3[1]);
    params["heading"] = XmlRpc::XmlRpcValue((double)pose3[2]);
    params["hz"] = XmlRpc::XmlRpcValue(hz_saved);
    params["optimization_stride"] = XmlRpc::XmlRpcValue(opt_stride);
    params["num_timesteps"] = XmlRpc::XmlRpcValue(T);
    params["use_feedback_gains"] = XmlRpc::XmlRpcValue(use_feedback_gains != 0);
    params["debug_mode"] = XmlRpc::XmlRpcValue(true);
    params["use_only_actual_state_controller"] = XmlRpc::XmlRpcValue(false);
    params["use_only_predicted_state_controller"] = XmlRpc::XmlRpcValue(false);
    params["profiler_max_iter"] = XmlRpc::XmlRpcValue(iterations);
    Controller *actual = new Controller(model, costs, nu_saved, init_u_saved, hz_saved, T, opt_stride, gamma_saved, iters, 0);
    Controller *predicted = new Controller(model, costs, nu_saved, init_u_saved, hz_saved, T, opt_stride, gamma_saved, iters, 0);
    ros::NodeHandle nh;
    AutorallyPlant robot(nh, &params);
    rec->T = T; rec->N = N; rec->noise_per_iter = (size_t)N * T * 2;
    rec->traj_cost = &traj_cost_of; rec->controls_of = &controls_of; rec->actual = actual; rec->predicted = predicted;
    curandCreateGenerator(&rec->twin, CURAND_RNG_PSEUDO_DEFAULT);
    curandSetPseudoRandomGeneratorSeed(rec->twin, 1234ULL);
    curandSetStream(rec->twin, 0);
    cudaMalloc((void **)&rec->eps_d, rec->noise_per_iter * sizeof(float));
    g_rec = rec;
    std::atomic<bool> is_alive(true);
    runControlLoop<Controller>(predicted, actual, &robot, &params, &is_alive);
    g_rec = nullptr;
    cudaError_t e = cudaDeviceSynchronize();
    curandDestroyGenerator(rec->twin);
    cudaFree(rec->eps_d);
    for (Controller *c : {actual, predicted}) {
      cudaFree(c->state_d_); cudaFree(c->nu_d_); cudaFree(c->traj_costs_d_); cudaFree(c->U_d_); cudaFree(c->du_d_);
      curandDestroyGenerator(c->gen_);
      delete c;
    }
    return e == cudaSuccess ? 0 : (int)e;
  }

  int time_compute(const float *state, int reps, float *ms_per_call) override {
    Eigen::Matrix<float, 7, 1> s;
    for (int i = 0; i < 7; i++) s(i) = state[i];
    std::vector<float> U0(ctrl->U_), h0(ctrl->control_hist_);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    ctrl->computeControl(s);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int r = 0; r < reps; r++) { ctrl->U_ = U0; ctrl->control_hist_ = h0; ctrl->computeControl(s); }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctrl->U_ = U0; ctrl->control_hist_ = h0;
    // keep the twin generator in step with the reference's generator
    const size_t count = (size_t)N * T * 2;
    for (int r = 0; r < (reps + 1) * iters; r++) curandGenerateNormal(twin, eps_d, count, 0.0f, 1.0f);
    cudaDeviceSynchronize();
    *ms_per_call = ms / reps;
    return 0;
  }
};

template <class MODEL, int ROLLOUTS, int BX, int BY>
int RefImpl<MODEL, ROLLOUTS, BX, BY>::time_kernels(const float *state, int reps, float *ms4) {
  // Device time of the reference's four GPU stages, each launched exactly as computeControl launches it
  // (PI/mppi_controller.cu:612-660) but through cudaLaunchKernel with pointers to the live model / cost objects, so that
  // the by-value kernel arguments are byte copies instead of C++ copies (launchRolloutKernel's `*mppi_costs` argument
  // deep-copies the std::vector<float4> costmap on the host at every launch, PI/mppi_controller.cu:285-286): what is
  // timed here is the kernels alone.  ms4 = {curandGenerateNormal, rolloutKernel, normExpKernel, weightedReductionKernel}.
  const size_t count = (size_t)N * T * 2;
  costs->paramsToDevice();
  model->paramsToDevice();
  cudaMemcpy(ctrl->state_d_, state, 7 * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(ctrl->U_d_, ctrl->U_.data(), 2 * T * sizeof(float), cudaMemcpyHostToDevice);
  std::vector<cudaEvent_t> ev(8 * (size_t)reps);
  for (auto &e : ev) cudaEventCreate(&e);
  int num_timesteps = T, opt_delay = opt_stride;
  float gamma = ctrl->gamma_, baseline = 0.0f, normalizer = 1.0f;
  void *roll_args[] = {&num_timesteps, &ctrl->state_d_, &ctrl->U_d_, &ctrl->du_d_, &ctrl->nu_d_, &ctrl->traj_costs_d_, model, costs, &opt_delay};
  void *norm_args[] = {&ctrl->traj_costs_d_, &gamma, &baseline};
  void *wred_args[] = {&ctrl->traj_costs_d_, &ctrl->du_d_, &ctrl->nu_d_, &normalizer, &num_timesteps};
  const dim3 rb(BX, BY, 1), rg((N - 1) / BX + 1, 1, 1), nb(BX, 1, 1), ng((N - 1) / BX + 1, 1, 1), wb((N - 1) / 64 + 1, 1, 1), wg(T, 1, 1);
  std::vector<float> host_costs(N);
  for (int r = -1; r < reps; r++) {  // r = -1: warm-up, and the baseline / normaliser of a real call
    cudaEvent_t *e = r < 0 ? nullptr : &ev[8 * (size_t)r];
    if (e) cudaEventRecord(e[0], 0);
    curandGenerateNormal(ctrl->gen_, ctrl->du_d_, count, 0.0f, 1.0f);  // the controller's own generator (PI/mppi_controller.cu:612)
    if (e) { cudaEventRecord(e[1], 0); cudaEventRecord(e[2], 0); }
    cudaLaunchKernel((const void *)rolloutKernel<MODEL, MPPICosts, ROLLOUTS, BX, BY>, rg, rb, roll_args, 0, 0);
    if (e) cudaEventRecord(e[3], 0);
    if (r < 0) {
      cudaMemcpy(host_costs.data(), ctrl->traj_costs_d_, N * sizeof(float), cudaMemcpyDeviceToHost);
      baseline = host_costs[0];
      for (int i = 0; i < N; i++) baseline = host_costs[i] < baseline ? host_costs[i] : baseline;
      normalizer = 0.0f;
      for (int i = 0; i < N; i++) normalizer += expf(-gamma * (host_costs[i] - baseline));
    }
    if (e) cudaEventRecord(e[4], 0);
    cudaLaunchKernel((const void *)normExpKernel<MODEL, MPPICosts, ROLLOUTS, BX, BY>, ng, nb, norm_args, 0, 0);
    if (e) { cudaEventRecord(e[5], 0); cudaEventRecord(e[6], 0); }
    cudaLaunchKernel((const void *)weightedReductionKernel<MODEL, MPPICosts, ROLLOUTS, BX, BY>, wg, wb, wred_args, 0, 0);
    if (e) cudaEventRecord(e[7], 0);
  }
  for (int r = -1; r < reps; r++) curandGenerateNormal(twin, eps_d, count, 0.0f, 1.0f);  // keep the twin generator in step
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) return (int)err;
  for (int k = 0; k < 4; k++) ms4[k] = 0.0f;
  for (int r = 0; r < reps; r++)
    for (int k = 0; k < 4; k++) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, ev[8 * (size_t)r + 2 * k], ev[8 * (size_t)r + 2 * k + 1]);
      ms4[k] += ms / reps;
    }
  for (auto &e : ev) cudaEventDestroy(e);
  return 0;
}

template <class NN, int NLAYERS, int ROLLOUTS, int BX, int BY>
RefBase *make_nn_t(const int *widths, const float *theta, int negate_yaw, const float *lo_hi, const float *costmap, int w, int h,
                 const ref_cost_params *cp, const float *nu, const float *init_u, int hz, int T, int opt_stride, float gamma,
                 int num_iters, int *rc) {
  auto *r = new RefImpl<NN, ROLLOUTS, BX, BY>();
  r->ranges[0] = make_float2(lo_hi[0], lo_hi[1]);
  r->ranges[1] = make_float2(lo_hi[2], lo_hi[3]);
  r->model = new NN(1.0 / hz, r->ranges);  // SRC/path_integral_main.cu:100
  r->model->negate_yaw_der = negate_yaw != 0;
  // theta is packed [W1|b1|W2|b2|W3|b3] row-major (PI/neural_net_model.cu:125-141); hand it over through setParams
  typedef Eigen::Matrix<float, -1, -1, Eigen::RowMajor> RowMat;
  RowMat W[NLAYERS - 1], B[NLAYERS - 1];
  size_t off = 0;
  for (int l = 0; l < NLAYERS - 1; l++) {
    const int nin = widths[l], nout = widths[l + 1];
    W[l] = RowMat::Zero(nout, nin);
    B[l] = RowMat::Zero(nout, 1);
    for (int j = 0; j < nout; j++) for (int k = 0; k < nin; k++) W[l](j, k) = theta[off + (size_t)j * nin + k];
    off += (size_t)nin * nout;
    for (int j = 0; j < nout; j++) B[l](j, 0) = theta[off + j];
    off += nout;
  }
  r->model->setParams(W, B);
  *rc = r->init_common(lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters);
  return r;
}

template <int ROLLOUTS, int BX, int BY>
RefBase *make_nn(const float *theta, int negate_yaw, const float *lo_hi, const float *costmap, int w, int h,
                 const ref_cost_params *cp, const float *nu, const float *init_u, int hz, int T, int opt_stride, float gamma,
                 int num_iters, int *rc) {
  static const int widths[4] = {6, 32, 32, 4};
  return make_nn_t<RefNN, 4, ROLLOUTS, BX, BY>(widths, theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, rc);
}

template <int ROLLOUTS, int BX, int BY>
RefBase *make_nn64(const float *theta, int negate_yaw, const float *lo_hi, const float *costmap, int w, int h,
                   const ref_cost_params *cp, const float *nu, const float *init_u, int hz, int T, int opt_stride, float gamma,
                   int num_iters, int *rc) {
  static const int widths[6] = {6, 64, 64, 64, 64, 4};
  return make_nn_t<RefNN64, 6, ROLLOUTS, BX, BY>(widths, theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, rc);
}

template <class NN, int ROLLOUTS, int BX, int BY, int... WIDTHS>
RefBase *make_nn_pack(const float *theta, int negate_yaw, const float *lo_hi, const float *costmap, int w, int h,
                      const ref_cost_params *cp, const float *nu, const float *init_u, int hz, int T, int opt_stride, float gamma,
                      int num_iters, int *rc) {
  static const int widths[sizeof...(WIDTHS)] = {WIDTHS...};
  return make_nn_t<NN, (int)sizeof...(WIDTHS), ROLLOUTS, BX, BY>(widths, theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, rc);
}

template <int ROLLOUTS, int BX, int BY>
RefBase *make_bf(const float *theta, const float *lo_hi, const float *costmap, int w, int h, const ref_cost_params *cp,
                 const float *nu, const float *init_u, int hz, int T, int opt_stride, float gamma, int num_iters, int *rc) {
  auto *r = new RefImpl<RefBF, ROLLOUTS, BX, BY>();
  r->ranges[0] = make_float2(lo_hi[0], lo_hi[1]);
  r->ranges[1] = make_float2(lo_hi[2], lo_hi[3]);
  r->model = new RefBF(1.0 / hz, r->ranges);
  Eigen::Matrix<float, 4, 25, Eigen::RowMajor> th;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 25; j++) th(i, j) = theta[i * 25 + j];
  r->model->setParams(th);
  *rc = r->init_common(lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters);
  return r;
}

}  // namespace

extern "C" {

enum { REF_NN_1920 = 0, REF_BF_2560 = 1, REF_NN_256 = 2, REF_NN_4096 = 3, REF_BF_256 = 4, REF_NN64_1920 = 5, REF_NN16_1920 = 6, REF_NN48_1920 = 7 };

const char *ref_version(void) { return "rdesc/autorally MPPIController (reference sources, sm_100a build, shimmed host libraries)"; }

// kind: REF_NN_1920 = MPPIController<NeuralNetModel<7,2,3,6,32,32,4>, MPPICosts, 1920, 8, 16>  (SRC/path_integral_main.cu:66-69)
//       REF_BF_2560 = MPPIController<GeneralizedLinear<CarBasisFuncs,7,2,25,CarKinematics,3>, MPPICosts, 2560, 16, 4>  (:71-74)
//       REF_NN_256 / REF_NN_4096 / REF_BF_256 = the same controllers with 256 / 4096 rollouts (small fixtures, ragged sizes)
//       REF_NN64_1920 = MPPIController<NeuralNetModel<7,2,3,6,64,64,64,64,4>, MPPICosts, 1920, 8, 16> (the fork's wider_deeper network)
int ref_create(int kind, const float *theta, int negate_yaw, const float *lo_hi, const float *costmap, int w, int h,
               const ref_cost_params *cp, const float *nu, const float *init_u, int hz, int T, int opt_stride, float gamma,
               int num_iters, void **out) {
  if (!theta || !lo_hi || !costmap || !cp || !nu || !init_u || !out) return -1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return -4; }
  int rc = 0;
  RefBase *r = nullptr;
  switch (kind) {
    case REF_NN_1920: r = make_nn<1920, 8, 16>(theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    case REF_BF_2560: r = make_bf<2560, 16, 4>(theta, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    case REF_NN_256: r = make_nn<256, 8, 16>(theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    case REF_NN_4096: r = make_nn<4096, 8, 16>(theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    case REF_BF_256: r = make_bf<256, 16, 4>(theta, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    case REF_NN64_1920: r = make_nn64<1920, 8, 16>(theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    // NeuralNetModel<7,2,3,6,16,16,4> and <7,2,3,6,48,4>: arbitrary layer packs (weights supplied by the caller)
    case REF_NN16_1920: r = make_nn_pack<RefNN16, 1920, 8, 16, 6, 16, 16, 4>(theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    case REF_NN48_1920: r = make_nn_pack<RefNN48, 1920, 8, 16, 6, 48, 4>(theta, negate_yaw, lo_hi, costmap, w, h, cp, nu, init_u, hz, T, opt_stride, gamma, num_iters, &rc); break;
    default: return -1;
  }
  if (rc) { delete r; return rc; }
  *out = r;
  return 0;
}
int ref_destroy(void *h) { delete static_cast<RefBase *>(h); return 0; }
int ref_num_rollouts(void *h) { return static_cast<RefBase *>(h)->num_rollouts(); }
int ref_set_controls(void *h, const float *U, const float *hist) { return static_cast<RefBase *>(h)->set_controls(U, hist); }
int ref_get_controls(void *h, float *U, float *hist) { return static_cast<RefBase *>(h)->get_controls(U, hist); }
int ref_slide(void *h, int stride) { return static_cast<RefBase *>(h)->slide(stride); }
// scalars = {normalizer_, trajectory_cost_}; weights[N] = exp(-gamma (c_i - min c)) of the last iteration;
// eps_out[num_iters][N][T][2] = the N(0,1) draws the call consumed.
int ref_compute_control(void *h, const float *state, float *eps_out, float *U_out, float *state_solution, float *control_solution,
                        float *scalars, float *weights) {
  return static_cast<RefBase *>(h)->compute(state, eps_out, U_out, state_solution, control_solution, scalars, weights);
}
int ref_rollout_costs(void *h, const float *state, const float *U, const float *eps, float *costs, float *V) {
  return static_cast<RefBase *>(h)->rollout_costs(state, U, eps, costs, V);
}
int ref_time_compute_control(void *h, const float *state, int reps, float *ms_per_call) {
  return static_cast<RefBase *>(h)->time_compute(state, reps, ms_per_call);
}
// ms4 = device time of {curandGenerateNormal, rolloutKernel, normExpKernel, weightedReductionKernel} per computeControl iteration
int ref_time_kernels(void *h, const float *state, int reps, float *ms4) { return static_cast<RefBase *>(h)->time_kernels(state, reps, ms4); }
// computeFeedbackGains around the controller's current solution (after ref_compute_control): gains [T][2][7], feedforward [T][2] or NULL
int ref_feedback_gains(void *h, const float *state, float *gains, float *feedforward) {
  return static_cast<RefBase *>(h)->feedback_gains(state, gains, feedforward);
}
// The reference's runControlLoop in debug mode (two controllers sharing model and costs) for `iterations` iterations from pose
// (x, y, heading).  Per iteration: states [7] / controls [2] handed to the plant, controller_used (0 actual, 1 predicted), the two
// controllers' trajectory costs [2], their control sequences [T][2], the gains handed over [T][2][7], and the N(0,1) draws both
// controllers consumed [N][T][2].  Any output may be NULL.
int ref_run_control_loop(void *h, const float *pose3, int iterations, int use_feedback_gains, float *states, float *controls, int *used,
                         float *tcost, float *u_actual, float *u_predicted, float *gains, float *eps) {
  LoopRecorder rec;
  int rc = static_cast<RefBase *>(h)->run_loop(pose3, iterations, use_feedback_gains, &rec);
  if (rc) return rc;
  if ((int)rec.used.size() != iterations) return -7;
  auto put = [](float *dst, const std::vector<float> &src) { if (dst) std::memcpy(dst, src.data(), src.size() * sizeof(float)); };
  put(states, rec.states); put(controls, rec.controls); put(tcost, rec.tcost); put(u_actual, rec.u_actual);
  put(u_predicted, rec.u_predicted); put(gains, rec.gains); put(eps, rec.eps);
  if (used) std::memcpy(used, rec.used.data(), rec.used.size() * sizeof(int));
  return 0;
}

}  // extern "C"
