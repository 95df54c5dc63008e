// Stand-ins for the reference's DDP headers (DDP/*.h need Eigen's decompositions and unsupported/NumericalDiff,
// which the Eigen stand-in does not provide).  They shadow <autorally_control/ddp/*.h> through include order so that
// MPPIController compiles; computeFeedbackGains is NOT exercised by the reference harness (it is off the GPU path,
// SURVEY.md section 8 f-2).  TEST INFRASTRUCTURE (oracle/refbuild.py).
#ifndef REF_SHIM_DDP_H_
#define REF_SHIM_DDP_H_
#include <Eigen/Dense>
namespace util {
struct DefaultLogger {};
}
// (the reference's DDP classes live in the global namespace)
template <class DYNAMICS_T>
struct ModelWrapperDDP {
  static const int StateSize = DYNAMICS_T::STATE_DIM;
  static const int ControlSize = DYNAMICS_T::CONTROL_DIM;
  DYNAMICS_T *model_;
  explicit ModelWrapperDDP(DYNAMICS_T *model) : model_(model) {}
};
template <class M>
struct OptimizerResult {
  Eigen::MatrixXf state_trajectory, control_trajectory;
};
template <class M>
struct TrackingCostDDP {
  typedef Eigen::Matrix<float, M::StateSize, M::StateSize> StateCostWeight;
  typedef Eigen::Matrix<float, M::ControlSize, M::ControlSize> ControlCostWeight;
  Eigen::MatrixXf traj_target_x_, traj_target_u_;
  TrackingCostDDP(const StateCostWeight &, const ControlCostWeight &, int T)
      : traj_target_x_(M::StateSize, T), traj_target_u_(M::ControlSize, T) {}
  void setTargets(float *x, float *u, int T) {
    for (int t = 0; t < T; t++) {
      for (int i = 0; i < M::StateSize; i++) traj_target_x_(i, t) = x[M::StateSize * t + i];
      for (int i = 0; i < M::ControlSize; i++) traj_target_u_(i, t) = u[M::ControlSize * t + i];
    }
  }
};
template <class M>
struct TrackingTerminalCost {
  typedef Eigen::Matrix<float, M::StateSize, M::StateSize> Hessian;
  Eigen::Matrix<float, M::StateSize, 1> xf;
  explicit TrackingTerminalCost(const Hessian &) {}
};
template <class M>
struct DDP {
  DDP(double, int, int, util::DefaultLogger *, bool) {}
  template <class S, class U, class RC, class TC, class B>
  OptimizerResult<M> run(const S &, const U &, M &, RC &, TC &, const B &, const B &) { return OptimizerResult<M>(); }
};
#endif
