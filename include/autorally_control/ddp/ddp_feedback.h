// ddp_feedback.h -- feedback gains around the MPPI solution: one DDP/iLQR iteration on a quadratic
// tracking cost, as MPPIController::computeFeedbackGains runs it in the reference
// (PI/mppi_controller.cu:402-445 -> DDP<...>::run with iterations = 1, DDP/ddp.h:49-157, tracking cost
// DDP/ddp_tracking_costs.h:7-117, model wrapper DDP/ddp_model_wrapper.h:9-80).  CPU only, off the GPU
// hot path; plain loops on small dense matrices instead of Eigen expression templates.
//
//   x_0 = x0, x_{k+1} = x_k + f(x_k, clamp(u_k)) dt                       (forward rollout)
//   A_k = I + df/dx dt, B_k = df/du dt                                      (analytic Jacobian if the model has
//                                                                            computeGrad, else central differences)
//   backward pass with l = (x-x*)'Q(x-x*) + (u-u*)'R(u-u*): gains K_k = -Quu^-1 Qux, k_k = -Quu^-1 Qu
//   forward pass with alpha = 1 (always accepted on the first iteration)
#ifndef MPPI_DDP_FEEDBACK_H_
#define MPPI_DDP_FEEDBACK_H_
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <type_traits>
#include <utility>
#include <vector>

#include <Eigen/Dense>

template <class DYNAMICS_T>
struct ModelWrapperDDP {
  typedef float Scalar;
  enum { StateSize = DYNAMICS_T::STATE_DIM, ControlSize = DYNAMICS_T::CONTROL_DIM };
  typedef Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, 1> State;
  typedef Eigen::Matrix<float, DYNAMICS_T::CONTROL_DIM, 1> Control;
  typedef Eigen::Matrix<float, DYNAMICS_T::STATE_DIM, DYNAMICS_T::STATE_DIM + DYNAMICS_T::CONTROL_DIM> Jacobian;
  typedef Eigen::Matrix<float, DYNAMICS_T::CONTROL_DIM, DYNAMICS_T::STATE_DIM> FeedbackGain;

  DYNAMICS_T *model_;
  Eigen::MatrixXf state, control;

  explicit ModelWrapperDDP(DYNAMICS_T *model) : model_(model), state(StateSize, 1), control(ControlSize, 1) {}

  /// continuous-time dynamics without constraint enforcement (DDP/ddp_model_wrapper.h:55-66)
  State f(const State &x, const Control &u) {
    for (int i = 0; i < StateSize; i++) state(i) = x(i);
    for (int i = 0; i < ControlSize; i++) control(i) = u(i);
    model_->computeKinematics(state);
    model_->computeDynamics(state, control);
    State dx;
    for (int i = 0; i < StateSize; i++) dx(i) = model_->state_der_(i);
    return dx;
  }

  Jacobian df(const State &x, const Control &u) { return df_impl(x, u, 0); }

 private:
  template <class M = DYNAMICS_T>
  auto df_impl(const State &x, const Control &u, int) -> decltype(std::declval<M &>().computeGrad(state, control), Jacobian()) {
    for (int i = 0; i < StateSize; i++) state(i) = x(i);
    for (int i = 0; i < ControlSize; i++) control(i) = u(i);
    model_->computeGrad(state, control);
    Jacobian j;
    for (int r = 0; r < StateSize; r++)
      for (int c = 0; c < StateSize + ControlSize; c++) j(r, c) = model_->jac_(r, c);
    return j;
  }
  // Central differences with the step rule of Eigen::NumericalDiff<..., Central>, which the reference uses for models
  // without computeGrad (DDP/ddp_dynamics.h:75-79): h = sqrt(epsilon) |v|, or sqrt(epsilon) when v == 0.
  Jacobian df_impl(const State &x, const Control &u, long) {
    Jacobian j;
    const float eps = std::sqrt(std::numeric_limits<float>::epsilon());
    for (int c = 0; c < StateSize + ControlSize; c++) {
      State xp = x, xm = x;
      Control up = u, um = u;
      const float v = c < StateSize ? x(c) : u(c - StateSize);
      float h = eps * std::fabs(v);
      if (h == 0.0f) h = eps;
      if (c < StateSize) { xp(c) += h; xm(c) -= h; } else { up(c - StateSize) += h; um(c - StateSize) -= h; }
      const State fp = f(xp, up), fm = f(xm, um);
      for (int r = 0; r < StateSize; r++) j(r, c) = (fp(r) - fm(r)) / (2.0f * h);
    }
    return j;
  }
};

template <class Dynamics>
struct OptimizerResult {
  typedef typename Dynamics::Scalar Scalar;
  int iterations = 0;
  int timesteps = 0;
  Scalar total_cost = 0;
  Eigen::MatrixXf cost;                ///< per-step cost of the accepted trajectory (1 x H)
  Eigen::MatrixXf state_trajectory;    ///< StateSize x H
  Eigen::MatrixXf control_trajectory;  ///< ControlSize x H
  std::vector<typename Dynamics::FeedbackGain> feedback_gain;  ///< H gains K_k (ControlSize x StateSize)
  Eigen::MatrixXf feedforward_gain;    ///< ControlSize x H
};

template <class DYNAMICS_T, class QMat, class RMat, class UVec>
OptimizerResult<ModelWrapperDDP<DYNAMICS_T> > ddp_feedback_gains(ModelWrapperDDP<DYNAMICS_T> &dyn, const Eigen::MatrixXf &x0,
                                                                 const std::vector<float> &target_x, const std::vector<float> &target_u,
                                                                 int H, float dt, const QMat &Q, const QMat &Qf, const RMat &R,
                                                                 const UVec &u_min, const UVec &u_max) {
  typedef ModelWrapperDDP<DYNAMICS_T> Dyn;
  const int N = Dyn::StateSize, M = Dyn::ControlSize;
  typedef typename Dyn::State State;
  typedef typename Dyn::Control Control;
  OptimizerResult<Dyn> out;
  out.iterations = 1;
  out.timesteps = H;
  std::vector<State> x(H);
  std::vector<Control> u(H);
  auto clampu = [&](Control &c) { for (int i = 0; i < M; i++) c(i) = std::max(u_min(i), std::min(u_max(i), c(i))); };
  for (int k = 0; k < H; k++)
    for (int i = 0; i < M; i++) u[k](i) = target_u[(size_t)M * k + i];
  for (int i = 0; i < N; i++) x[0](i) = x0(i);
  for (int i = 1; i < H; i++) {
    if (i < H - 1) clampu(u[i - 1]);
    const State d = dyn.f(x[i - 1], u[i - 1]);
    for (int j = 0; j < N; j++) x[i](j) = x[i - 1](j) + d(j) * dt;
  }
  // linearisation and cost derivatives
  std::vector<std::vector<float> > A(H, std::vector<float>(N * N)), B(H, std::vector<float>(N * M));
  std::vector<std::vector<float> > lx(H, std::vector<float>(N)), lu(H, std::vector<float>(M));
  for (int k = 0; k < H; k++) {
    const typename Dyn::Jacobian J = dyn.df(x[k], u[k]);
    for (int r = 0; r < N; r++) {
      for (int c = 0; c < N; c++) A[k][r * N + c] = J(r, c) * dt + (r == c ? 1.0f : 0.0f);
      for (int c = 0; c < M; c++) B[k][r * M + c] = J(r, N + c) * dt;
    }
    for (int r = 0; r < N; r++) {
      float acc = 0;
      for (int c = 0; c < N; c++) acc += Q(r, c) * (x[k](c) - target_x[(size_t)N * k + c]);
      lx[k][r] = acc;
    }
    for (int r = 0; r < M; r++) {
      float acc = 0;
      for (int c = 0; c < M; c++) acc += R(r, c) * (u[k](c) - target_u[(size_t)M * k + c]);
      lu[k][r] = acc;
    }
  }
  // boundary condition: terminal cost (x - x*_{H-1})' Qf (x - x*_{H-1})
  std::vector<float> Vxx(N * N), Vx(N);
  for (int r = 0; r < N; r++) {
    float acc = 0;
    for (int c = 0; c < N; c++) {
      Vxx[r * N + c] = 0.5f * (Qf(r, c) + Qf(c, r));
      acc += Qf(r, c) * (x[H - 1](c) - target_x[(size_t)N * (H - 1) + c]);
    }
    Vx[r] = acc;
  }
  out.feedback_gain.assign(H, typename Dyn::FeedbackGain());
  out.feedforward_gain = Eigen::MatrixXf::Zero(M, H);
  std::vector<float> VA(N * N), VB(N * M), qx(N), qu(M), qux(M * N), qxx(N * N), quu(M * M), K(M * N), kff(M);
  for (int k = H - 2; k >= 0; k--) {
    const std::vector<float> &Ak = A[k], &Bk = B[k];
    for (int r = 0; r < N; r++) {  // VA = Vxx A, VB = Vxx B
      for (int c = 0; c < N; c++) { float a = 0; for (int m = 0; m < N; m++) a += Vxx[r * N + m] * Ak[m * N + c]; VA[r * N + c] = a; }
      for (int c = 0; c < M; c++) { float a = 0; for (int m = 0; m < N; m++) a += Vxx[r * N + m] * Bk[m * M + c]; VB[r * M + c] = a; }
    }
    for (int r = 0; r < N; r++) { float a = lx[k][r] * dt; for (int m = 0; m < N; m++) a += Ak[m * N + r] * Vx[m]; qx[r] = a; }
    for (int r = 0; r < M; r++) { float a = lu[k][r] * dt; for (int m = 0; m < N; m++) a += Bk[m * M + r] * Vx[m]; qu[r] = a; }
    for (int r = 0; r < M; r++)
      for (int c = 0; c < N; c++) { float a = 0; for (int m = 0; m < N; m++) a += Bk[m * M + r] * VA[m * N + c]; qux[r * N + c] = a; }
    for (int r = 0; r < N; r++)
      for (int c = 0; c < N; c++) { float a = Q(r, c) * dt; for (int m = 0; m < N; m++) a += Ak[m * N + r] * VA[m * N + c]; qxx[r * N + c] = a; }
    for (int r = 0; r < M; r++)
      for (int c = 0; c < M; c++) { float a = R(r, c) * dt; for (int m = 0; m < N; m++) a += Bk[m * M + r] * VB[m * M + c]; quu[r * M + c] = a; }
    // solve quu [K | kff] = -[qux | qu] by Gaussian elimination with partial pivoting (M is tiny)
    std::vector<float> aug(M * (M + N + 1));
    const int W = M + N + 1;
    for (int r = 0; r < M; r++) {
      for (int c = 0; c < M; c++) aug[r * W + c] = quu[r * M + c];
      for (int c = 0; c < N; c++) aug[r * W + M + c] = -qux[r * N + c];
      aug[r * W + M + N] = -qu[r];
    }
    for (int p = 0; p < M; p++) {
      int best = p;
      for (int r = p + 1; r < M; r++) if (std::fabs(aug[r * W + p]) > std::fabs(aug[best * W + p])) best = r;
      if (std::fabs(aug[best * W + p]) < 1e-20f) { fprintf(stderr, "DDP: singular Quu\n"); return out; }
      if (best != p) for (int c = 0; c < W; c++) std::swap(aug[p * W + c], aug[best * W + c]);
      for (int r = 0; r < M; r++) {
        if (r == p) continue;
        const float fct = aug[r * W + p] / aug[p * W + p];
        for (int c = p; c < W; c++) aug[r * W + c] -= fct * aug[p * W + c];
      }
    }
    for (int r = 0; r < M; r++) {
      for (int c = 0; c < N; c++) K[r * N + c] = aug[r * W + M + c] / aug[r * W + r];
      kff[r] = aug[r * W + M + N] / aug[r * W + r];
    }
    for (int r = 0; r < M; r++) {
      for (int c = 0; c < N; c++) out.feedback_gain[k](r, c) = K[r * N + c];
      out.feedforward_gain(r, k) = kff[r];
    }
    // value function: Vxx = sym(qxx + qux' K), Vx = qx + qux' k
    std::vector<float> nV(N * N);
    for (int r = 0; r < N; r++)
      for (int c = 0; c < N; c++) { float a = qxx[r * N + c]; for (int m = 0; m < M; m++) a += qux[m * N + r] * K[m * N + c]; nV[r * N + c] = a; }
    for (int r = 0; r < N; r++) {
      for (int c = 0; c < N; c++) Vxx[r * N + c] = 0.5f * (nV[r * N + c] + nV[c * N + r]);
      float a = qx[r];
      for (int m = 0; m < M; m++) a += qux[m * N + r] * kff[m];
      Vx[r] = a;
    }
  }
  // forward pass, alpha = 1
  out.state_trajectory = Eigen::MatrixXf::Zero(N, H);
  out.control_trajectory = Eigen::MatrixXf::Zero(M, H);
  out.cost = Eigen::MatrixXf::Zero(1, H);
  State xn = x[0];
  for (int k = 0; k < H; k++) {
    for (int i = 0; i < N; i++) out.state_trajectory(i, k) = xn(i);
    if (k == H - 1) break;
    Control un;
    for (int r = 0; r < M; r++) {
      float a = u[k](r) + out.feedforward_gain(r, k);
      for (int c = 0; c < N; c++) a += out.feedback_gain[k](r, c) * (xn(c) - x[k](c));
      un(r) = a;
    }
    clampu(un);
    for (int i = 0; i < M; i++) out.control_trajectory(i, k) = un(i);
    float c = 0;
    for (int r = 0; r < N; r++) for (int cc = 0; cc < N; cc++) c += (xn(r) - target_x[(size_t)N * k + r]) * Q(r, cc) * (xn(cc) - target_x[(size_t)N * k + cc]);
    for (int r = 0; r < M; r++) for (int cc = 0; cc < M; cc++) c += (un(r) - target_u[(size_t)M * k + r]) * R(r, cc) * (un(cc) - target_u[(size_t)M * k + cc]);
    out.cost(0, k) = c * dt;
    out.total_cost += c * dt;
    const State d = dyn.f(xn, un);
    for (int j = 0; j < N; j++) xn(j) += d(j) * dt;
  }
  return out;
}

#endif
