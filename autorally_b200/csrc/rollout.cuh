// rollout.cuh -- the fused rollout kernel: control perturbation, clamp, running-mean cost, dynamics,
// Euler step and crash bookkeeping for T sequential steps, entirely in registers.
// Replaces rolloutKernel (PI/mppi_controller.cu:72-184) and the device members it calls
// (PI/neural_net_model.cu:311-410, PI/generalized_linear.cu:168-245, PI/costs.cu:301-409).
#pragma once
#include "device_common.cuh"
#include "dynamics.cuh"
#include "philox.cuh"

namespace mppi {

// The pieces of MPPICosts::computeCost (PI/costs.cu:396-409, parts :307-393) that do not depend on the sticky crash
// flag.  Two exact shortcuts, both uniform branches on kernel parameters:
//   * both control-cost coefficients zero (the launch-file default): the term is (+-0)/nu^2 added to +0 = +0;
//   * an affine costmap transform (r_c1.z = r_c2.z = 0, trs.z = 1, what loadTrackData builds, PI/costs.cu:226-229):
//     w is exactly 1 and u/w, v/w are u, v, so the four IEEE divisions are skipped.
struct StepCostParts {
  float pre;    // control + speed (summed before the crash cost, PI/costs.cu:403)
  float track;  // track cost
  float stab;   // stabilizing cost
  bool boundary;
};

__device__ __forceinline__ float costmap_lookup(const DevCostParams &cp, cudaTextureObject_t tex, float x, float y) {
  const float uu = __fadd_rn(fmaf(cp.c1x, x, __fmul_rn(cp.c2x, y)), cp.tx);  // coorTransform, PI/costs.cu:351-357
  const float vv = __fadd_rn(fmaf(cp.c1y, x, __fmul_rn(cp.c2y, y)), cp.ty);
  if (cp.affine) return tex2D<float>(tex, uu, vv);
  const float ww = __fadd_rn(fmaf(cp.c1z, x, __fmul_rn(cp.c2z, y)), cp.tz);
  return tex2D<float>(tex, __fdiv_rn(uu, ww), __fdiv_rn(vv, ww));
}

// The two costmap fetches of the track cost (:359-380): front / back of the car with the fast intrinsics the reference
// uses.  Separate from the rest so that a kernel can issue them early and consume the texels later.
__device__ __forceinline__ void track_lookups(const DevCostParams &cp, cudaTextureObject_t tex, float x, float y, float yaw,
                                              float &front, float &back) {
  const float cy = __cosf(yaw), sy = __sinf(yaw);
  front = costmap_lookup(cp, tex, fmaf(0.5f, cy, x), fmaf(0.5f, sy, y));
  back = costmap_lookup(cp, tex, fmaf(-0.5f, cy, x), fmaf(-0.5f, sy, y));
}

// control cost (:307-313); depends on the controls only
__device__ __forceinline__ float cost_control_part(const DevCostParams &cp, float u0, float u1, float du0, float du1, float nu0, float nu1) {
  float control = 0.0f;
  if (cp.has_control_cost) {
    control = __fadd_rn(control, __fdiv_rn(__fmul_rn(__fmul_rn(cp.steering_coeff, du0), __fsub_rn(u0, du0)), __fmul_rn(nu0, nu0)));
    control = __fadd_rn(control, __fdiv_rn(__fmul_rn(__fmul_rn(cp.throttle_coeff, du1), __fsub_rn(u1, du1)), __fmul_rn(nu1, nu1)));
  }
  return control;
}

// control cost + speed cost (:315-326): the part summed before the crash cost
__device__ __forceinline__ float cost_pre_part(const DevCostParams &cp, float control, float vx) {
  const float err = __fsub_rn(vx, cp.desired_speed);
  const float sc = cp.l1_cost ? fabsf(err) : __fmul_rn(err, err);
  return __fadd_rn(control, __fmul_rn(cp.speed_coeff, sc));
}

// track cost from the two texels (:381-393) and whether the boundary was hit
__device__ __forceinline__ float cost_track_part(const DevCostParams &cp, float front, float back, bool &boundary) {
  const float track = __fmul_rn(__fadd_rn(fabsf(front), fabsf(back)), 0.5f);  // "/2.0" in double is exact
  boundary = (front >= cp.boundary_threshold || back >= cp.boundary_threshold);
  return (fabsf(track) < cp.track_slop) ? 0.0f : __fmul_rn(cp.track_coeff, track);
}

// stabilizing cost (:337-349); the reference compares |u_x| against the double 0.001
__device__ __forceinline__ float cost_stab_part(const DevCostParams &cp, float vx, float vy) {
  float stab = 0.0f;
  if (fabsf(vx) >= 0.001f) {  // |u_x| > 0.001 (double)  <=>  |u_x| >= 0.001f because 0.001f > 0.001
    const float slip = -atanf(__fdiv_rn(vy, fabsf(vx)));
    stab = __fmul_rn(cp.slip_penalty, __fmul_rn(slip, slip));
    if (fabsf(slip) > cp.max_slip_ang) stab = __fadd_rn(stab, cp.crash_coeff);
  }
  return stab;
}

__device__ __forceinline__ StepCostParts step_cost_from_lookups(const DevCostParams &cp, float front, float back, float vx, float vy,
                                                                float u0, float u1, float du0, float du1, float nu0, float nu1) {
  StepCostParts r;
  r.track = cost_track_part(cp, front, back, r.boundary);
  r.pre = cost_pre_part(cp, cost_control_part(cp, u0, u1, du0, du1, nu0, nu1), vx);
  r.stab = cost_stab_part(cp, vx, vy);
  return r;
}

__device__ __forceinline__ StepCostParts step_cost_parts(const DevCostParams &cp, cudaTextureObject_t tex, float x, float y,
                                                         float yaw, float vx, float vy, float u0, float u1, float du0,
                                                         float du1, float nu0, float nu1) {
  float front, back;
  track_lookups(cp, tex, x, y, yaw, front, back);
  return step_cost_from_lookups(cp, front, back, vx, vy, u0, u1, du0, du1, nu0, nu1);
}

// computeCost given the parts: the crash flag is sticky, the boundary hit is charged in the same step (:396-409)
__device__ __forceinline__ float running_cost_from_parts(const DevCostParams &cp, const StepCostParts &c, int &crash) {
  if (c.boundary) crash = 1;
  const float crash_cost = crash > 0 ? cp.crash_cost_on : 0.0f;  // (:328-335, :402)
  float cost = __fadd_rn(__fadd_rn(__fadd_rn(c.pre, crash_cost), c.track), c.stab);
  if (cost > 1e12f || isnan(cost)) cost = 1e12f;
  return cost;
}

// MPPICosts::computeCost: the track cost runs before the crash cost, so a boundary hit is charged in the same
// step; `crash` is sticky.
__device__ __forceinline__ float running_cost_step(const DevCostParams &cp, cudaTextureObject_t tex,
                                                   const float (&s)[S_DIM], float u0, float u1, float du0,
                                                   float du1, float nu0, float nu1, int &crash) {
  const StepCostParts c = step_cost_parts(cp, tex, s[0], s[1], s[2], s[4], s[5], u0, u1, du0, du1, nu0, nu1);
  return running_cost_from_parts(cp, c, crash);
}

// One thread owns DYN::R consecutive rollouts.  Grid covers B * n_local rollouts; n_local is a
// multiple of 64, so a warp never straddles two controllers and is either fully valid or idle.
// FUSED: the noise is drawn in place from the Philox stream (philox.cuh) instead of being read from `du`.
template <class DYN, int BLOCK, int MINB = 1, bool FUSED = false>
__global__ void __launch_bounds__(BLOCK, MINB) rollout_kernel(const __grid_constant__ RolloutParams p) {
  constexpr int R = DYN::R;
  extern __shared__ float4 smem4[];
  float *sw = reinterpret_cast<float *>(smem4);
  for (int i = threadIdx.x; i < DYN::SMEM_FLOATS / 4; i += BLOCK) smem4[i] = reinterpret_cast<const float4 *>(p.theta_t)[i];
  __syncthreads();
  // per-thread staging column (policies that need dynamically indexed storage), [slot][thread] after the weights
  float *tsm = sw + ((DYN::SMEM_FLOATS + 3) & ~3) + (DYN::THREAD_SMEM_FLOATS ? 2 * threadIdx.x : 0);

  const long long total = (long long)p.B * p.n_local;
  const long long g0 = ((long long)blockIdx.x * BLOCK + threadIdx.x) * R;
  const bool valid = g0 < total;
  unsigned int best = 0xffffffffu;
  int ctrl = 0;
  if (valid) {
    ctrl = (int)(g0 / p.n_local);
    const int lr0 = (int)(g0 - (long long)ctrl * p.n_local);
    const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
    const float2 *U = reinterpret_cast<const float2 *>(inbox + INBOX_U);
    float s[R][S_DIM];
    float running[R];
    int crash[R];
    bool noise_free[R], pure_noise[R];
    float2 *row[R];
    FusedNoise fz[R];
    const uint32_t call = FUSED ? *p.call_ptr : 0u;
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
      for (int k = 0; k < S_DIM; k++) s[r][k] = inbox[INBOX_STATE + k];
      running[r] = 0.0f;
      crash[r] = 0;
      const int rg = p.r_begin + lr0 + r;  // global rollout index drives the bookkeeping (R2)
      noise_free[r] = (rg == 0);
      pure_noise[r] = (rg >= p.pure_noise_from);
      row[r] = reinterpret_cast<float2 *>(p.du) + (size_t)(g0 + r) * p.T;
      fz[r].r_global = (uint32_t)rg; fz[r].b_global = (uint32_t)(p.b_begin + ctrl); fz[r].call = call;
      fz[r].seed_lo = p.seed_lo; fz[r].seed_hi = p.seed_hi; fz[r].held = make_float2(0.0f, 0.0f);
    }
    // one rollout per thread (the latency-bound shapes: basis functions, 64-wide FP32 kernel): fetch the costmap texels a
    // step ahead; the two-rollouts-per-thread FFMA2 kernel is register-bound and keeps the plain order
    constexpr bool EARLY_TEXELS = (R == 1);
    float front[R], back[R];
#pragma unroll
    for (int r = 0; r < R; r++) { front[r] = 0.0f; back[r] = 0.0f; }
    float2 e_next[R];
#pragma unroll
    for (int r = 0; r < R; r++) e_next[r] = (EARLY_TEXELS && !FUSED) ? row[r][0] : make_float2(0.0f, 0.0f);
    for (int i = 0; i < p.T; i++) {
      const float2 Ui = U[i];
      float in[6][R];
      float du[R][2], u[R][2];
#pragma unroll
      for (int r = 0; r < R; r++) {
        // PI/mppi_controller.cu:130-155
        float2 e;
        if (FUSED) {
          e = fz[r].step(i);
        } else if (EARLY_TEXELS) {  // one rollout per thread: the noise of the next step is requested a step ahead as well
          e = e_next[r];
          if (i + 1 < p.T) e_next[r] = row[r][i + 1];
        } else {
          e = row[r][i];
        }
        if (noise_free[r] || i < p.opt_delay) {
          du[r][0] = 0.0f; du[r][1] = 0.0f;
          u[r][0] = Ui.x; u[r][1] = Ui.y;
        } else if (pure_noise[r]) {
          du[r][0] = __fmul_rn(e.x, p.nu0); du[r][1] = __fmul_rn(e.y, p.nu1);
          u[r][0] = du[r][0]; u[r][1] = du[r][1];
        } else {
          du[r][0] = __fmul_rn(e.x, p.nu0); du[r][1] = __fmul_rn(e.y, p.nu1);
          u[r][0] = __fadd_rn(Ui.x, du[r][0]); u[r][1] = __fadd_rn(Ui.y, du[r][1]);
        }
        row[r][i] = make_float2(u[r][0], u[r][1]);  // un-clamped write-back (:153)
        // enforceConstraints, PI/neural_net_model.cu:311-323
        u[r][0] = u[r][0] < p.lo0 ? p.lo0 : (u[r][0] > p.hi0 ? p.hi0 : u[r][0]);
        u[r][1] = u[r][1] < p.lo1 ? p.lo1 : (u[r][1] > p.hi1 ? p.hi1 : u[r][1]);
        if (i > 0) {
          // running mean of the step costs, PI/mppi_controller.cu:162-165: float difference,
          // double divide (by a tabulated reciprocal) and double accumulate, float store.
          float c;
          if (EARLY_TEXELS)
            c = running_cost_from_parts(p.cp, step_cost_from_lookups(p.cp, front[r], back[r], s[r][4], s[r][5], u[r][0], u[r][1], du[r][0],
                                                                      du[r][1], p.nu0, p.nu1), crash[r]);
          else
            c = running_cost_step(p.cp, p.tex, s[r], u[r][0], u[r][1], du[r][0], du[r][1], p.nu0, p.nu1, crash[r]);
          running[r] = (float)((double)running[r] + (double)__fsub_rn(c, running[r]) * p.inv_step[i]);
        }
        in[0][r] = s[r][3]; in[1][r] = s[r][4]; in[2][r] = s[r][5]; in[3][r] = s[r][6];
        in[4][r] = u[r][0]; in[5][r] = u[r][1];
      }
      if (EARLY_TEXELS) {
        // x, y, yaw do not depend on the dynamics model: they are advanced first (same operations, same order), so that the
        // two costmap texels of the NEXT step's cost are in flight during the whole dynamics evaluation
#pragma unroll
        for (int r = 0; r < R; r++) {
          float sn, cs;
          sincosf(s[r][2], &sn, &cs);
          const float d0 = fmaf(cs, s[r][4], -__fmul_rn(sn, s[r][5]));
          const float d1 = fmaf(sn, s[r][4], __fmul_rn(cs, s[r][5]));
          const float d2 = p.negate_yaw ? -s[r][6] : s[r][6];
          s[r][0] = fmaf(d0, p.dt, s[r][0]);
          s[r][1] = fmaf(d1, p.dt, s[r][1]);
          s[r][2] = fmaf(d2, p.dt, s[r][2]);
          if (i + 1 < p.T) track_lookups(p.cp, p.tex, s[r][0], s[r][1], s[r][2], front[r], back[r]);
        }
      }
      float dyn_out[4][R];
      DYN::deriv(sw, tsm, in, dyn_out);
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (!EARLY_TEXELS) {
          // kinematics, PI/neural_net_model.cu:346-355 (precise sinf/cosf)
          float sn, cs;
          sincosf(s[r][2], &sn, &cs);
          const float d0 = fmaf(cs, s[r][4], -__fmul_rn(sn, s[r][5]));
          const float d1 = fmaf(sn, s[r][4], __fmul_rn(cs, s[r][5]));
          const float d2 = p.negate_yaw ? -s[r][6] : s[r][6];
          // incrementState, PI/neural_net_model.cu:334-344
          s[r][0] = fmaf(d0, p.dt, s[r][0]);
          s[r][1] = fmaf(d1, p.dt, s[r][1]);
          s[r][2] = fmaf(d2, p.dt, s[r][2]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) s[r][3 + k] = fmaf(dyn_out[k][r], p.dt, s[r][3 + k]);
        // getCrash, PI/costs.cu:301-305 (the reference compares against the double 1.57)
        if (fabsf(s[r][3]) >= 1.57f) crash[r] = 1;  // > 1.57 (double)  <=>  >= 1.57f because 1.57f > 1.57
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      p.costs[g0 + r] = running[r];  // + terminalCost == 0 (PI/costs.cu:411-414)
      p.crash[g0 + r] = (unsigned char)crash[r];
      const unsigned int o = float_to_ordered(running[r]);
      best = o < best ? o : best;
    }
  }
  // min-cost baseline (host loop at PI/mppi_controller.cu:627-632): one redux + one atomic per warp
  const unsigned int wbest = __reduce_min_sync(0xffffffffu, best);
  const int wctrl = __shfl_sync(0xffffffffu, ctrl, 0);
  if ((threadIdx.x & 31) == 0 && wbest != 0xffffffffu) atomicMin(p.baseline + wctrl, wbest);
}

}  // namespace mppi
