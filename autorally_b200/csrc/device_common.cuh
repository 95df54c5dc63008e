// device_common.cuh -- shared device-side types and math for the MPPI kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mppi {

constexpr int S_DIM = 7;
constexpr int C_DIM = 2;

// Device copy of MPPICosts::CostParams (PI/costs.cuh:67-85).  It travels in the kernel parameter
// block (constant bank), so every access is a uniform constant read; the reference re-reads it
// through a global pointer (params_d_->...) on every use.
struct DevCostParams {
  float desired_speed, speed_coeff, track_coeff, max_slip_ang, slip_penalty, track_slop, crash_coeff;
  float steering_coeff, throttle_coeff, boundary_threshold;
  float crash_cost_on;  // (float)((1.0 - (double)discount) * (double)crash_coeff), PI/costs.cu:402
  int l1_cost;
  int has_control_cost;  // steering_coeff != 0 || throttle_coeff != 0
  int affine;            // r_c1.z == 0 && r_c2.z == 0 && trs.z == 1: the projective divide is the identity
  float c1x, c1y, c1z, c2x, c2y, c2z, tx, ty, tz;  // r_c1, r_c2, trs
};

// Everything one rollout launch needs; passed by value (__grid_constant__).
struct RolloutParams {
  const float *inbox;      // [B][inbox_stride]: state[7] | hist[4] | pad[1] | U[T][2]
  float *du;               // [B][n_local][T][2]  noise in, sampled (un-clamped) controls out
  float *costs;            // [B][n_local]
  unsigned char *crash;    // [B][n_local]
  unsigned int *baseline;  // [B] order-preserving uint encoding of the minimum cost
  const float *theta_t;    // packed transposed weights (see NetLayout in dynamics_nn.cuh) / BF theta
  const float *theta_fold; // 6-32-32-4 only: the same layout with the tanh scale and affine map folded in (fold_nn32, mppi_b200.cu)
  const double *inv_step;  // [T]: 1.0 / (1.0 * i)
  int inbox_stride;
  int n_local, n_global, r_begin, T, B, opt_delay, pure_noise_from;
  float nu0, nu1, lo0, hi0, lo1, hi1, dt;
  int negate_yaw;
  DevCostParams cp;
  cudaTextureObject_t tex;
  // Noise generated in place (philox.cuh): fused_noise != 0 -> the kernel draws eps[r][t][:] itself from the Philox stream
  // (same counters as sample_noise_kernel, bit-identical values) instead of reading it from `du`; `du` then only receives
  // the sampled controls.  call_ptr is the device-resident compute-call counter finalize_kernel advances.
  int fused_noise, b_begin;
  uint32_t seed_lo, seed_hi;
  const uint32_t *call_ptr;
};

constexpr int INBOX_STATE = 0;
constexpr int INBOX_HIST = 7;
constexpr int INBOX_U = 12;

// Programmatic dependent launch (PDL, sm_90+): a kernel launched with the programmatic-stream-serialization attribute
// may start while its predecessor is still running; pdl_wait() blocks until the predecessor grid has completed and its
// writes are visible, pdl_trigger() lets the successor grid be scheduled early.  Both are no-ops for a normal launch.
// Used to overlap each kernel's prologue (weights into registers, launch latency) with the tail of the previous one.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// Order-preserving float <-> uint map so the minimum cost can be taken with integer atomics / redux.
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) on MUFU.EX2 + MUFU.RCP: 3 FP32 + 2 SFU instructions, absolute error
// < 4e-7 over the whole range (tests/test_tanh_accuracy), saturating to +-1 without branches.  The
// reference calls CUDA's precise tanhf (2 ulp); tanh.approx.f32 (2^-11) would not hold the 1e-4
// parity tolerance over 100 recurrent steps.
__device__ __forceinline__ float tanh_fast(float x) {
  float e = ex2_approx(x * 2.88539008177792681472f);  // 2 * log2(e)
  float r = rcp_approx(e + 1.0f);
  return fmaf(-2.0f, r, 1.0f);
}

// Packed pair version (two rollouts in one 64-bit register pair): FMUL2 / FADD2 / FFMA2 halve the
// FP32 issue slots; the two MUFU calls per element stay scalar.
// r = 1 / (2^x + 1): the core of tanh(y) = 1 - 2 r with x = 2 log2(e) y.  The latency kernels of the 6-32-32-4 network take x
// from weights that already carry the factor 2 log2(e) and feed r, not tanh, to the next layer, whose weights carry the map
// 1 - 2 r (W h + b = (b + rowsum W) - 2 W r): three dependent instructions per activation instead of five.
__device__ __forceinline__ float recip_core(float x) { return rcp_approx(ex2_approx(x) + 1.0f); }
__device__ __forceinline__ float2 recip_core2(float2 x) {
  const float2 d = __fadd2_rn(make_float2(ex2_approx(x.x), ex2_approx(x.y)), make_float2(1.0f, 1.0f));
  return make_float2(rcp_approx(d.x), rcp_approx(d.y));
}

__device__ __forceinline__ float2 tanh_fast2(float2 x) {
  const float2 t = __fmul2_rn(x, make_float2(2.88539008177792681472f, 2.88539008177792681472f));
  const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  const float2 r = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  return __ffma2_rn(make_float2(-2.0f, -2.0f), r, make_float2(1.0f, 1.0f));
}

}  // namespace mppi
