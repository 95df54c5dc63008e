// dynamics.cuh -- device dynamics models for the fused rollout kernel.
//
//  * NetChain<R, 6,32,32,4>: the NeuralNetModel MLP (PI/neural_net_model.cu:357-410) for R rollouts
//    held in one thread's registers.  Weights are staged once per CTA in shared memory in a
//    transposed layout Wt[k][j], so for a fixed input k the OUT weights are contiguous: one
//    broadcast LDS.128 feeds 4*R FFMAs.  Per output neuron the products are accumulated for k
//    ascending from 0 with FMA and the bias is added afterwards -- the reference's order.
//  * CarBasisModel: GeneralizedLinear<CarBasisFuncs,7,2,25,CarKinematics,3>
//    (PI/generalized_linear.cu:225-245, PI/car_bfs.cuh:44-120) as straight-line code with the shared
//    sub-expressions (front slip tangent, rear slip ratio, sin(steer)) evaluated once.
#pragma once
#include "device_common.cuh"

namespace mppi {

// ---- packed transposed parameter layout ------------------------------------------------------
// For widths w0..wL: for each layer l: Wt_l[w_l][w_{l+1}] (k-major) followed by b_l[w_{l+1}].
template <int... W>
struct NetShape;
template <int A, int B2>
struct NetShape<A, B2> {
  static constexpr int NPARAMS = A * B2 + B2;
  static constexpr int LAST = B2;
  static constexpr int FIRST = A;
};
template <int A, int B2, int C, int... Rest>
struct NetShape<A, B2, C, Rest...> {
  static constexpr int NPARAMS = A * B2 + B2 + NetShape<B2, C, Rest...>::NPARAMS;
  static constexpr int LAST = NetShape<B2, C, Rest...>::LAST;
  static constexpr int FIRST = A;
};

template <int R, int IN, int OUT, bool ACT>
__device__ __forceinline__ void dense_layer(const float *__restrict__ sw, const float (&a)[IN][R], float (&o)[OUT][R]) {
  static_assert(OUT % 4 == 0, "layer width must be a multiple of 4");
  const float4 *W4 = reinterpret_cast<const float4 *>(sw);
#pragma unroll
  for (int j = 0; j < OUT; j++)
#pragma unroll
    for (int r = 0; r < R; r++) o[j][r] = 0.0f;
#pragma unroll
  for (int k = 0; k < IN; k++) {
#pragma unroll
    for (int j4 = 0; j4 < OUT / 4; j4++) {
      const float4 w = W4[k * (OUT / 4) + j4];
#pragma unroll
      for (int r = 0; r < R; r++) {
        o[4 * j4 + 0][r] = fmaf(w.x, a[k][r], o[4 * j4 + 0][r]);
        o[4 * j4 + 1][r] = fmaf(w.y, a[k][r], o[4 * j4 + 1][r]);
        o[4 * j4 + 2][r] = fmaf(w.z, a[k][r], o[4 * j4 + 2][r]);
        o[4 * j4 + 3][r] = fmaf(w.w, a[k][r], o[4 * j4 + 3][r]);
      }
    }
  }
  const float4 *B4 = reinterpret_cast<const float4 *>(sw + IN * OUT);
#pragma unroll
  for (int j4 = 0; j4 < OUT / 4; j4++) {
    const float4 b = B4[j4];
#pragma unroll
    for (int r = 0; r < R; r++) {
      float v0 = o[4 * j4 + 0][r] + b.x, v1 = o[4 * j4 + 1][r] + b.y;
      float v2 = o[4 * j4 + 2][r] + b.z, v3 = o[4 * j4 + 3][r] + b.w;
      o[4 * j4 + 0][r] = ACT ? tanh_fast(v0) : v0;
      o[4 * j4 + 1][r] = ACT ? tanh_fast(v1) : v1;
      o[4 * j4 + 2][r] = ACT ? tanh_fast(v2) : v2;
      o[4 * j4 + 3][r] = ACT ? tanh_fast(v3) : v3;
    }
  }
}

template <int R, int IN, int OUT, int... Rest>
struct NetChain {
  using Shape = NetShape<IN, OUT, Rest...>;
  static constexpr int NPARAMS = Shape::NPARAMS;
  static constexpr int LAST = Shape::LAST;
  __device__ __forceinline__ static void forward(const float *__restrict__ sw, const float (&a)[IN][R], float (&out)[LAST][R]) {
    if constexpr (sizeof...(Rest) == 0) {
      dense_layer<R, IN, OUT, false>(sw, a, out);
    } else {
      float h[OUT][R];
      dense_layer<R, IN, OUT, true>(sw, a, h);
      NetChain<R, OUT, Rest...>::forward(sw + IN * OUT + OUT, h, out);
    }
  }
};

// ---- packed variant: two rollouts per thread in f32x2 registers ---------------------------------
// sm_100 has a 2-wide FP32 FMA (PTX fma.rn.f32x2, SASS FFMA2) whose second source may be a single
// 32-bit register or uniform register broadcast to both halves.  With the two rollouts of a thread
// packed as (r0, r1) the scalar weight broadcasts for free, the FMA issue slots halve, and each half
// still accumulates k ascending with FMA, bias last -- bit-identical to the scalar layer.
// WSRC: 0 = weights from shared memory (LDS.128), 1 = from the constant bank (LDCU.128 -> uniform regs).
#ifndef MPPI_CONST_THETA_FLOATS
#define MPPI_CONST_THETA_FLOATS 1412
#endif
__constant__ float c_theta[MPPI_CONST_THETA_FLOATS];

template <int WSRC>
__device__ __forceinline__ float4 load_w4(const float *__restrict__ sw, int const_off, int idx4) {
  if constexpr (WSRC == 0) return reinterpret_cast<const float4 *>(sw)[idx4];
  else return reinterpret_cast<const float4 *>(c_theta + const_off)[idx4];
}

template <int IN, int OUT, bool ACT, int WSRC>
__device__ __forceinline__ void dense_layer_p2(const float *__restrict__ sw, int const_off, const float2 (&a)[IN], float2 (&o)[OUT]) {
  static_assert(OUT % 4 == 0, "layer width must be a multiple of 4");
#pragma unroll
  for (int j = 0; j < OUT; j++) o[j] = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int k = 0; k < IN; k++) {
#pragma unroll
    for (int j4 = 0; j4 < OUT / 4; j4++) {
      const float4 w = load_w4<WSRC>(sw, const_off, k * (OUT / 4) + j4);
      o[4 * j4 + 0] = __ffma2_rn(make_float2(w.x, w.x), a[k], o[4 * j4 + 0]);
      o[4 * j4 + 1] = __ffma2_rn(make_float2(w.y, w.y), a[k], o[4 * j4 + 1]);
      o[4 * j4 + 2] = __ffma2_rn(make_float2(w.z, w.z), a[k], o[4 * j4 + 2]);
      o[4 * j4 + 3] = __ffma2_rn(make_float2(w.w, w.w), a[k], o[4 * j4 + 3]);
    }
  }
#pragma unroll
  for (int j4 = 0; j4 < OUT / 4; j4++) {
    const float4 b = load_w4<WSRC>(sw + IN * OUT, const_off + IN * OUT, j4);
    o[4 * j4 + 0] = __fadd2_rn(o[4 * j4 + 0], make_float2(b.x, b.x));
    o[4 * j4 + 1] = __fadd2_rn(o[4 * j4 + 1], make_float2(b.y, b.y));
    o[4 * j4 + 2] = __fadd2_rn(o[4 * j4 + 2], make_float2(b.z, b.z));
    o[4 * j4 + 3] = __fadd2_rn(o[4 * j4 + 3], make_float2(b.w, b.w));
    if (ACT) {
      o[4 * j4 + 0] = tanh_fast2(o[4 * j4 + 0]);
      o[4 * j4 + 1] = tanh_fast2(o[4 * j4 + 1]);
      o[4 * j4 + 2] = tanh_fast2(o[4 * j4 + 2]);
      o[4 * j4 + 3] = tanh_fast2(o[4 * j4 + 3]);
    }
  }
}

template <int WSRC, int IN, int OUT, int... Rest>
struct NetChainP2 {
  using Shape = NetShape<IN, OUT, Rest...>;
  static constexpr int LAST = Shape::LAST;
  __device__ __forceinline__ static void forward(const float *__restrict__ sw, int const_off, const float2 (&a)[IN], float2 (&out)[LAST]) {
    if constexpr (sizeof...(Rest) == 0) {
      dense_layer_p2<IN, OUT, false, WSRC>(sw, const_off, a, out);
    } else {
      float2 h[OUT];
      dense_layer_p2<IN, OUT, true, WSRC>(sw, const_off, a, h);
      NetChainP2<WSRC, OUT, Rest...>::forward(sw + IN * OUT + OUT, const_off + IN * OUT + OUT, h, out);
    }
  }
};

template <int WSRC_, int... W>
struct NeuralNetDynP2 {
  static constexpr int R = 2;
  static constexpr int WSRC = WSRC_;
  static constexpr int SMEM_FLOATS = WSRC_ == 0 ? NetShape<W...>::NPARAMS : 0;
  static constexpr int NPARAMS = NetShape<W...>::NPARAMS;
  static constexpr int THREAD_SMEM_FLOATS = 0;
  __device__ __forceinline__ static void deriv(const float *__restrict__ sw, float *, const float (&in)[6][2], float (&out)[4][2]) {
    float2 a[6], o[4];
#pragma unroll
    for (int k = 0; k < 6; k++) a[k] = make_float2(in[k][0], in[k][1]);
    NetChainP2<WSRC_, W...>::forward(sw, 0, a, o);
#pragma unroll
    for (int k = 0; k < 4; k++) { out[k][0] = o[k].x; out[k][1] = o[k].y; }
  }
};

// ---- compact variant of the packed pair: layer 2 as a ROLLED loop over chunks of 8 outputs -----------------------
// ncu on the fully unrolled pair kernel (profiles/ncu_1m_r01b.txt) shows the instruction fetch as its first stall
// reason ("no_instruction", 21% of the warp cycles): one timestep is ~2900 straight-line instructions = 46 KB, far
// beyond the instruction caches, and 168 live registers allow only 3 warps per scheduler to hide it.  Here the 32 x 32
// layer runs as 4 iterations of one 8-output body (64 LDS.128 + 256 FFMA2 + 8 bias / tanh): its outputs go to a
// per-thread shared-memory column (dynamic index, which registers cannot do) from which layer 3 reads them back.
// k still ascends from 0 with FMA and the bias is added last, so results are bit-identical to NetChainP2.
// Cost: 32 STS.64 + 32 LDS.64 per timestep; gain: ~1000 fewer instructions of code and 48 fewer live registers.
template <int BLOCK>
struct NeuralNetDynP2Compact {
  static constexpr int R = 2;
  static constexpr int SMEM_FLOATS = NetShape<6, 32, 32, 4>::NPARAMS;
  static constexpr int THREAD_SMEM_FLOATS = 64;  // h2[32] as float2, laid out [j][thread]
  static constexpr int kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;
  __device__ __forceinline__ static void deriv(const float *__restrict__ sw, float *__restrict__ tsm, const float (&in)[6][2],
                                               float (&out)[4][2]) {
    float2 a[6], h1[32];
#pragma unroll
    for (int k = 0; k < 6; k++) a[k] = make_float2(in[k][0], in[k][1]);
    dense_layer_p2<6, 32, true, 0>(sw, 0, a, h1);
    float2 *h2 = reinterpret_cast<float2 *>(tsm);
#pragma unroll 1
    for (int c = 0; c < 4; c++) {
      const float4 *W = reinterpret_cast<const float4 *>(sw + kW2 + 8 * c);
      float2 o[8];
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < 32; k++) {
        const float4 wa = W[k * 8], wb = W[k * 8 + 1];
        o[0] = __ffma2_rn(make_float2(wa.x, wa.x), h1[k], o[0]); o[1] = __ffma2_rn(make_float2(wa.y, wa.y), h1[k], o[1]);
        o[2] = __ffma2_rn(make_float2(wa.z, wa.z), h1[k], o[2]); o[3] = __ffma2_rn(make_float2(wa.w, wa.w), h1[k], o[3]);
        o[4] = __ffma2_rn(make_float2(wb.x, wb.x), h1[k], o[4]); o[5] = __ffma2_rn(make_float2(wb.y, wb.y), h1[k], o[5]);
        o[6] = __ffma2_rn(make_float2(wb.z, wb.z), h1[k], o[6]); o[7] = __ffma2_rn(make_float2(wb.w, wb.w), h1[k], o[7]);
      }
      const float4 ba = *reinterpret_cast<const float4 *>(sw + kB2 + 8 * c), bb = *reinterpret_cast<const float4 *>(sw + kB2 + 8 * c + 4);
      const float bj[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int j = 0; j < 8; j++) h2[(8 * c + j) * BLOCK] = tanh_fast2(__fadd2_rn(o[j], make_float2(bj[j], bj[j])));
    }
    float2 y[4];
#pragma unroll
    for (int j = 0; j < 4; j++) y[j] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 32; k++) {
      const float2 g = h2[k * BLOCK];
      const float4 w = reinterpret_cast<const float4 *>(sw + kW3)[k];
      y[0] = __ffma2_rn(make_float2(w.x, w.x), g, y[0]); y[1] = __ffma2_rn(make_float2(w.y, w.y), g, y[1]);
      y[2] = __ffma2_rn(make_float2(w.z, w.z), g, y[2]); y[3] = __ffma2_rn(make_float2(w.w, w.w), g, y[3]);
    }
    const float4 b3 = *reinterpret_cast<const float4 *>(sw + kB3);
    y[0] = __fadd2_rn(y[0], make_float2(b3.x, b3.x)); y[1] = __fadd2_rn(y[1], make_float2(b3.y, b3.y));
    y[2] = __fadd2_rn(y[2], make_float2(b3.z, b3.z)); y[3] = __fadd2_rn(y[3], make_float2(b3.w, b3.w));
#pragma unroll
    for (int k = 0; k < 4; k++) { out[k][0] = y[k].x; out[k][1] = y[k].y; }
  }
};

// Dynamics policies consumed by rollout_kernel: deriv() maps [roll, u_x, u_y, yaw_rate, steer,
// throttle] to d/dt [roll, u_x, u_y, yaw_rate] for R rollouts.
template <int R_, int... W>
struct NeuralNetDyn {
  static constexpr int R = R_;
  static constexpr int SMEM_FLOATS = NetShape<W...>::NPARAMS;
  using Chain = NetChain<R_, W...>;
  static constexpr int THREAD_SMEM_FLOATS = 0;
  static_assert(NetShape<W...>::FIRST == 6 && NetShape<W...>::LAST == 4, "6 inputs, 4 outputs");
  __device__ __forceinline__ static void deriv(const float *__restrict__ sw, float *, const float (&in)[6][R_], float (&out)[4][R_]) {
    Chain::forward(sw, in, out);
  }
};

// x / c without the slow-path branch of the IEEE division sequence (FCHK + call), which serialises the 22 divisions of
// the basis functions behind each other.  Markstein's refinement: with rc = RN(1 / c), q = RN(x rc), r = x - q c (exact,
// one FMA), RN(q + r rc) is the correctly rounded quotient (outside overflow / underflow), i.e. the bits of x / c.
__device__ __forceinline__ float div_refined(float x, float c, float rc) {
  const float q = __fmul_rn(x, rc);
  const float r = fmaf(-q, c, x);
  return fmaf(r, rc, q);
}
// divisor of basis function i (PI/car_bfs.cuh), folded into the staged weights
constexpr double kBfDivisor[25] = {1.0, 10.0, 1200.0, 1440000.0, 1728000000.0, 25.0, 10.0, 10.0, 1.0, 40.0, 1400.0, 1960000.0, 2744000000.0,
                                   40.0, 1600.0, 64000.0, 50.0, 1.0, 1.0, 3.0, 5.0, 100.0, 1000.0, 1.0, 1.0};

struct CarBasisDyn {
  static constexpr int R = 1;
  static constexpr int SMEM_FLOATS = 100;  // theta TRANSPOSED and pre-divided [25][4] (mppi_set_bf_params): one LDS.128 feeds the four outputs of a basis function
  static constexpr int THREAD_SMEM_FLOATS = 0;
  __device__ __forceinline__ static void deriv(const float *__restrict__ sw, float *, const float (&in)[6][1], float (&out)[4][1]) {
    const float roll = in[0][0], vx = in[1][0], vy = in[2][0], wz = in[3][0], steer = in[4][0], thr = in[5][0];
    const bool moving = vx >= 0.1f;  // (double)vx > .1  <=>  vx >= 0.1f because 0.1f > 0.1
    // the three divisions by u_x share one reciprocal (MUFU.RCP + one Newton step: within 1 ulp, then Markstein's step)
    const float r0 = rcp_approx(vx);
    const float rvx = fmaf(fmaf(-vx, r0, 1.0f), r0, r0);
    const float ratio_y = div_refined(vy, vx, rvx);
    // front: tan(atan(vy/vx + .45 wz/vx) - steer); rear: vy/vx - .35 wz/vx
    const float front_arg = moving ? (ratio_y + div_refined(0.45f * wz, vx, rvx)) : 0.0f;
    // tan(atan(a) - s) = (a - tan s) / (1 + a tan s): tan(steer) depends on the control only, i.e. it is off the critical path of
    // the state recursion, and the atanf / tanf pair that sat on it (~100 instructions, ~230 cycles) becomes one refined division.
    // The two forms differ by a few float ulps (each of atanf, tanf is accurate to 2-4 ulp itself).
    // sin and tan of the steering angle from ONE sincosf (tan = sin / cos by a refined division; |steer| <= 1 after the clamp)
    float ss, cs;
    sincosf(steer, &ss, &cs);
    const float c0 = rcp_approx(cs);
    const float ts = div_refined(ss, cs, fmaf(fmaf(-cs, c0, 1.0f), c0, c0));
    float tf = -ts;
    if (moving) {
      const float den = fmaf(front_arg, ts, 1.0f);
      const float d0 = rcp_approx(den);
      tf = div_refined(front_arg - ts, den, fmaf(fmaf(-den, d0, 1.0f), d0, d0));
    }
    const float rear = ratio_y - div_refined(0.35f * wz, vx, rvx);
    // The constant divisors of the basis functions (PI/car_bfs.cuh: u_x / 10, sin(d) tan(a_f) / 1200, ...) are folded into the
    // staged weights (theta_t[i][j] = theta[j][i] / c_i, kBfDivisor, mppi_set_bf_params): phi here are the numerators, and the
    // 22 refined divisions of a timestep (66 instructions, 15 % of the kernel) are gone.  RN(w / c) x instead of w RN(x / c):
    // one float ulp per term.
    float phi[25];
    phi[0] = thr;
    phi[1] = vx;
    phi[2] = ss * tf;
    phi[3] = ss * tf * fabsf(tf);
    phi[4] = ss * (tf * tf * tf);
    phi[5] = wz * vy;
    phi[6] = wz;
    phi[7] = vy;
    phi[8] = ss;
    phi[9] = moving ? ratio_y : 0.0f;
    phi[10] = tf;
    phi[11] = tf * fabsf(tf);
    phi[12] = tf * tf * tf;
    phi[13] = moving ? rear : 0.0f;
    phi[14] = moving ? rear * fabsf(rear) : 0.0f;
    phi[15] = moving ? rear * rear * rear : 0.0f;
    phi[16] = wz * vx;
    phi[17] = roll;
    phi[18] = roll * wz;
    phi[19] = roll * vx;
    phi[20] = roll * vx * wz;
    phi[21] = vx * vx;
    phi[22] = vx * vx * vx;
    phi[23] = thr * thr;
    phi[24] = thr * thr * thr;
    // theta . phi, i ascending per output (the order of PI/generalized_linear.cu:225-245 with one y-thread); the weights of
    // basis function i for the four outputs arrive as one 128-bit load (100 scalar loads were 22 % of the kernel's instructions)
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    const float4 *sw4 = reinterpret_cast<const float4 *>(sw);
#pragma unroll
    for (int i = 0; i < 25; i++) {
      const float4 w = sw4[i];
      a0 = fmaf(w.x, phi[i], a0); a1 = fmaf(w.y, phi[i], a1); a2 = fmaf(w.z, phi[i], a2); a3 = fmaf(w.w, phi[i], a3);
    }
    out[0][0] = a0; out[1][0] = a1; out[2][0] = a2; out[3][0] = a3;
  }
};

}  // namespace mppi
