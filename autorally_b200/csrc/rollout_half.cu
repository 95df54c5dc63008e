// rollout_half.cu -- the latency kernel for NeuralNetModel<7,2,3,6,32,32,4>: ONE ROLLOUT PER HALF-WARP (16 lanes),
// two rollouts per warp, one warp per CTA.
//
// ncu on the earlier latency kernels (profiles/ncu_1920_r01*.txt) shows a warp spends ~5.5 cycles per issued
// instruction whatever the decomposition (dependent FMA chains, MUFU, shared-memory exchanges, in-order issue), so
// the time of one timestep is set by the NUMBER OF INSTRUCTIONS a warp issues for it.  This kernel minimises that:
//
//  * lane l of a half-warp owns hidden neurons 2l, 2l+1 of both hidden layers, all their weights in registers as
//    float2 pairs, so every MLP multiply-add is a 2-wide FFMA2 with the activation broadcast from a scalar register:
//    layer 2 is 32 FFMA2 per lane (4 accumulator pairs, 8-deep chains) instead of 64 FFMA;
//  * layer 3 is (4 outputs) x (4 quarters of k) over the 16 lanes: 4 FFMA2, then the 16 partial sums meet in shared memory
//    (one store, four broadcast loads, two levels of adds: a shorter chain than a two-level xor tree plus a broadcast);
//  * the weights are the folded copy (fold_nn32, mppi_b200.cu): tanh's factor 2 log2(e) and the map 1 - 2 r live in the
//    weights, the activations that cross lanes are r = 1 / (2^x + 1) -- three dependent instructions instead of five;
//  * the two rollouts of a warp share every instruction (one LDS / STS / SHFL / MUFU serves both);
//  * per block of 16 timesteps, lane l prepares timestep i0 + l in parallel (noise, perturbation, un-clamped
//    write-back, clamp; PI/mppi_controller.cu:130-159) and evaluates its running cost afterwards in parallel
//    (positions by a sequential FMA prefix, sincosf, costmap fetches, PI/costs.cu:307-393; sticky crash flag as a
//    prefix-OR over ballots);
//  * the step costs go to shared memory and their mean (the reference's running-mean recursion,
//    PI/mppi_controller.cu:162-165) is taken once at the end, in double, by the 16 lanes in parallel: the recursion's
//    float -> double -> float chain stalled the in-order pipeline when interleaved with the MLP (14% of the stall samples)
//    and cost 2.8 us when replayed serially at the end.
//
// Layers 2 and 3 sum k in four interleaved partial sums (see warp_mlp.cuh for the numerical note).
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {
constexpr int kW1 = 0, kB1 = 192, kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;  // packed transposed layout

__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

// MINB = 12 one-warp CTAs per SM (166 registers) for up to 3552 rollouts: with a tighter cap ptxas funnels the eight
// LDS.128 of a layer through one register quad and every group of four FFMA2 waits a full shared-memory latency
// (1920 rollouts: 55.4 us capped at 128 registers, 49.6 us at 166).  MINB = 16 (128 registers) keeps larger rollout
// counts in fewer waves, which matters more there (4096 rollouts: 70 us vs 90 us).
template <int MINB>
__global__ void __launch_bounds__(32, MINB) rollout_half_kernel(const __grid_constant__ RolloutParams p) {
  extern __shared__ float4 smem4[];
  float *xbuf = reinterpret_cast<float *>(smem4);  // per half-warp: h1[32], h2[32]  -> 128 floats per warp
  const int lane = threadIdx.x, l = lane & 15, hw = lane >> 4, gb = lane & 16;
  float *myx = xbuf + hw * 64;
  float *pbuf = xbuf + 128 + hw * 16;               // per half-warp: the 16 partial sums of the output layer
  const int T = p.T;
  float *scost = xbuf + 160 + hw * T;               // [2][T] step costs for the deferred running mean
  const unsigned full = 0xffffffffu;
  const long long gro = (long long)blockIdx.x * 2 + hw;  // rollout index over B * n_local (even, so both are valid)
  const int ctrl = (int)(gro / p.n_local);
  const int lr = (int)(gro - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;

  // ---- lane-resident weight slices: neurons (2l, 2l+1) of layers 1 and 2, output (l & 3) over k in [8q, 8q+8) ----
  const float *th = p.theta_fold;  // tanh scale and affine map folded into the weights (fold_nn32): activations travel as r
  float2 w1[6], w2[32];
#pragma unroll
  for (int k = 0; k < 6; k++) w1[k] = *reinterpret_cast<const float2 *>(th + kW1 + k * 32 + 2 * l);
#pragma unroll
  for (int k = 0; k < 32; k++) w2[k] = *reinterpret_cast<const float2 *>(th + kW2 + k * 32 + 2 * l);
  const float2 b1 = *reinterpret_cast<const float2 *>(th + kB1 + 2 * l);
  const float2 b2 = *reinterpret_cast<const float2 *>(th + kB2 + 2 * l);
  const int jo = l & 3, q = l >> 2;
  float2 w3[4];  // (k, k+1) pairs of this lane's quarter
#pragma unroll
  for (int m = 0; m < 4; m++) w3[m] = make_float2(th[kW3 + (8 * q + 2 * m) * 4 + jo], th[kW3 + (8 * q + 2 * m + 1) * 4 + jo]);
  const float2 b3q = make_float2(q == 0 ? th[kB3 + jo] : 0.0f, 0.0f);  // the output bias opens the partial sum of the first quarter

  const float2 *Ug = reinterpret_cast<const float2 *>(inbox + INBOX_U);
  float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)gro * T;
  pdl_trigger();
  pdl_wait();  // everything above reads model parameters only; noise and inbox come from the sampler kernel
  float xcur = inbox[INBOX_STATE + 0], ycur = inbox[INBOX_STATE + 1], yaw = inbox[INBOX_STATE + 2];
  float roll = inbox[INBOX_STATE + 3], vx = inbox[INBOX_STATE + 4], vy = inbox[INBOX_STATE + 5], wz = inbox[INBOX_STATE + 6];
  const int rg = p.r_begin + lr;  // the GLOBAL rollout index drives the bookkeeping (R2)
  const bool noise_free = (rg == 0), pure_noise = (rg >= p.pure_noise_from);
  bool crash_in = false;
  // noise and nominal control of this lane's timestep, fetched one block ahead of their use
  float2 e_next = l < T ? row[l] : make_float2(0.0f, 0.0f);
  float2 U_next = l < T ? Ug[l] : make_float2(0.0f, 0.0f);

  for (int i0 = 0; i0 < T; i0 += 16) {
    const int nb = min(16, T - i0);
    const bool mine = l < nb;
    const int im = i0 + l;
    // ---- this lane's timestep: control perturbation (PI/mppi_controller.cu:130-155) ----
    const float2 e = e_next, Ui = U_next;
    if (im + 16 < T) { e_next = row[im + 16]; U_next = Ug[im + 16]; }
    float du0, du1, u0m, u1m;
    if (noise_free || im < p.opt_delay) {
      du0 = 0.0f; du1 = 0.0f; u0m = Ui.x; u1m = Ui.y;
    } else if (pure_noise) {
      du0 = __fmul_rn(e.x, p.nu0); du1 = __fmul_rn(e.y, p.nu1); u0m = du0; u1m = du1;
    } else {
      du0 = __fmul_rn(e.x, p.nu0); du1 = __fmul_rn(e.y, p.nu1);
      u0m = __fadd_rn(Ui.x, du0); u1m = __fadd_rn(Ui.y, du1);
    }
    if (mine) row[im] = make_float2(u0m, u1m);  // un-clamped write-back (:153)
    u0m = u0m < p.lo0 ? p.lo0 : (u0m > p.hi0 ? p.hi0 : u0m);  // enforceConstraints, PI/neural_net_model.cu:311-323
    u1m = u1m < p.lo1 ? p.lo1 : (u1m > p.hi1 ? p.hi1 : u1m);

    // ---- phase A: the serial recursion ----
    float r_yaw = 0.0f, r_vx = 0.0f, r_vy = 0.0f;
    bool r_roll = false;
    // the controls of a timestep are fetched (shuffles) one timestep ahead: they do not depend on the state
    float u0n = __shfl_sync(full, u0m, gb), u1n = __shfl_sync(full, u1m, gb);
    for (int ii = 0; ii < nb; ii++) {
      const float u0 = u0n, u1 = u1n;
      u0n = __shfl_sync(full, u0m, gb | ((ii + 1) & 15)); u1n = __shfl_sync(full, u1m, gb | ((ii + 1) & 15));
      if (l == ii) { r_yaw = yaw; r_vx = vx; r_vy = vy; }
      // layer 1: neurons (2l, 2l+1); two interleaved partial sums (the bias opens one of them), the controls -- which arrive
      // through shuffles -- last
      float2 t = __ffma2_rn(w1[0], bcast2(roll), b1), tb = __fmul2_rn(w1[1], bcast2(vx));
      t = __ffma2_rn(w1[2], bcast2(vy), t); tb = __ffma2_rn(w1[3], bcast2(wz), tb);
      t = __ffma2_rn(w1[4], bcast2(u0), t); tb = __ffma2_rn(w1[5], bcast2(u1), tb);
      *reinterpret_cast<float2 *>(myx + 2 * l) = recip_core2(__fadd2_rn(t, tb));
      __syncwarp();
      // layer 2: four accumulator pairs over k mod 4
      float2 a0 = b2, a1 = make_float2(0.0f, 0.0f), a2 = a1, a3 = a1;  // the bias opens the first partial sum
#pragma unroll
      for (int k4 = 0; k4 < 8; k4++) {
        const float4 hv = reinterpret_cast<const float4 *>(myx)[k4];
        a0 = __ffma2_rn(w2[4 * k4 + 0], bcast2(hv.x), a0); a1 = __ffma2_rn(w2[4 * k4 + 1], bcast2(hv.y), a1);
        a2 = __ffma2_rn(w2[4 * k4 + 2], bcast2(hv.z), a2); a3 = __ffma2_rn(w2[4 * k4 + 3], bcast2(hv.w), a3);
      }
      const float2 g = recip_core2(__fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3)));
      *reinterpret_cast<float2 *>(myx + 32 + 2 * l) = g;
      __syncwarp();
      // layer 3: output jo over this lane's quarter of k, (even, odd) k packed; xor tree over the 4 quarters
      const float4 g0 = reinterpret_cast<const float4 *>(myx + 32 + 8 * q)[0], g1 = reinterpret_cast<const float4 *>(myx + 32 + 8 * q)[1];
      float2 s2 = __ffma2_rn(w3[0], make_float2(g0.x, g0.y), b3q), s3 = __fmul2_rn(w3[1], make_float2(g0.z, g0.w));
      s2 = __ffma2_rn(w3[2], make_float2(g1.x, g1.y), s2);
      s3 = __ffma2_rn(w3[3], make_float2(g1.z, g1.w), s3);
      s2 = __fadd2_rn(s2, s3);
      // the 16 partial sums (4 outputs x 4 quarters of k) meet in shared memory: one store, four broadcast loads and two levels of
      // adds instead of a two-level xor tree plus a broadcast (three dependent shuffles)
      pbuf[l] = s2.x + s2.y;  // pbuf[4 q + jo]
      __syncwarp();
      const float4 q0 = reinterpret_cast<const float4 *>(pbuf)[0], q1 = reinterpret_cast<const float4 *>(pbuf)[1];
      const float4 q2 = reinterpret_cast<const float4 *>(pbuf)[2], q3 = reinterpret_cast<const float4 *>(pbuf)[3];
      const float o0 = (q0.x + q1.x) + (q2.x + q3.x), o1 = (q0.y + q1.y) + (q2.y + q3.y);
      const float o2 = (q0.z + q1.z) + (q2.z + q3.z), o3 = (q0.w + q1.w) + (q2.w + q3.w);
      // incrementState, PI/neural_net_model.cu:334-344 (kinematics of x, y are deferred to phase B)
      yaw = fmaf(p.negate_yaw ? -wz : wz, p.dt, yaw);
      roll = fmaf(o0, p.dt, roll); vx = fmaf(o1, p.dt, vx); vy = fmaf(o2, p.dt, vy); wz = fmaf(o3, p.dt, wz);
      if (l == ii) r_roll = fabsf(roll) >= 1.57f;  // getCrash after the update (PI/costs.cu:301-305)
    }

    // ---- phase B: lane l evaluates timestep i0 + l of its rollout ----
    float sn, cs;
    sincosf(r_yaw, &sn, &cs);
    const float d0 = fmaf(cs, r_vx, -__fmul_rn(sn, r_vy));  // kinematics, PI/neural_net_model.cu:346-355
    const float d1 = fmaf(sn, r_vx, __fmul_rn(cs, r_vy));
    float px = 0.0f, py = 0.0f;
    // sequential Euler prefix of x, y over the block (the reference's order).  Fully unrolled so the 32 shuffles are in
    // flight together; lanes beyond the end of the horizon hold r_vx = r_vy = 0, i.e. contribute exact zeros.
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (l == j) { px = xcur; py = ycur; }
      xcur = fmaf(__shfl_sync(full, d0, gb | j), p.dt, xcur);
      ycur = fmaf(__shfl_sync(full, d1, gb | j), p.dt, ycur);
    }
    const bool costed = mine && im > 0;  // step 0 is never costed (PI/mppi_controller.cu:162)
    StepCostParts cpart = {0.0f, 0.0f, 0.0f, false};
    if (costed) cpart = step_cost_parts(p.cp, p.tex, px, py, r_yaw, r_vx, r_vy, u0m, u1m, du0, du1, p.nu0, p.nu1);
    const unsigned bbits = (__ballot_sync(full, costed && cpart.boundary) >> gb) & 0xffffu;
    const unsigned rbits = (__ballot_sync(full, mine && r_roll) >> gb) & 0xffffu;
    const unsigned upto = (2u << l) - 1u;  // bits 0..l
    // the boundary flag of step i is raised before step i's crash cost, the roll flag after step i's update
    const bool crash_used = crash_in || (bbits & upto) || (rbits & (upto >> 1));
    float cost = __fadd_rn(__fadd_rn(__fadd_rn(cpart.pre, crash_used ? p.cp.crash_cost_on : 0.0f), cpart.track), cpart.stab);
    if (cost > 1e12f || isnan(cost)) cost = 1e12f;
    crash_in = crash_in || bbits || rbits;
    if (mine) scost[im] = cost;
  }
  __syncwarp();
  // ---- running mean of the step costs (PI/mppi_controller.cu:162-165) ----
  // The reference's recursion running += (c_i - running) / i is the arithmetic mean of c_1 .. c_{T-1}, rounded to float at
  // every step.  Replaying it in order was a serial chain of T - 1 float -> double -> float updates (56 cycles each: 2.8 us
  // of this kernel's 47 at T = 100); here the half-warp sums the costs in DOUBLE (fixed order: lane l takes steps 1 + l,
  // 17 + l, ..., then an xor tree) and rounds once: a few float ulps (~1e-7 relative) from the recursion's value, inside
  // the 1e-4 cost tolerance like every other difference between this kernel and the reference's arithmetic.
  double csum = 0.0;
  for (int i = 1 + l; i < T; i += 16) csum += (double)scost[i];
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) csum += __shfl_xor_sync(full, csum, o);
  const float running = T > 1 ? (float)(csum * __ldg(p.inv_step + (T - 1))) : 0.0f;
  if (l == 0) {
    p.costs[gro] = running;  // + terminalCost == 0 (PI/costs.cu:411-414)
    p.crash[gro] = (unsigned char)(crash_in ? 1 : 0);
    atomicMin(p.baseline + ctrl, float_to_ordered(running));  // min-cost baseline (host loop at :627-632)
  }
}

}  // namespace

cudaError_t launch_rollout_nn32_half(const RolloutParams &p, cudaStream_t st, bool pdl) {
  const long long total = (long long)p.B * p.n_local;  // multiple of 64
  const size_t smem = (160 + 2 * (size_t)p.T) * sizeof(float);
  const bool roomy = total / 2 <= 148LL * 12;  // every CTA resident at once with the 166-register build
  if (smem > 48 * 1024) {
    cudaError_t e = roomy ? cudaFuncSetAttribute(rollout_half_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(rollout_half_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(total / 2)); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return roomy ? cudaLaunchKernelEx(&cfg, rollout_half_kernel<12>, p) : cudaLaunchKernelEx(&cfg, rollout_half_kernel<16>, p);
}

}  // namespace mppi
