// rollout_lanes.cu -- latency variants of the fused rollout kernel for NeuralNetModel<7,2,3,6,32,32,4>:
// ONE ROLLOUT ACROSS L = 8 / 16 / 32 LANES, with the cost evaluation deferred and spread over the lanes.
//
// Why: at the reference's own size (1920 rollouts) one thread per rollout is 60 warps on a machine with
// 592 warp schedulers, each walking a ~2000-instruction dependent chain per timestep.  Only the
// recursion  (roll, u_x, u_y, yaw_rate, yaw)_{i+1} = F(.., u_i)  is inherently serial.  So:
//
//  phase A, every timestep (serial): perturb / clamp the control, run the MLP with lane l owning
//    hidden neurons [l*32/L, (l+1)*32/L) -- the role BDIM_Y plays in the reference
//    (PI/mppi_controller.cu:275-278) -- and Euler-step the 5 recursive state variables.  Layer 1 and 2
//    weights live in registers (L >= 16) for all T steps; activations cross lanes through a
//    double-buffered 128-byte shared-memory slot per rollout (one __syncwarp per exchange); layer 3 is
//    split as (4 outputs) x (L/4 chunks of k) with a xor-shuffle tree over the chunks.
//  phase B, every L timesteps (parallel): lane l takes timestep i0+l of the block: precise sincosf of
//    its yaw, the x / y positions by a sequential FMA prefix over the block (same order as the
//    reference's Euler steps), then the whole running cost of that step -- track lookups, speed,
//    slip, control cost (PI/costs.cu:307-393).  The sticky crash flag is a prefix-OR over ballots, and
//    the running mean (PI/mppi_controller.cu:162-165) is replayed sequentially from the L step costs.
//
// Per-timestep work drops from ~650 warp instructions (everything replicated in 8 lanes) to ~140, and
// 1920 rollouts become 1920 warps: 3.2 per scheduler, enough to overlap the dependent chains.
// Noise is read and the sampled controls written back L timesteps at a time, coalesced.
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {
constexpr int kNP = 1412;  // 6-32-32-4 packed transposed parameters (dynamics.cuh layout)
constexpr int kW1 = 0, kB1 = 192, kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;
constexpr int kBlock = 128;
constexpr int kMaxT = 512;  // nominal controls / reciprocal table staged in shared memory up to this T

template <int L>
__global__ void __launch_bounds__(kBlock) rollout_lanes_kernel(const __grid_constant__ RolloutParams p) {
  constexpr int NPL = 32 / L;        // hidden neurons per lane
  constexpr int KPL = 128 / L;       // layer-3 inputs per lane (4 outputs x L/4 chunks)
  constexpr int RPB = kBlock / L;    // rollouts per CTA
  __shared__ float4 sw4[kNP / 4];
  __shared__ float4 xch4[RPB * 4 * 8];  // per rollout: h1[2][32] and h2[2][32] floats
  __shared__ float2 sU[kMaxT];
  __shared__ double sInv[kMaxT];
  const int tid = threadIdx.x, lane = tid & 31, l = lane & (L - 1);
  const int gb = lane & ~(L - 1);    // first lane of this rollout's group
  const int T = p.T;
  for (int i = tid; i < kNP / 4; i += kBlock) sw4[i] = reinterpret_cast<const float4 *>(p.theta_t)[i];
  const long long gro = (long long)blockIdx.x * RPB + (tid / L);
  const int ctrl = (int)(gro / p.n_local);
  const int lr = (int)(gro - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  const bool staged = T <= kMaxT;
  if (staged)
    for (int i = tid; i < T; i += kBlock) { sU[i] = reinterpret_cast<const float2 *>(inbox + INBOX_U)[i]; sInv[i] = p.inv_step[i]; }
  __syncthreads();
  const float *sw = reinterpret_cast<const float *>(sw4);
  const float2 *Ug = reinterpret_cast<const float2 *>(inbox + INBOX_U);

  // ---- per-lane weight slices, resident in registers for all T steps ----
  float w1[NPL][6], b1[NPL], b2[NPL];
#pragma unroll
  for (int n = 0; n < NPL; n++) {
    const int j = NPL * l + n;
#pragma unroll
    for (int k = 0; k < 6; k++) w1[n][k] = sw[kW1 + k * 32 + j];
    b1[n] = sw[kB1 + j];
    b2[n] = sw[kB2 + j];
  }
  constexpr bool W2_IN_REGS = (NPL <= 2);
  float w2r[W2_IN_REGS ? NPL : 1][32];
  if (W2_IN_REGS) {
#pragma unroll
    for (int n = 0; n < NPL; n++)
#pragma unroll
      for (int k = 0; k < 32; k++) w2r[n][k] = sw[kW2 + k * 32 + NPL * l + n];
  }
  // layer 3: lane (jo = l & 3, chunk = l >> 2) owns output jo over k in [chunk*KPL, (chunk+1)*KPL)
  const int jo = l & 3, chunk = l >> 2;
  float w3[KPL];
#pragma unroll
  for (int kk = 0; kk < KPL; kk++) w3[kk] = sw[kW3 + (chunk * KPL + kk) * 4 + jo];
  const float b3 = sw[kB3 + jo];
  float *xbuf = reinterpret_cast<float *>(xch4) + (tid / L) * 128;

  // replicated recursive state
  float roll = inbox[INBOX_STATE + 3], vx = inbox[INBOX_STATE + 4], vy = inbox[INBOX_STATE + 5], wz = inbox[INBOX_STATE + 6];
  float yaw = inbox[INBOX_STATE + 2];
  float xcur = inbox[INBOX_STATE + 0], ycur = inbox[INBOX_STATE + 1];
  float running = 0.0f;
  bool crash_in = false;
  const int rg = p.r_begin + lr;
  const bool noise_free = (rg == 0), pure_noise = (rg >= p.pure_noise_from);
  float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)gro * T;
  const unsigned full = 0xffffffffu;
  const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << gb);

  for (int i0 = 0; i0 < T; i0 += L) {
    const int nb = min(L, T - i0);
    const bool mine = l < nb;
    const float2 e_mine = mine ? row[i0 + l] : make_float2(0.0f, 0.0f);
    // this lane's record of timestep i0 + l
    float2 wb = make_float2(0.0f, 0.0f);
    float r_yaw = 0.0f, r_vx = 0.0f, r_vy = 0.0f, r_u0 = 0.0f, r_u1 = 0.0f, r_du0 = 0.0f, r_du1 = 0.0f;
    bool r_roll = false;
    // ------------------------------ phase A: the serial recursion ------------------------------
    for (int ii = 0; ii < nb; ii++) {
      const int i = i0 + ii;
      const float ex = __shfl_sync(full, e_mine.x, gb | ii);
      const float ey = __shfl_sync(full, e_mine.y, gb | ii);
      const float2 Ui = staged ? sU[i] : Ug[i];
      float du0, du1, u0, u1;
      if (noise_free || i < p.opt_delay) {
        du0 = 0.0f; du1 = 0.0f; u0 = Ui.x; u1 = Ui.y;
      } else if (pure_noise) {
        du0 = __fmul_rn(ex, p.nu0); du1 = __fmul_rn(ey, p.nu1); u0 = du0; u1 = du1;
      } else {
        du0 = __fmul_rn(ex, p.nu0); du1 = __fmul_rn(ey, p.nu1);
        u0 = __fadd_rn(Ui.x, du0); u1 = __fadd_rn(Ui.y, du1);
      }
      const float u0raw = u0, u1raw = u1;
      u0 = u0 < p.lo0 ? p.lo0 : (u0 > p.hi0 ? p.hi0 : u0);
      u1 = u1 < p.lo1 ? p.lo1 : (u1 > p.hi1 ? p.hi1 : u1);
      if (l == ii) {
        wb = make_float2(u0raw, u1raw);  // un-clamped write-back (PI/mppi_controller.cu:153)
        r_yaw = yaw; r_vx = vx; r_vy = vy; r_u0 = u0; r_u1 = u1; r_du0 = du0; r_du1 = du1;
      }
      // layer 1
      const float a0[6] = {roll, vx, vy, wz, u0, u1};
      float h[NPL];
#pragma unroll
      for (int n = 0; n < NPL; n++) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 6; k++) t = fmaf(w1[n][k], a0[k], t);
        h[n] = tanh_fast(t + b1[n]);
      }
      float *h1buf = xbuf + (i & 1) * 32;
      if constexpr (NPL == 4) *reinterpret_cast<float4 *>(h1buf + 4 * l) = make_float4(h[0], h[1], h[2], h[3]);
      else if constexpr (NPL == 2) *reinterpret_cast<float2 *>(h1buf + 2 * l) = make_float2(h[0], h[1]);
      else h1buf[l] = h[0];
      __syncwarp();
      // layer 2: k ascending, as the reference accumulates (PI/neural_net_model.cu:388-399)
      float acc[NPL];
#pragma unroll
      for (int n = 0; n < NPL; n++) acc[n] = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < 8; k4++) {
        const float4 hv = reinterpret_cast<const float4 *>(h1buf)[k4];
        const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int k = 4 * k4 + kk;
          if constexpr (W2_IN_REGS) {
#pragma unroll
            for (int n = 0; n < NPL; n++) acc[n] = fmaf(w2r[n][k], hk[kk], acc[n]);
          } else {
            const float4 w = sw4[(kW2 + k * 32) / 4 + l];
            acc[0] = fmaf(w.x, hk[kk], acc[0]); acc[1] = fmaf(w.y, hk[kk], acc[1]);
            acc[2] = fmaf(w.z, hk[kk], acc[2]); acc[3] = fmaf(w.w, hk[kk], acc[3]);
          }
        }
      }
      float g[NPL];
#pragma unroll
      for (int n = 0; n < NPL; n++) g[n] = tanh_fast(acc[n] + b2[n]);
      float *h2buf = xbuf + 64 + (i & 1) * 32;
      if constexpr (NPL == 4) *reinterpret_cast<float4 *>(h2buf + 4 * l) = make_float4(g[0], g[1], g[2], g[3]);
      else if constexpr (NPL == 2) *reinterpret_cast<float2 *>(h2buf + 2 * l) = make_float2(g[0], g[1]);
      else h2buf[l] = g[0];
      __syncwarp();
      // layer 3: partial dot over this lane's chunk, xor tree over the chunks, gather the 4 outputs
      float part = 0.0f;
#pragma unroll
      for (int q = 0; q < KPL / 4; q++) {
        const float4 gv = reinterpret_cast<const float4 *>(h2buf + chunk * KPL)[q];
        part = fmaf(w3[4 * q + 0], gv.x, part); part = fmaf(w3[4 * q + 1], gv.y, part);
        part = fmaf(w3[4 * q + 2], gv.z, part); part = fmaf(w3[4 * q + 3], gv.w, part);
      }
#pragma unroll
      for (int m = 4; m < L; m <<= 1) part += __shfl_xor_sync(full, part, m);
      part += b3;
      const float o0 = __shfl_sync(full, part, gb | 0), o1 = __shfl_sync(full, part, gb | 1);
      const float o2 = __shfl_sync(full, part, gb | 2), o3 = __shfl_sync(full, part, gb | 3);
      // Euler step of the recursive variables (incrementState, PI/neural_net_model.cu:334-344)
      const float d2 = p.negate_yaw ? -wz : wz;
      yaw = fmaf(d2, p.dt, yaw);
      roll = fmaf(o0, p.dt, roll); vx = fmaf(o1, p.dt, vx); vy = fmaf(o2, p.dt, vy); wz = fmaf(o3, p.dt, wz);
      if (l == ii) r_roll = fabsf(roll) >= 1.57f;  // getCrash after the update (PI/costs.cu:301-305)
    }
    if (mine) row[i0 + l] = wb;
    // ------------------------------ phase B: lane l evaluates timestep i0 + l ------------------------------
    float sn, cs;
    sincosf(r_yaw, &sn, &cs);
    const float d0 = fmaf(cs, r_vx, -__fmul_rn(sn, r_vy));  // kinematics, PI/neural_net_model.cu:346-355
    const float d1 = fmaf(sn, r_vx, __fmul_rn(cs, r_vy));
    float myx = 0.0f, myy = 0.0f;
    for (int j = 0; j < nb; j++) {  // sequential Euler prefix of x, y over the block
      if (l == j) { myx = xcur; myy = ycur; }
      xcur = fmaf(__shfl_sync(full, d0, gb | j), p.dt, xcur);
      ycur = fmaf(__shfl_sync(full, d1, gb | j), p.dt, ycur);
    }
    const bool costed = mine && (i0 + l) > 0;  // step 0 is never costed (PI/mppi_controller.cu:162)
    StepCostParts cpart = {0.0f, 0.0f, 0.0f, false};
    if (costed) cpart = step_cost_parts(p.cp, p.tex, myx, myy, r_yaw, r_vx, r_vy, r_u0, r_u1, r_du0, r_du1, p.nu0, p.nu1);
    const unsigned bbits = (__ballot_sync(full, costed && cpart.boundary) & gmask) >> gb;
    const unsigned rbits = (__ballot_sync(full, mine && r_roll) & gmask) >> gb;
    const unsigned upto = (l == 31) ? 0xffffffffu : ((2u << l) - 1u);  // bits 0..l
    const bool crash_used = crash_in || (bbits & upto) || (rbits & (upto >> 1));
    float cost = __fadd_rn(__fadd_rn(__fadd_rn(cpart.pre, crash_used ? p.cp.crash_cost_on : 0.0f), cpart.track), cpart.stab);
    if (cost > 1e12f || isnan(cost)) cost = 1e12f;
    crash_in = crash_in || bbits || rbits;
    for (int j = 0; j < nb; j++) {  // replay the running mean in step order (float diff, double update)
      const float cj = __shfl_sync(full, cost, gb | j);
      const int i = i0 + j;
      if (i > 0) running = (float)((double)running + (double)__fsub_rn(cj, running) * (staged ? sInv[i] : p.inv_step[i]));
    }
  }
  if (l == 0) {
    p.costs[gro] = running;
    p.crash[gro] = (unsigned char)(crash_in ? 1 : 0);
  }
  const unsigned int wbest = __reduce_min_sync(full, float_to_ordered(running));
  if (lane == 0) atomicMin(p.baseline + ctrl, wbest);
}

}  // namespace

cudaError_t launch_rollout_nn32_lanes(const RolloutParams &p, cudaStream_t st, int lanes) {
  const long long total = (long long)p.B * p.n_local;  // multiple of 64, so every CTA is full
  switch (lanes) {
    case 32: rollout_lanes_kernel<32><<<(unsigned)(total / (kBlock / 32)), kBlock, 0, st>>>(p); break;
    case 16: rollout_lanes_kernel<16><<<(unsigned)(total / (kBlock / 16)), kBlock, 0, st>>>(p); break;
    default: rollout_lanes_kernel<8><<<(unsigned)(total / (kBlock / 8)), kBlock, 0, st>>>(p); break;
  }
  return cudaGetLastError();
}

}  // namespace mppi
