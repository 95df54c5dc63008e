// rollout_warp32.cu -- the latency kernel for NeuralNetModel<7,2,3,6,32,32,4> at SMALL rollout counts (AUTO: up to 512):
// ONE ROLLOUT PER WARP, one warp per CTA.
//
// Replaces rolloutKernel (PI/mppi_controller.cu:72-184) + computeDynamics (PI/neural_net_model.cu:357-410), like
// rollout_half.cu, whose bookkeeping, deferred cost evaluation and summation orders it shares.  A warp spends ~5 cycles per
// issued instruction in these dependent chains whatever the decomposition (profiles/ncu_1920_r02.txt), so the time of a
// timestep is the number of instructions ONE warp issues for it: the half-warp kernel issues 136 per timestep for its two
// rollouts, this kernel ~75 for its single one:
//
//  * lane l owns hidden neuron l of both hidden layers, its 6 + 32 weights in registers: layer 2 is 32 FFMA in four
//    interleaved partial sums (k mod 4) against activations read back as eight broadcast float4s;
//  * layer 3 is (4 outputs) x (8 octets of k) over the 32 lanes: 4 FFMA + a three-level xor tree;
//  * per block of 32 timesteps, lane l prepares timestep i0 + l (noise, perturbation, un-clamped write-back, clamp;
//    PI/mppi_controller.cu:130-159) and evaluates its running cost afterwards (positions by a sequential FMA prefix,
//    sincosf, costmap fetches, PI/costs.cu:307-393; sticky crash flag as a prefix-OR over ballots);
//  * the step costs go to shared memory and their mean (PI/mppi_controller.cu:162-165) is taken once at the end, in
//    double, by the 32 lanes in parallel (see rollout_half.cu).
//
// Measured (rollout kernel, 100 timesteps; profiles/exp_pipe64_r02.txt section 4): 256 rollouts 31.0 us (half-warp kernel
// 34.6), 512: 32.3 (34.7), 1024: 34.8 (34.9), 1920: 44.9 (40.3), 4096: 80 (63).  It loses once the SMs fill up because every
// activation is delivered to 32 lanes instead of 16: the eight broadcast float4 loads of layer 2 alone are 32 cycles of the
// SM's 128 B / cycle shared-memory pipe per rollout and timestep (the half-warp kernel: 16), 13 warps per SM at 1920 rollouts.
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {
constexpr int kW1 = 0, kB1 = 192, kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;  // packed transposed layout

__global__ void __launch_bounds__(32, 16) rollout_warp32_kernel(const __grid_constant__ RolloutParams p) {
  extern __shared__ float4 smem4[];
  float *xbuf = reinterpret_cast<float *>(smem4);  // h1[32], h2[32]
  const int lane = threadIdx.x;
  const int T = p.T;
  float *scost = xbuf + 64;  // [T] step costs for the deferred running mean
  const unsigned full = 0xffffffffu;
  const long long gro = blockIdx.x;  // rollout index over B * n_local
  const int ctrl = (int)(gro / p.n_local);
  const int lr = (int)(gro - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;

  // ---- lane-resident weight slices: neuron `lane` of layers 1 and 2, output (lane & 3) over k in [4 o, 4 o + 4) ----
  const float *th = p.theta_fold;  // tanh scale and affine map folded into the weights (fold_nn32): activations travel as r
  float w1[6], w2[32], w3[4];
#pragma unroll
  for (int k = 0; k < 6; k++) w1[k] = th[kW1 + k * 32 + lane];
#pragma unroll
  for (int k = 0; k < 32; k++) w2[k] = th[kW2 + k * 32 + lane];
  const float b1 = th[kB1 + lane], b2 = th[kB2 + lane];
  const int jo = lane & 3, oct = lane >> 2;
#pragma unroll
  for (int m = 0; m < 4; m++) w3[m] = th[kW3 + (4 * oct + m) * 4 + jo];
  const float b3o = oct == 0 ? th[kB3 + jo] : 0.0f;  // the output bias opens the partial sum of the first octet

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
  float2 e_next = lane < T ? row[lane] : make_float2(0.0f, 0.0f);
  float2 U_next = lane < T ? Ug[lane] : make_float2(0.0f, 0.0f);

  for (int i0 = 0; i0 < T; i0 += 32) {
    const int nb = min(32, T - i0);
    const bool mine = lane < nb;
    const int im = i0 + lane;
    // ---- this lane's timestep: control perturbation (PI/mppi_controller.cu:130-155) ----
    const float2 e = e_next, Ui = U_next;
    if (im + 32 < T) { e_next = row[im + 32]; U_next = Ug[im + 32]; }
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
    for (int ii = 0; ii < nb; ii++) {
      const float u0 = __shfl_sync(full, u0m, ii), u1 = __shfl_sync(full, u1m, ii);
      if (lane == ii) { r_yaw = yaw; r_vx = vx; r_vy = vy; }
      // layer 1: neuron `lane`; two interleaved partial sums, bias last
      float ta = fmaf(w1[0], roll, b1), tb = __fmul_rn(w1[1], vx);  // the bias opens one of the partial sums
      ta = fmaf(w1[2], vy, ta); tb = fmaf(w1[3], wz, tb);
      ta = fmaf(w1[4], u0, ta); tb = fmaf(w1[5], u1, tb);
      xbuf[lane] = recip_core(__fadd_rn(ta, tb));
      __syncwarp();
      // layer 2: four partial sums over k mod 4
      float a0 = b2, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < 8; k4++) {
        const float4 hv = reinterpret_cast<const float4 *>(xbuf)[k4];
        a0 = fmaf(w2[4 * k4 + 0], hv.x, a0); a1 = fmaf(w2[4 * k4 + 1], hv.y, a1);
        a2 = fmaf(w2[4 * k4 + 2], hv.z, a2); a3 = fmaf(w2[4 * k4 + 3], hv.w, a3);
      }
      xbuf[32 + lane] = recip_core(__fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3)));
      __syncwarp();
      // layer 3: output jo over this lane's octet of k; xor tree over the 8 octets
      const float4 gv = reinterpret_cast<const float4 *>(xbuf + 32)[oct];
      float part = __fadd_rn(fmaf(w3[2], gv.z, fmaf(w3[0], gv.x, b3o)), fmaf(w3[3], gv.w, __fmul_rn(w3[1], gv.y)));
      part = __fadd_rn(part, __shfl_xor_sync(full, part, 4));
      part = __fadd_rn(part, __shfl_xor_sync(full, part, 8));
      part = __fadd_rn(part, __shfl_xor_sync(full, part, 16));
      const float o0 = __shfl_sync(full, part, 0), o1 = __shfl_sync(full, part, 1);
      const float o2 = __shfl_sync(full, part, 2), o3 = __shfl_sync(full, part, 3);
      // incrementState, PI/neural_net_model.cu:334-344 (kinematics of x, y are deferred to phase B)
      yaw = fmaf(p.negate_yaw ? -wz : wz, p.dt, yaw);
      roll = fmaf(o0, p.dt, roll); vx = fmaf(o1, p.dt, vx); vy = fmaf(o2, p.dt, vy); wz = fmaf(o3, p.dt, wz);
      if (lane == ii) r_roll = fabsf(roll) >= 1.57f;  // getCrash after the update (PI/costs.cu:301-305)
    }

    // ---- phase B: lane l evaluates timestep i0 + l ----
    float sn, cs;
    sincosf(r_yaw, &sn, &cs);
    const float d0 = fmaf(cs, r_vx, -__fmul_rn(sn, r_vy));  // kinematics, PI/neural_net_model.cu:346-355
    const float d1 = fmaf(sn, r_vx, __fmul_rn(cs, r_vy));
    float px = 0.0f, py = 0.0f;
    // sequential Euler prefix of x, y over the block (the reference's order).  Fully unrolled so the shuffles are in flight
    // together; lanes beyond the end of the horizon hold r_vx = r_vy = 0, i.e. contribute exact zeros.
#pragma unroll
    for (int j = 0; j < 32; j++) {
      if (lane == j) { px = xcur; py = ycur; }
      xcur = fmaf(__shfl_sync(full, d0, j), p.dt, xcur);
      ycur = fmaf(__shfl_sync(full, d1, j), p.dt, ycur);
    }
    const bool costed = mine && im > 0;  // step 0 is never costed (PI/mppi_controller.cu:162)
    StepCostParts cpart = {0.0f, 0.0f, 0.0f, false};
    if (costed) cpart = step_cost_parts(p.cp, p.tex, px, py, r_yaw, r_vx, r_vy, u0m, u1m, du0, du1, p.nu0, p.nu1);
    const unsigned bbits = __ballot_sync(full, costed && cpart.boundary);
    const unsigned rbits = __ballot_sync(full, mine && r_roll);
    const unsigned upto = (2u << lane) - 1u;  // bits 0..lane
    // the boundary flag of step i is raised before step i's crash cost, the roll flag after step i's update
    const bool crash_used = crash_in || (bbits & upto) || (rbits & (upto >> 1));
    float cost = __fadd_rn(__fadd_rn(__fadd_rn(cpart.pre, crash_used ? p.cp.crash_cost_on : 0.0f), cpart.track), cpart.stab);
    if (cost > 1e12f || isnan(cost)) cost = 1e12f;
    crash_in = crash_in || bbits || rbits;
    if (mine) scost[im] = cost;
  }
  __syncwarp();
  // ---- running mean of the step costs (PI/mppi_controller.cu:162-165) = their arithmetic mean, summed in double over the
  //      32 lanes in a fixed order and rounded once (see rollout_half.cu) ----
  double csum = 0.0;
  for (int i = 1 + lane; i < T; i += 32) csum += (double)scost[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(full, csum, o);
  const float running = T > 1 ? (float)(csum * __ldg(p.inv_step + (T - 1))) : 0.0f;
  if (lane == 0) {
    p.costs[gro] = running;  // + terminalCost == 0 (PI/costs.cu:411-414)
    p.crash[gro] = (unsigned char)(crash_in ? 1 : 0);
    atomicMin(p.baseline + ctrl, float_to_ordered(running));  // min-cost baseline (host loop at :627-632)
  }
}

}  // namespace

cudaError_t launch_rollout_nn32_warp(const RolloutParams &p, cudaStream_t st, bool pdl) {
  const long long total = (long long)p.B * p.n_local;  // multiple of 64
  const size_t smem = (64 + (size_t)p.T) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rollout_warp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)total); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rollout_warp32_kernel, p);
}

}  // namespace mppi
