// rollout_pipe64.cu -- the latency kernel for the 64-wide network NeuralNetModel<7,2,3,6,64,64,64,64,4> (the fork's
// SRC/params/models/wider_deeper_network_08_20_2020.npz) at controller sizes (a few thousand rollouts).
//
// Replaces rolloutKernel (PI/mppi_controller.cu:72-184) + computeDynamics (PI/neural_net_model.cu:357-410).  A 64 x 64
// layer is 4096 weights: far too many for the registers of the threads of one rollout, and re-reading them from shared
// memory for every rollout and timestep is what bounds the run-time layer kernel (rollout_generic.cu).  Here the LAYERS are
// spread over the warps of a CTA instead, and the rollouts flow through them:
//
//   one LAYER WARP per hidden layer h = 2 .. NHID keeps its 64 x 64 weights in REGISTERS -- lane l owns neurons 2l, 2l+1:
//     128 weights, held as FFMA2 pairs over consecutive inputs (W[2i][n], W[2i+1][n]) -- and does nothing but that layer for
//     the two rollouts of a pair: the activations arrive as broadcast float4 loads whose halves (h[2i], h[2i+1]) are the
//     other FFMA2 operand as they are, so the even and the odd inputs of a neuron accumulate in the two halves of one
//     register pair: 128 FFMA2 and 32 loads per pair and layer;
//   two OWNER WARPS (one per rollout of a pair) own the rollouts: output layer (inputs split over the lanes + butterfly),
//     Euler update of the recursion (PI/neural_net_model.cu:334-344) on state kept in registers, first layer (6 inputs) of
//     the next timestep, and per block of 32 timesteps the controls (PI/mppi_controller.cu:130-159) and the running costs
//     (PI/costs.cu:307-409), evaluated with lane = timestep exactly as in rollout_generic.cu.
//
// A CTA carries 2 NHID rollouts as NHID pairs.  In tick tau the owners work on pair tau mod NHID and the warp of hidden layer
// h on pair (tau - h + 1) mod NHID, so every pair meets the owners, layer 2, ..., layer NHID in consecutive ticks and is back
// at the owners one tick later for its next timestep: the ring is always full, every warp works in every tick, and a tick
// ends with one CTA barrier.  Weights are read once per kernel; per timestep and rollout only 5 x 64 activations cross
// shared memory.  Measured (profiles/exp_pipe64_r02.txt): a tick is ~700 cycles for the layer warps (a lone warp issues an
// FFMA2 every ~5 cycles) and ~800 for the owners (a chain of shared-memory and shuffle round trips queued behind the layer
// warps' loads); splitting a layer over two warps (by neurons or by inputs) or keeping both rollouts of a pair in one owner
// did not shorten it.
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {

constexpr int PW = 64;  // hidden width

template <int NHID>
struct PipeGeo {
  static constexpr int NP = NHID;            // pairs in flight = warps
  static constexpr int ROLLOUTS = 2 * NP;    // rollouts per CTA
  static constexpr int WARPS = 2 + (NHID - 1);  // two owners, one per 64 x 64 layer
  static constexpr int THREADS = 32 * WARPS;
  // packed transposed parameters (per layer Wt[k][j] then b[j])
  static constexpr int TH_W1 = 0, TH_B1 = 6 * PW;
  __host__ __device__ static constexpr int th_w(int h) { return 7 * PW + (h - 1) * (PW * PW + PW); }  // hidden layer h >= 1 -> h + 1
  __host__ __device__ static constexpr int th_b(int h) { return th_w(h) + PW * PW; }
  static constexpr int TH_WL = 7 * PW + (NHID - 1) * (PW * PW + PW), TH_BL = TH_WL + PW * 4;
  // shared memory (floats)
  static constexpr int HB = PW;                                   // one activation vector
  static constexpr int OFF_H = 0;                                 // [stage NHID][pair NP][2][HB]
  static constexpr int OFF_CTL = OFF_H + NHID * NP * 2 * HB;      // [pair][2][32][4]: clamped controls, perturbation
  static constexpr int OFF_REC = OFF_CTL + NP * 2 * 32 * 4;       // [pair][2][32][4]: yaw, u_x, u_y before the step; rolled-over flag after it
  static constexpr int OFF_COST = OFF_REC + NP * 2 * 32 * 4;      // [pair][2][Tpad]
};

template <int NHID>
__global__ void __launch_bounds__(PipeGeo<NHID>::THREADS, 2) rollout_pipe64_kernel(const __grid_constant__ RolloutParams p) {
  using G = PipeGeo<NHID>;
  constexpr int NP = G::NP;
  extern __shared__ float4 psm4[];
  float *sm = reinterpret_cast<float *>(psm4);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = p.T, Tpad = (T + 3) & ~3;
  const unsigned full = 0xffffffffu;
  const long long g_cta = (long long)blockIdx.x * G::ROLLOUTS;  // rollouts per controller are a multiple of 64: one controller per CTA
  const int ctrl = (int)(g_cta / p.n_local);
  const int lr_cta = (int)(g_cta - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  const float *th = p.theta_t;

  // The two roles are separate loops (disjoint register lifetimes: 128 weight registers live only in the layer warps) that
  // meet at the same CTA barrier once per tick.
  auto tick_barrier = []() { asm volatile("bar.sync 0;" ::: "memory"); };
#ifdef PIPE_EXP  // timing experiment: work cycles per tick of each role (ticks 200 .. 207 of CTA 0), written over rollout 0's controls
  __shared__ unsigned stamp_sm[G::WARPS][8];
  unsigned t_begin = 0;
#define PIPE_T0() do { t_begin = clock(); } while (0)
#define PIPE_T1() do { if (blockIdx.x == 0 && lane == 0 && tau >= 200 && tau < 208) stamp_sm[warp][tau - 200] = clock() - t_begin; } while (0)
#else
#define PIPE_T0() do { } while (0)
#define PIPE_T1() do { } while (0)
#endif
  const int nticks = (T + 1) * NP;

  if (warp >= 2) {
    // ---- layer warp: hidden layer h -> h + 1 (h = 1 .. NHID-1), neurons 2l, 2l+1 ----
    const int h = warp - 1;
    float2 wk[2][PW / 2];  // wk[j][i] = (W[2i][2l+j], W[2i+1][2l+j])
#pragma unroll
    for (int i = 0; i < PW / 2; i++) {
      const float2 r0 = *reinterpret_cast<const float2 *>(th + G::th_w(h) + (2 * i) * PW + 2 * lane);
      const float2 r1 = *reinterpret_cast<const float2 *>(th + G::th_w(h) + (2 * i + 1) * PW + 2 * lane);
      wk[0][i] = make_float2(r0.x, r1.x);
      wk[1][i] = make_float2(r0.y, r1.y);
    }
    const float2 bias2 = *reinterpret_cast<const float2 *>(th + G::th_b(h) + 2 * lane);
    tick_barrier();
    for (int tau = 0; tau < nticks; tau++) {
      PIPE_T0();
      const int rel = tau - h;
      const int v = rel >= 0 ? rel / NP : -1, q = rel - v * NP;  // this layer's visit number and pair
      if (v >= 0 && v < T) {
        const float *in0 = sm + G::OFF_H + ((h - 1) * NP + q) * 2 * G::HB, *in1 = in0 + G::HB;
        float *out0 = sm + G::OFF_H + (h * NP + q) * 2 * G::HB;
        // per rollout and neuron two FFMA2 chains (input pairs i even / odd), each summing even inputs in .x and odd ones in .y
        float2 a[2][2], b[2][2];
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int c = 0; c < 2; c++) { a[j][c] = make_float2(0.0f, 0.0f); b[j][c] = make_float2(0.0f, 0.0f); }
#pragma unroll
        for (int kq = 0; kq < PW / 4; kq++) {
          const float4 x = *reinterpret_cast<const float4 *>(in0 + 4 * kq);
          const float4 y = *reinterpret_cast<const float4 *>(in1 + 4 * kq);
#pragma unroll
          for (int j = 0; j < 2; j++) {
            a[j][0] = __ffma2_rn(wk[j][2 * kq], make_float2(x.x, x.y), a[j][0]);
            b[j][0] = __ffma2_rn(wk[j][2 * kq], make_float2(y.x, y.y), b[j][0]);
            a[j][1] = __ffma2_rn(wk[j][2 * kq + 1], make_float2(x.z, x.w), a[j][1]);
            b[j][1] = __ffma2_rn(wk[j][2 * kq + 1], make_float2(y.z, y.w), b[j][1]);
          }
        }
        const float2 p0 = __fadd2_rn(a[0][0], a[0][1]), p1 = __fadd2_rn(a[1][0], a[1][1]);
        const float2 q0 = __fadd2_rn(b[0][0], b[0][1]), q1 = __fadd2_rn(b[1][0], b[1][1]);
        const float2 ta = tanh_fast2(__fadd2_rn(make_float2(__fadd_rn(p0.x, p0.y), __fadd_rn(p1.x, p1.y)), bias2));
        const float2 tb = tanh_fast2(__fadd2_rn(make_float2(__fadd_rn(q0.x, q0.y), __fadd_rn(q1.x, q1.y)), bias2));
        reinterpret_cast<float2 *>(out0)[lane] = ta;
        reinterpret_cast<float2 *>(out0 + G::HB)[lane] = tb;
      }
      PIPE_T1();
      tick_barrier();
    }
    return;
  }

  // ---- owner warp e = 0 / 1: rollout e of every pair.  First layer (neurons 2l, 2l+1) and output layer (neuron j = lane % 4,
  //      inputs k = lane / 4 + 8 i) in registers ----
  const int e = warp;
  float2 w1[6];
  float wo[8];
#pragma unroll
  for (int k = 0; k < 6; k++) w1[k] = *reinterpret_cast<const float2 *>(th + G::TH_W1 + k * PW + 2 * lane);
  const float2 bias2 = *reinterpret_cast<const float2 *>(th + G::TH_B1 + 2 * lane);
#pragma unroll
  for (int i = 0; i < 8; i++) wo[i] = th[G::TH_WL + ((lane >> 2) + 8 * i) * 4 + (lane & 3)];
  const float bo = th[G::TH_BL + (lane & 3)];
  const uint32_t call = p.fused_noise ? *p.call_ptr : 0u;
  // The states of the warp's rollouts live in its registers, replicated in every lane (every lane performs the same
  // update); `q` is a compile-time index because the tick loop is unrolled over the pairs.
  float roll[NP], vx[NP], vy[NP], wz[NP], yaw[NP], xc[NP], yc[NP];
  bool crashed[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) {
    xc[q] = inbox[INBOX_STATE + 0]; yc[q] = inbox[INBOX_STATE + 1]; yaw[q] = inbox[INBOX_STATE + 2];
    roll[q] = inbox[INBOX_STATE + 3]; vx[q] = inbox[INBOX_STATE + 4]; vy[q] = inbox[INBOX_STATE + 5];
    wz[q] = inbox[INBOX_STATE + 6]; crashed[q] = false;
  }
  tick_barrier();

  for (int v = 0; v <= T; v++) {
#pragma unroll
    for (int q = 0; q < NP; q++) {
#ifdef PIPE_EXP
      const int tau = v * NP + q;
#endif
      PIPE_T0();
      // ---- visit v of pair q: finish timestep v - 1 (output layer, update), start timestep v (first layer) ----
      const int r = q * 2 + e;  // rollout within the CTA
      float *ctl = sm + G::OFF_CTL + r * 32 * 4;
      float *rec = sm + G::OFF_REC + r * 32 * 4;
      if (v >= 1) {
        const int ip = v - 1;
        // output layer: lane = 4 s + j sums inputs k = s (mod 8) of output j (PI/neural_net_model.cu:379-405)
        const float *in = sm + G::OFF_H + ((NHID - 1) * NP + q) * 2 * G::HB + e * G::HB;
        float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          acc0 = fmaf(wo[i], in[(lane >> 2) + 8 * i], acc0);
          acc1 = fmaf(wo[i + 1], in[(lane >> 2) + 8 * (i + 1)], acc1);
          acc2 = fmaf(wo[i + 2], in[(lane >> 2) + 8 * (i + 2)], acc2);
          acc3 = fmaf(wo[i + 3], in[(lane >> 2) + 8 * (i + 3)], acc3);
        }
        float o = __fadd_rn(__fadd_rn(acc0, acc1), __fadd_rn(acc2, acc3));
#pragma unroll
        for (int d = 4; d < 32; d <<= 1) o = __fadd_rn(o, __shfl_xor_sync(full, o, d));
        o = __fadd_rn(o, bo);  // lanes j (mod 4): state derivative of roll, u_x, u_y, yaw rate
        const float d0 = __shfl_sync(full, o, 0), d1 = __shfl_sync(full, o, 1), d2 = __shfl_sync(full, o, 2), d3 = __shfl_sync(full, o, 3);
        // incrementState, PI/neural_net_model.cu:334-344 (the kinematics of x, y are evaluated with the costs)
        yaw[q] = fmaf(p.negate_yaw ? -wz[q] : wz[q], p.dt, yaw[q]);
        roll[q] = fmaf(d0, p.dt, roll[q]); vx[q] = fmaf(d1, p.dt, vx[q]);
        vy[q] = fmaf(d2, p.dt, vy[q]); wz[q] = fmaf(d3, p.dt, wz[q]);
        if (lane == 0) rec[(ip & 31) * 4 + 3] = fabsf(roll[q]) >= 1.57f ? 1.0f : 0.0f;  // getCrash after the update (PI/costs.cu:301-305)
        if ((ip & 31) == 31 || ip == T - 1) {
          // ---- the block of timesteps i0 .. ip is complete: lane l evaluates timestep i0 + l ----
          __syncwarp();
          const int i0 = ip & ~31, nb = ip - i0 + 1, im = i0 + lane;
          const bool mine = lane < nb;
          float *scost = sm + G::OFF_COST + (size_t)r * Tpad;
          const float4 rc = mine ? reinterpret_cast<const float4 *>(rec)[lane] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          const float4 cu = mine ? reinterpret_cast<const float4 *>(ctl)[lane] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          float sn, cs;
          sincosf(rc.x, &sn, &cs);
          const float k0 = fmaf(cs, rc.y, -__fmul_rn(sn, rc.z));  // kinematics, PI/neural_net_model.cu:346-355
          const float k1 = fmaf(sn, rc.y, __fmul_rn(cs, rc.z));
          float px = 0.0f, py = 0.0f;
          // sequential Euler prefix of x, y over the block (the reference's order); lanes past the end contribute exact zeros
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (lane == j) { px = xc[q]; py = yc[q]; }
            xc[q] = fmaf(__shfl_sync(full, k0, j), p.dt, xc[q]);
            yc[q] = fmaf(__shfl_sync(full, k1, j), p.dt, yc[q]);
          }
          const bool costed = mine && im > 0;  // step 0 is never costed (PI/mppi_controller.cu:162)
          StepCostParts cpart = {0.0f, 0.0f, 0.0f, false};
          if (costed) cpart = step_cost_parts(p.cp, p.tex, px, py, rc.x, rc.y, rc.z, cu.x, cu.y, cu.z, cu.w, p.nu0, p.nu1);
          const unsigned bbits = __ballot_sync(full, costed && cpart.boundary);
          const unsigned rbits = __ballot_sync(full, mine && rc.w != 0.0f);
          const unsigned upto = (2u << lane) - 1u;  // bits 0..lane
          // the boundary flag of step i is raised before step i's crash cost, the roll flag after step i's update
          const bool crash_used = crashed[q] || (bbits & upto) || (rbits & (upto >> 1));
          float cost = __fadd_rn(__fadd_rn(__fadd_rn(cpart.pre, crash_used ? p.cp.crash_cost_on : 0.0f), cpart.track), cpart.stab);
          if (cost > 1e12f || isnan(cost)) cost = 1e12f;
          if (mine) scost[im] = cost;
          crashed[q] = crashed[q] || bbits || rbits;
          __syncwarp();
        }
      }
      if (v < T) {
        if ((v & 31) == 0) {
          // ---- controls of timesteps v .. v + 31: lane l prepares timestep v + l (PI/mppi_controller.cu:130-155) ----
          const int im = v + lane;
          const bool mine = im < T;
          const float2 Ui = mine ? reinterpret_cast<const float2 *>(inbox + INBOX_U)[im] : make_float2(0.0f, 0.0f);
          const int rg = p.r_begin + lr_cta + r;  // the GLOBAL rollout index drives the bookkeeping (R2)
          float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)(g_cta + r) * T;
          float2 ez = make_float2(0.0f, 0.0f);
          if (mine) {
            if (p.fused_noise) {
              const float4 z = philox_normal4((uint32_t)(im >> 1), (uint32_t)rg, call, (uint32_t)(p.b_begin + ctrl), p.seed_lo, p.seed_hi);
              ez = (im & 1) ? make_float2(z.z, z.w) : make_float2(z.x, z.y);
            } else {
              ez = row[im];
            }
          }
          float du0, du1, u0, u1;
          if (rg == 0 || im < p.opt_delay) {
            du0 = 0.0f; du1 = 0.0f; u0 = Ui.x; u1 = Ui.y;
          } else if (rg >= p.pure_noise_from) {
            du0 = __fmul_rn(ez.x, p.nu0); du1 = __fmul_rn(ez.y, p.nu1); u0 = du0; u1 = du1;
          } else {
            du0 = __fmul_rn(ez.x, p.nu0); du1 = __fmul_rn(ez.y, p.nu1);
            u0 = __fadd_rn(Ui.x, du0); u1 = __fadd_rn(Ui.y, du1);
          }
          if (mine) row[im] = make_float2(u0, u1);  // un-clamped write-back (:153)
          u0 = u0 < p.lo0 ? p.lo0 : (u0 > p.hi0 ? p.hi0 : u0);  // enforceConstraints, PI/neural_net_model.cu:311-323
          u1 = u1 < p.lo1 ? p.lo1 : (u1 > p.hi1 ? p.hi1 : u1);
          reinterpret_cast<float4 *>(ctl)[lane] = make_float4(u0, u1, du0, du1);
          __syncwarp();
        }
        // first layer of timestep v: input [roll, u_x, u_y, yaw rate, steering, throttle] (PI/neural_net_model.cu:372-377)
        const float2 u = *reinterpret_cast<const float2 *>(ctl + (v & 31) * 4);
        if (lane == 0)  // the state this step's running cost is evaluated on (the roll flag follows after the update)
          *reinterpret_cast<float2 *>(rec + (v & 31) * 4) = make_float2(yaw[q], vx[q]);
        if (lane == 1) rec[(v & 31) * 4 + 2] = vy[q];
        // k ascending in two interleaved partial sums, bias last
        float2 a = __ffma2_rn(w1[0], make_float2(roll[q], roll[q]), make_float2(0.0f, 0.0f));
        float2 b = __ffma2_rn(w1[1], make_float2(vx[q], vx[q]), make_float2(0.0f, 0.0f));
        a = __ffma2_rn(w1[2], make_float2(vy[q], vy[q]), a);
        b = __ffma2_rn(w1[3], make_float2(wz[q], wz[q]), b);
        a = __ffma2_rn(w1[4], make_float2(u.x, u.x), a);
        b = __ffma2_rn(w1[5], make_float2(u.y, u.y), b);
        const float2 t = tanh_fast2(__fadd2_rn(__fadd2_rn(a, b), bias2));
        float *out = sm + G::OFF_H + (0 * NP + q) * 2 * G::HB + e * G::HB;
        reinterpret_cast<float2 *>(out)[lane] = t;
      }
      PIPE_T1();
      tick_barrier();
    }
  }
#ifdef PIPE_EXP
  if (blockIdx.x == 0 && lane < 8) {
    float2 *row0 = reinterpret_cast<float2 *>(p.du);
    for (int w = 0; w < G::WARPS; w++) row0[w * 8 + lane] = make_float2((float)stamp_sm[w][lane], (float)w);
  }
#endif

  // ---- running mean of the step costs (PI/mppi_controller.cu:162-165) = their arithmetic mean, summed in double over the
  //      32 lanes in a fixed order and rounded once (see rollout_half.cu); min-cost baseline (host loop at :627-632) ----
  __syncwarp();
#pragma unroll
  for (int q = 0; q < NP; q++) {
    const int r = q * 2 + e;
    const float *scost = sm + G::OFF_COST + (size_t)r * Tpad;
    double csum = 0.0;
    for (int i = 1 + lane; i < T; i += 32) csum += (double)scost[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(full, csum, o);
    const float running = T > 1 ? (float)(csum * __ldg(p.inv_step + (T - 1))) : 0.0f;
    if (lane == 0) {
      p.costs[g_cta + r] = running;  // + terminalCost == 0 (PI/costs.cu:411-414)
      p.crash[g_cta + r] = (unsigned char)(crashed[q] ? 1 : 0);
      atomicMin(p.baseline + ctrl, float_to_ordered(running));
    }
  }
}

}  // namespace

// Shared memory per CTA grows with the horizon (the step costs of its rollouts); beyond the limit the caller uses another kernel.
bool rollout_pipe64_fits(int T) {
  using G = PipeGeo<4>;
  return (size_t)(G::OFF_COST + G::ROLLOUTS * ((T + 3) & ~3)) * sizeof(float) <= 100 * 1024;
}

cudaError_t launch_rollout_nn64_pipe(const RolloutParams &p, cudaStream_t st) {
  using G = PipeGeo<4>;
  const long long total = (long long)p.B * p.n_local;  // multiple of 64
  const size_t smem = (size_t)(G::OFF_COST + G::ROLLOUTS * ((p.T + 3) & ~3)) * sizeof(float);
  if (smem > 100 * 1024) return cudaErrorInvalidValue;
  static int opted_in_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && !opted_in_dev[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(rollout_pipe64_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    opted_in_dev[dev & 63] = 1;
  }
  rollout_pipe64_kernel<4><<<(unsigned)(total / G::ROLLOUTS), G::THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mppi
