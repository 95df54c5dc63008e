// rollout_generic.cu -- the rollout kernel for ANY NeuralNetModel<7,2,3, 6, ..., 4> layer pack (run-time widths <= 128, up
// to 16 layers, FP32): what the reference's variadic template covers (PI/neural_net_model.cuh:48-62, PI/meta_math.h:13-46,
// device forward pass PI/neural_net_model.cu:357-410) beyond the two shapes that have dedicated kernels.
//
// Shape: ONE WARP PER R ROLLOUTS (R = 1 or 2), the lanes splitting the neurons of a layer (what BDIM_Y does in the
// reference: the y-threads of a block split the neurons of one rollout, PI/neural_net_model.cu:379-405).  Lane l owns the
// NPL = 1, 2 or 4 adjacent neurons l NPL .. l NPL + NPL - 1 of a layer of up to 32 NPL neurons; the weights are staged once per
// CTA in shared memory as Wt[k][32 NPL] (rows padded with zeros), so a lane fetches its NPL weights of input k with one
// vector load, and the activations of the R rollouts ping-pong through a per-warp buffer [k][R] that is read four inputs at
// a time as broadcast float4s.  The kernel is bound by shared-memory bandwidth (every warp streams the whole network once
// per timestep), which is why a warp carries two rollouts when there are enough of them: each weight fetched feeds both.
// A network too large for shared memory is read in place through the read-only path (unpadded rows, scalar loads).
// Every neuron sums its products in k order (in two or four interleaved partial sums when the lane has fewer than eight
// independent chains) and adds the bias last.
//
// As in rollout_half.cu only the 4-variable recursion (roll, u_x, u_y, yaw rate) is serial.  Per block of 32 timesteps
// lane l prepares timestep i0 + l (noise -- read from the buffer or drawn in place from the Philox stream --, control
// perturbation, un-clamped write-back, clamp; PI/mppi_controller.cu:130-159) and afterwards evaluates that timestep's
// running cost in parallel (positions by a sequential FMA prefix = the reference's Euler order, precise sincosf, both
// costmap texels, all cost terms, PI/costs.cu:307-409; the sticky crash flag as a prefix-OR over ballots); the running
// mean (PI/mppi_controller.cu:162-165) is the arithmetic mean of the step costs, summed in double at the end.
#include <cstdlib>
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {

constexpr int GEN_MAX_LAYERS = 16;
constexpr int GEN_MAX_WIDTH = 128;

struct GenericNet {
  int num_layers;                 // entries of width[]
  int width[GEN_MAX_LAYERS];      // 6, hidden..., 4
  int npl[GEN_MAX_LAYERS];        // neurons per lane of the layer that transition l produces: 1, 2 or 4; -4 / -8: narrow_layer<4 / 8>
  int w_off[GEN_MAX_LAYERS];      // offset of Wt_l[k][j] in the packed transposed parameters (theta_t)
  int b_off[GEN_MAX_LAYERS];      // offset of b_l[j] there
  int pw_off[GEN_MAX_LAYERS];     // the same two in the padded shared-memory copy: Wt_l[k][32 npl], b_l[32 npl]
  int pb_off[GEN_MAX_LAYERS];
  int npadded;                    // floats of the padded copy (multiple of 4)
};

template <int N>
struct Vec;
template <>
struct Vec<1> { using T = float; };
template <>
struct Vec<2> { using T = float2; };
template <>
struct Vec<4> { using T = float4; };

// One layer for this lane's NPL neurons and the warp's R rollouts.  PAD: W / b are the padded shared-memory copy.
// cur[k * R + r], nxt[j * R + r] in shared memory (16-byte aligned).
template <int NPL, int R, bool PAD>
__device__ __forceinline__ void wide_layer(const float *__restrict__ W, const float *__restrict__ b, int nin, int nout, bool act,
                                           const float *__restrict__ cur, float *__restrict__ nxt, int lane) {
  constexpr int KS = (NPL * R >= 8) ? 1 : (NPL * R >= 4) ? 2 : 4;  // interleaved partial sums per neuron: eight chains in flight where possible
  float acc[NPL][R][KS];
#pragma unroll
  for (int n = 0; n < NPL; n++)
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
      for (int s = 0; s < KS; s++) acc[n][r][s] = 0.0f;
  const int j0 = lane * NPL;
  int jc[NPL];  // unpadded rows: columns past the end shadow column 0 (their results land in unused activation slots)
#pragma unroll
  for (int n = 0; n < NPL; n++) jc[n] = (j0 + n < nout) ? j0 + n : 0;
  auto load_w = [&](int k, float (&w)[NPL]) {
    if (PAD) {
      const typename Vec<NPL>::T v = *reinterpret_cast<const typename Vec<NPL>::T *>(W + (size_t)(k * 32 + lane) * NPL);
      const float *vf = reinterpret_cast<const float *>(&v);
#pragma unroll
      for (int n = 0; n < NPL; n++) w[n] = vf[n];
    } else {
#pragma unroll
      for (int n = 0; n < NPL; n++) w[n] = __ldg(W + (size_t)k * nout + jc[n]);
    }
  };
  // groups of four inputs; lanes with few accumulators unroll further so that more loads are in flight ahead of the FMAs
  // (a warp often has its scheduler to itself here, so nothing else hides the shared-memory latency)
  constexpr int UNROLL = (NPL * R >= 8) ? 1 : (NPL * R >= 4) ? 2 : 4;
  int k = 0;
#pragma unroll UNROLL
  for (; k + 3 < nin; k += 4) {
    float a[4][R];
    if (R == 1) {
      const float4 v = *reinterpret_cast<const float4 *>(cur + k);
      a[0][0] = v.x; a[1][0] = v.y; a[2][0] = v.z; a[3][0] = v.w;
    } else {
      const float4 v0 = *reinterpret_cast<const float4 *>(cur + 2 * k), v1 = *reinterpret_cast<const float4 *>(cur + 2 * k + 4);
      a[0][0] = v0.x; a[0][R - 1] = v0.y; a[1][0] = v0.z; a[1][R - 1] = v0.w;
      a[2][0] = v1.x; a[2][R - 1] = v1.y; a[3][0] = v1.z; a[3][R - 1] = v1.w;
    }
    float w[4][NPL];
#pragma unroll
    for (int i = 0; i < 4; i++) load_w(k + i, w[i]);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int n = 0; n < NPL; n++)
#pragma unroll
        for (int r = 0; r < R; r++) acc[n][r][i % KS] = fmaf(w[i][n], a[i][r], acc[n][r][i % KS]);
  }
  for (; k < nin; k++) {
    float w[NPL];
    load_w(k, w);
#pragma unroll
    for (int n = 0; n < NPL; n++)
#pragma unroll
      for (int r = 0; r < R; r++) acc[n][r][0] = fmaf(w[n], cur[k * R + r], acc[n][r][0]);
  }
  float out[NPL * R];
#pragma unroll
  for (int n = 0; n < NPL; n++) {
    const float bias = PAD ? b[j0 + n] : (j0 + n < nout ? __ldg(b + j0 + n) : 0.0f);
#pragma unroll
    for (int r = 0; r < R; r++) {
      float t = acc[n][r][0];
      if (KS == 2) t = __fadd_rn(t, acc[n][r][KS - 1]);
      if (KS == 4) t = __fadd_rn(__fadd_rn(t, acc[n][r][1 % KS]), __fadd_rn(acc[n][r][2 % KS], acc[n][r][3 % KS]));
      t = __fadd_rn(t, bias);
      out[n * R + r] = act ? tanh_fast(t) : t;
    }
  }
  // the lane's NPL R results are contiguous in the [j][R] buffer
  float *dst = nxt + (size_t)j0 * R;
  if (NPL * R == 1) dst[0] = out[0];
  else if (NPL * R == 2) *reinterpret_cast<float2 *>(dst) = make_float2(out[0], out[1 % (NPL * R)]);
  else {
#pragma unroll
    for (int q = 0; q < NPL * R / 4; q++)
      reinterpret_cast<float4 *>(dst)[q] = make_float4(out[(4 * q) % (NPL * R)], out[(4 * q + 1) % (NPL * R)], out[(4 * q + 2) % (NPL * R)],
                                                      out[(4 * q + 3) % (NPL * R)]);
  }
}

// A layer of at most 8 neurons (the output layer): one neuron per lane would leave most of the warp idle behind an nin-long
// chain, so the lanes split the inputs as well -- lane = s NP + j sums inputs k = s (mod 32 / NP) of neuron j, and a butterfly
// over s finishes the sums.  PAD: W is [k][NP] (zero rows / columns beyond the layer), so lane l reads W[32 i + l].
template <int NP, int R, bool PAD>
__device__ __forceinline__ void narrow_layer(const float *__restrict__ W, const float *__restrict__ b, int nin, int nout, bool act,
                                             const float *__restrict__ cur, float *__restrict__ nxt, int lane) {
  constexpr int S = 32 / NP;
  const int j = lane % NP, s = lane / NP;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll 4
  for (int k = s, i = 0; k < nin; k += S, i++) {
    const float w = PAD ? W[32 * i + lane] : (j < nout ? __ldg(W + (size_t)k * nout + j) : 0.0f);
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = fmaf(w, cur[k * R + r], acc[r]);
  }
#pragma unroll
  for (int o = NP; o < 32; o <<= 1)
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = __fadd_rn(acc[r], __shfl_xor_sync(0xffffffffu, acc[r], o));
  if (lane < nout) {
    const float bias = PAD ? b[lane] : __ldg(b + lane);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const float t = __fadd_rn(acc[r], bias);
      nxt[lane * R + r] = act ? tanh_fast(t) : t;
    }
  }
}

// R: rollouts per warp.  SMEM_W: the weights are staged (padded) in shared memory.
template <int R, bool SMEM_W>
__global__ void __launch_bounds__(512, 1) rollout_generic_kernel(const __grid_constant__ RolloutParams p, const __grid_constant__ GenericNet net) {
  extern __shared__ float4 gsm4[];
  float *sm = reinterpret_cast<float *>(gsm4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int T = p.T, Tpad = (T + 3) & ~3;
  const int L = net.num_layers;
  const int wfloats = SMEM_W ? net.npadded : 0;
  if (SMEM_W) {
    for (int i = tid; i < wfloats / 4; i += blockDim.x) gsm4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();
    for (int l = 0; l + 1 < L; l++) {
      const int nin = net.width[l], nout = net.width[l + 1], row = net.npl[l] < 0 ? -net.npl[l] : 32 * net.npl[l];
      for (int i = tid; i < nin * nout; i += blockDim.x) {
        const int k = i / nout, j = i - k * nout;
        sm[net.pw_off[l] + k * row + j] = p.theta_t[net.w_off[l] + i];
      }
      for (int j = tid; j < nout; j += blockDim.x) sm[net.pb_off[l] + j] = p.theta_t[net.b_off[l] + j];
    }
  }
  __syncthreads();
  float *act = sm + wfloats + (size_t)warp * (2 * GEN_MAX_WIDTH * R + R * Tpad);
  float *scost = act + 2 * GEN_MAX_WIDTH * R;  // [R][Tpad] step costs for the deferred running mean
  const unsigned full = 0xffffffffu;
  const long long total = (long long)p.B * p.n_local;  // rollouts per controller are a multiple of 64: a warp's R rollouts share one
  const long long g0 = ((long long)blockIdx.x * nwarps + warp) * R;
  if (g0 >= total) return;  // whole warps only, and no CTA-wide barrier follows
  const int ctrl = (int)(g0 / p.n_local);
  const int lr0 = (int)(g0 - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  const float2 *Ug = reinterpret_cast<const float2 *>(inbox + INBOX_U);
  const uint32_t call = p.fused_noise ? *p.call_ptr : 0u;
  float xcur[R], ycur[R], yaw[R], roll[R], vx[R], vy[R], wz[R];
  float2 *row[R];
  int rg[R];
  bool crash_in[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    xcur[r] = inbox[INBOX_STATE + 0]; ycur[r] = inbox[INBOX_STATE + 1]; yaw[r] = inbox[INBOX_STATE + 2];
    roll[r] = inbox[INBOX_STATE + 3]; vx[r] = inbox[INBOX_STATE + 4]; vy[r] = inbox[INBOX_STATE + 5]; wz[r] = inbox[INBOX_STATE + 6];
    row[r] = reinterpret_cast<float2 *>(p.du) + (size_t)(g0 + r) * T;
    rg[r] = p.r_begin + lr0 + r;  // the GLOBAL rollout index drives the bookkeeping (R2)
    crash_in[r] = false;
  }

  for (int i0 = 0; i0 < T; i0 += 32) {
    const int nb = min(32, T - i0);
    const bool mine = lane < nb;
    const int im = i0 + lane;
    // ---- this lane's timestep: control perturbation (PI/mppi_controller.cu:130-155) ----
    float du0[R], du1[R], u0m[R], u1m[R];
    const float2 Ui = mine ? Ug[im] : make_float2(0.0f, 0.0f);
#pragma unroll
    for (int r = 0; r < R; r++) {
      float2 e = make_float2(0.0f, 0.0f);
      if (mine) {
        if (p.fused_noise) {
          const float4 z = philox_normal4((uint32_t)(im >> 1), (uint32_t)rg[r], call, (uint32_t)(p.b_begin + ctrl), p.seed_lo, p.seed_hi);
          e = (im & 1) ? make_float2(z.z, z.w) : make_float2(z.x, z.y);
        } else {
          e = row[r][im];
        }
      }
      if (rg[r] == 0 || im < p.opt_delay) {
        du0[r] = 0.0f; du1[r] = 0.0f; u0m[r] = Ui.x; u1m[r] = Ui.y;
      } else if (rg[r] >= p.pure_noise_from) {
        du0[r] = __fmul_rn(e.x, p.nu0); du1[r] = __fmul_rn(e.y, p.nu1); u0m[r] = du0[r]; u1m[r] = du1[r];
      } else {
        du0[r] = __fmul_rn(e.x, p.nu0); du1[r] = __fmul_rn(e.y, p.nu1);
        u0m[r] = __fadd_rn(Ui.x, du0[r]); u1m[r] = __fadd_rn(Ui.y, du1[r]);
      }
      if (mine) row[r][im] = make_float2(u0m[r], u1m[r]);  // un-clamped write-back (:153)
      u0m[r] = u0m[r] < p.lo0 ? p.lo0 : (u0m[r] > p.hi0 ? p.hi0 : u0m[r]);  // enforceConstraints, PI/neural_net_model.cu:311-323
      u1m[r] = u1m[r] < p.lo1 ? p.lo1 : (u1m[r] > p.hi1 ? p.hi1 : u1m[r]);
    }

    // ---- phase A: the serial recursion, R rollouts side by side ----
    float r_yaw[R], r_vx[R], r_vy[R];
    bool r_roll[R];
#pragma unroll
    for (int r = 0; r < R; r++) { r_yaw[r] = 0.0f; r_vx[r] = 0.0f; r_vy[r] = 0.0f; r_roll[r] = false; }
    for (int ii = 0; ii < nb; ii++) {
      float *cur = act, *nxt = act + GEN_MAX_WIDTH * R;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const float u0 = __shfl_sync(full, u0m[r], ii), u1 = __shfl_sync(full, u1m[r], ii);
        if (lane == ii) { r_yaw[r] = yaw[r]; r_vx[r] = vx[r]; r_vy[r] = vy[r]; }
        // network input [roll, u_x, u_y, yaw rate, steering, throttle] (PI/neural_net_model.cu:372-377)
        const float myin = lane == 0 ? roll[r] : lane == 1 ? vx[r] : lane == 2 ? vy[r] : lane == 3 ? wz[r] : lane == 4 ? u0 : u1;
        if (lane < 6) cur[lane * R + r] = myin;
      }
      __syncwarp();
      for (int l = 0; l + 1 < L; l++) {
        const float *W = SMEM_W ? sm + net.pw_off[l] : p.theta_t + net.w_off[l];
        const float *b = SMEM_W ? sm + net.pb_off[l] : p.theta_t + net.b_off[l];
        const int nin = net.width[l], nout = net.width[l + 1];
        const bool tanh_layer = l + 2 < L;
        switch (net.npl[l]) {
          case -4: narrow_layer<4, R, SMEM_W>(W, b, nin, nout, tanh_layer, cur, nxt, lane); break;
          case -8: narrow_layer<8, R, SMEM_W>(W, b, nin, nout, tanh_layer, cur, nxt, lane); break;
          case 1: wide_layer<1, R, SMEM_W>(W, b, nin, nout, tanh_layer, cur, nxt, lane); break;
          case 2: wide_layer<2, R, SMEM_W>(W, b, nin, nout, tanh_layer, cur, nxt, lane); break;
          default: wide_layer<4, R, SMEM_W>(W, b, nin, nout, tanh_layer, cur, nxt, lane); break;
        }
        __syncwarp();
        float *t = cur; cur = nxt; nxt = t;
      }
      float o[4][R];
      if (R == 1) {
        const float4 v = *reinterpret_cast<const float4 *>(cur);
        o[0][0] = v.x; o[1][0] = v.y; o[2][0] = v.z; o[3][0] = v.w;
      } else {
        const float4 v0 = *reinterpret_cast<const float4 *>(cur), v1 = *reinterpret_cast<const float4 *>(cur + 4);
        o[0][0] = v0.x; o[0][R - 1] = v0.y; o[1][0] = v0.z; o[1][R - 1] = v0.w;
        o[2][0] = v1.x; o[2][R - 1] = v1.y; o[3][0] = v1.z; o[3][R - 1] = v1.w;
      }
      __syncwarp();  // everybody has the outputs before the next step's inputs overwrite a buffer
#pragma unroll
      for (int r = 0; r < R; r++) {
        // incrementState, PI/neural_net_model.cu:334-344 (kinematics of x, y are deferred to phase B)
        yaw[r] = fmaf(p.negate_yaw ? -wz[r] : wz[r], p.dt, yaw[r]);
        roll[r] = fmaf(o[0][r], p.dt, roll[r]); vx[r] = fmaf(o[1][r], p.dt, vx[r]);
        vy[r] = fmaf(o[2][r], p.dt, vy[r]); wz[r] = fmaf(o[3][r], p.dt, wz[r]);
        if (lane == ii) r_roll[r] = fabsf(roll[r]) >= 1.57f;  // getCrash after the update (PI/costs.cu:301-305)
      }
    }

    // ---- phase B: lane l evaluates timestep i0 + l of each rollout ----
#pragma unroll
    for (int r = 0; r < R; r++) {
      float sn, cs;
      sincosf(r_yaw[r], &sn, &cs);
      const float d0 = fmaf(cs, r_vx[r], -__fmul_rn(sn, r_vy[r]));  // kinematics, PI/neural_net_model.cu:346-355
      const float d1 = fmaf(sn, r_vx[r], __fmul_rn(cs, r_vy[r]));
      float px = 0.0f, py = 0.0f;
      // sequential Euler prefix of x, y over the block (the reference's order); lanes beyond the end of the horizon hold
      // r_vx = r_vy = 0 and contribute exact zeros
#pragma unroll
      for (int j = 0; j < 32; j++) {
        if (lane == j) { px = xcur[r]; py = ycur[r]; }
        xcur[r] = fmaf(__shfl_sync(full, d0, j), p.dt, xcur[r]);
        ycur[r] = fmaf(__shfl_sync(full, d1, j), p.dt, ycur[r]);
      }
      const bool costed = mine && im > 0;  // step 0 is never costed (PI/mppi_controller.cu:162)
      StepCostParts cpart = {0.0f, 0.0f, 0.0f, false};
      if (costed) cpart = step_cost_parts(p.cp, p.tex, px, py, r_yaw[r], r_vx[r], r_vy[r], u0m[r], u1m[r], du0[r], du1[r], p.nu0, p.nu1);
      const unsigned bbits = __ballot_sync(full, costed && cpart.boundary);
      const unsigned rbits = __ballot_sync(full, mine && r_roll[r]);
      const unsigned upto = (2u << lane) - 1u;  // bits 0..lane
      // the boundary flag of step i is raised before step i's crash cost, the roll flag after step i's update
      const bool crash_used = crash_in[r] || (bbits & upto) || (rbits & (upto >> 1));
      float cost = __fadd_rn(__fadd_rn(__fadd_rn(cpart.pre, crash_used ? p.cp.crash_cost_on : 0.0f), cpart.track), cpart.stab);
      if (cost > 1e12f || isnan(cost)) cost = 1e12f;
      crash_in[r] = crash_in[r] || bbits || rbits;
      if (mine) scost[r * Tpad + im] = cost;
    }
  }
  __syncwarp();
  // ---- running mean of the step costs (PI/mppi_controller.cu:162-165) = their arithmetic mean: summed in double over the
  //      32 lanes in a fixed order and rounded once (see rollout_half.cu) ----
#pragma unroll
  for (int r = 0; r < R; r++) {
    double csum = 0.0;
    for (int i = 1 + lane; i < T; i += 32) csum += (double)scost[r * Tpad + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(full, csum, o);
    const float running = T > 1 ? (float)(csum * __ldg(p.inv_step + (T - 1))) : 0.0f;
    if (lane == 0) {
      p.costs[g0 + r] = running;  // + terminalCost == 0 (PI/costs.cu:411-414)
      p.crash[g0 + r] = (unsigned char)(crash_in[r] ? 1 : 0);
      atomicMin(p.baseline + ctrl, float_to_ordered(running));  // min-cost baseline (host loop at :627-632)
    }
  }
}

template <int R, bool SMEM_W>
cudaError_t launch_generic_t(const RolloutParams &p, const GenericNet &net, unsigned grid, int nwarps, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rollout_generic_kernel<R, SMEM_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  rollout_generic_kernel<R, SMEM_W><<<grid, nwarps * 32, smem, st>>>(p, net);
  return cudaGetLastError();
}

}  // namespace

// net_structure = {6, hidden..., 4}; the packed transposed parameters (per layer Wt[k][j] then b[j]) are p.theta_t
cudaError_t launch_rollout_generic(const RolloutParams &p, cudaStream_t st, const int *net_structure, int num_layers) {
  if (num_layers < 2 || num_layers > GEN_MAX_LAYERS) return cudaErrorInvalidValue;
  GenericNet net{};
  net.num_layers = num_layers;
  int off = 0, poff = 0;
  for (int l = 0; l < num_layers; l++) {
    if (net_structure[l] < 1 || net_structure[l] > GEN_MAX_WIDTH) return cudaErrorInvalidValue;
    net.width[l] = net_structure[l];
    if (l + 1 < num_layers) {
      const int nin = net_structure[l], nout = net_structure[l + 1];
      // narrow layers fed by at least 32 inputs split the inputs over the lanes too (below that the butterfly costs more
      // than the short chain it removes)
      net.npl[l] = (nout <= 8 && nin >= 32) ? (nout <= 4 ? -4 : -8) : nout <= 32 ? 1 : nout <= 64 ? 2 : 4;
      net.w_off[l] = off; off += nin * nout;
      net.b_off[l] = off; off += nout;
      if (net.npl[l] < 0) {
        const int np = -net.npl[l], slices = 32 / np;
        net.pw_off[l] = poff; poff += ((nin + slices - 1) / slices) * 32;  // [k][np], whole 32-float rows of the lane-indexed view
        net.pb_off[l] = poff; poff += 32;
      } else {
        net.pw_off[l] = poff; poff += nin * 32 * net.npl[l];
        net.pb_off[l] = poff; poff += 32 * net.npl[l];
      }
    }
  }
  net.npadded = (poff + 3) & ~3;
  const long long total = (long long)p.B * p.n_local;  // multiple of 64
  // two rollouts per warp once there are enough warps to occupy the machine (the kernel is shared-memory-bandwidth bound and
  // every weight fetched then feeds both); one per warp below that, where the serial chain of a timestep is what counts
  int R = total >= 148 * 8 ? 2 : 1;
  if (const char *e = getenv("MPPI_GENERIC_R")) R = atoi(e) == 1 ? 1 : 2;
  const long long warps = total / R;
  // CTAs spread over the SMs; large problems share one copy of the weights among sixteen warps
  int nwarps = (int)((warps + 147) / 148);
  nwarps = nwarps < 2 ? 2 : nwarps > 16 ? 16 : nwarps;
  const size_t wbytes = (size_t)net.npadded * sizeof(float);
  const size_t per_warp = (size_t)(2 * GEN_MAX_WIDTH * R + R * ((p.T + 3) & ~3)) * sizeof(float);
  size_t smem = 0;
  bool in_smem = false;
  for (;;) {  // long horizons: fewer warps per CTA, then weights through the read-only path
    in_smem = wbytes + nwarps * per_warp <= 200 * 1024;
    smem = (in_smem ? wbytes : 0) + nwarps * per_warp;
    if (smem <= 227 * 1024 || nwarps == 1) break;
    nwarps >>= 1;
  }
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((warps + nwarps - 1) / nwarps);
  if (R == 2) return in_smem ? launch_generic_t<2, true>(p, net, grid, nwarps, smem, st) : launch_generic_t<2, false>(p, net, grid, nwarps, smem, st);
  return in_smem ? launch_generic_t<1, true>(p, net, grid, nwarps, smem, st) : launch_generic_t<1, false>(p, net, grid, nwarps, smem, st);
}

}  // namespace mppi
