// rollout_generic.cu -- the rollout kernel for ANY NeuralNetModel<7,2,3, 6, ..., 4> layer pack (run-time widths <= 128, up
// to 16 layers, FP32): what the reference's variadic template covers (PI/neural_net_model.cuh:48-62, PI/meta_math.h:13-46,
// device forward pass PI/neural_net_model.cu:357-410) beyond the two shapes that have dedicated kernels.
//
// Shape: ONE ROLLOUT PER WARP (what BDIM_Y does in the reference: the y-threads of a block split the neurons of one
// rollout, PI/neural_net_model.cu:379-405).  Lane l owns neurons l, l+32, l+64, l+96 of every layer; activations
// ping-pong through a per-warp shared-memory buffer; the transposed weights Wt[k][j] are staged once per CTA in shared
// memory (or read through the read-only path when the network does not fit), so a warp's weight loads are contiguous in j.
// Every neuron sums its products for k ascending with FMA and adds the bias last -- the reference's order.
//
// As in rollout_half.cu only the 4-variable recursion (roll, u_x, u_y, yaw rate) is serial.  Per block of 32 timesteps
// lane l prepares timestep i0 + l (noise -- read from the buffer or drawn in place from the Philox stream --, control
// perturbation, un-clamped write-back, clamp; PI/mppi_controller.cu:130-159) and afterwards evaluates that timestep's
// running cost in parallel (positions by a sequential FMA prefix = the reference's Euler order, precise sincosf, both
// costmap texels, all cost terms, PI/costs.cu:307-409; the sticky crash flag as a prefix-OR over ballots); the running
// mean (float difference, double update, PI/mppi_controller.cu:162-165) is replayed in order at the end.
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {

constexpr int GEN_MAX_LAYERS = 16;
constexpr int GEN_MAX_WIDTH = 128;
constexpr int GEN_ACT = GEN_MAX_WIDTH + 4;  // floats per activation buffer

struct GenericNet {
  int num_layers;                 // entries of width[]
  int width[GEN_MAX_LAYERS];      // 6, hidden..., 4
  int w_off[GEN_MAX_LAYERS];      // offset of Wt_l[k][j] in the packed transposed parameters
  int b_off[GEN_MAX_LAYERS];      // offset of b_l[j]
  int nparams;
  int weights_in_smem;
};

// one layer for this lane's neurons; cur / nxt in shared memory; W points at Wt[k][j], b at the bias
__device__ __forceinline__ void generic_layer(const float *__restrict__ W, const float *__restrict__ b, int nin, int nout, bool act,
                                              const float *__restrict__ cur, float *__restrict__ nxt, int lane) {
  if (nout <= 32) {
    // one neuron per lane: k runs in four interleaved partial sums (k mod 4), as in the half-warp kernel and WarpMlp32 -- the
    // dependent FMA chain is nin / 4 long instead of nin, and eight weight loads are in flight ahead of it
    if (lane < nout) {
      const float *w = W + lane;
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
      int k = 0;
#pragma unroll 2
      for (; k + 3 < nin; k += 4) {
        const float w0 = w[(k + 0) * nout], w1 = w[(k + 1) * nout], w2 = w[(k + 2) * nout], w3 = w[(k + 3) * nout];
        a0 = fmaf(w0, cur[k + 0], a0); a1 = fmaf(w1, cur[k + 1], a1); a2 = fmaf(w2, cur[k + 2], a2); a3 = fmaf(w3, cur[k + 3], a3);
      }
      for (; k < nin; k++) a0 = fmaf(w[k * nout], cur[k], a0);
      const float t = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3)), b[lane]);
      nxt[lane] = act ? tanh_fast(t) : t;
    }
    return;
  }
  // up to four neurons per lane, four independent FMA chains (k ascending, bias last: the reference's order per neuron);
  // lanes past the end shadow neuron `lane` and store nothing
  const int j0 = lane, j1 = lane + 32, j2 = lane + 64, j3 = lane + 96;
  const bool v1 = j1 < nout, v2 = j2 < nout, v3 = j3 < nout;
  const float *w0 = W + j0, *w1 = W + (v1 ? j1 : j0), *w2 = W + (v2 ? j2 : j0), *w3 = W + (v3 ? j3 : j0);
  float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll 4
  for (int k = 0; k < nin; k++) {
    const float a = cur[k];
    const int o = k * nout;
    t0 = fmaf(w0[o], a, t0); t1 = fmaf(w1[o], a, t1); t2 = fmaf(w2[o], a, t2); t3 = fmaf(w3[o], a, t3);
  }
  t0 = __fadd_rn(t0, b[j0]);
  nxt[j0] = act ? tanh_fast(t0) : t0;
  if (v1) { t1 = __fadd_rn(t1, b[j1]); nxt[j1] = act ? tanh_fast(t1) : t1; }
  if (v2) { t2 = __fadd_rn(t2, b[j2]); nxt[j2] = act ? tanh_fast(t2) : t2; }
  if (v3) { t3 = __fadd_rn(t3, b[j3]); nxt[j3] = act ? tanh_fast(t3) : t3; }
}

// SMEM_W: the weights are staged in shared memory (the compiler then emits LDS for them instead of generic loads)
template <bool SMEM_W>
__global__ void __launch_bounds__(512, 1) rollout_generic_kernel(const __grid_constant__ RolloutParams p, const __grid_constant__ GenericNet net) {
  extern __shared__ float4 gsm4[];
  float *sm = reinterpret_cast<float *>(gsm4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int T = p.T;
  const int wfloats = SMEM_W ? ((net.nparams + 3) & ~3) : 0;
  if (SMEM_W)
    for (int i = tid; i < wfloats / 4; i += blockDim.x) gsm4[i] = reinterpret_cast<const float4 *>(p.theta_t)[i];
  const float *W = SMEM_W ? sm : p.theta_t;
  __syncthreads();
  float *act = sm + wfloats + (size_t)warp * (2 * GEN_ACT + ((T + 3) & ~3));
  float *scost = act + 2 * GEN_ACT;  // [T] step costs for the deferred running mean
  const unsigned full = 0xffffffffu;
  const long long total = (long long)p.B * p.n_local;
  const long long gro = (long long)blockIdx.x * nwarps + warp;
  if (gro >= total) return;  // whole warps only, and no CTA-wide barrier follows
  const int ctrl = (int)(gro / p.n_local);
  const int lr = (int)(gro - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  const float2 *Ug = reinterpret_cast<const float2 *>(inbox + INBOX_U);
  float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)gro * T;
  float xcur = inbox[INBOX_STATE + 0], ycur = inbox[INBOX_STATE + 1], yaw = inbox[INBOX_STATE + 2];
  float roll = inbox[INBOX_STATE + 3], vx = inbox[INBOX_STATE + 4], vy = inbox[INBOX_STATE + 5], wz = inbox[INBOX_STATE + 6];
  const int rg = p.r_begin + lr;  // the GLOBAL rollout index drives the bookkeeping (R2)
  const bool noise_free = (rg == 0), pure_noise = (rg >= p.pure_noise_from);
  const uint32_t call = p.fused_noise ? *p.call_ptr : 0u;
  bool crash_in = false;
  const int L = net.num_layers;

  for (int i0 = 0; i0 < T; i0 += 32) {
    const int nb = min(32, T - i0);
    const bool mine = lane < nb;
    const int im = i0 + lane;
    // ---- this lane's timestep: control perturbation (PI/mppi_controller.cu:130-155) ----
    float2 e = make_float2(0.0f, 0.0f), Ui = make_float2(0.0f, 0.0f);
    if (mine) {
      Ui = Ug[im];
      if (p.fused_noise) {
        const float4 z = philox_normal4((uint32_t)(im >> 1), (uint32_t)rg, call, (uint32_t)(p.b_begin + ctrl), p.seed_lo, p.seed_hi);
        e = (im & 1) ? make_float2(z.z, z.w) : make_float2(z.x, z.y);
      } else {
        e = row[im];
      }
    }
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
      float *cur = act, *nxt = act + GEN_ACT;
      // network input [roll, u_x, u_y, yaw rate, steering, throttle] (PI/neural_net_model.cu:372-377)
      const float myin = lane == 0 ? roll : lane == 1 ? vx : lane == 2 ? vy : lane == 3 ? wz : lane == 4 ? u0 : u1;
      if (lane < 6) cur[lane] = myin;
      __syncwarp();
      for (int l = 0; l + 1 < L; l++) {
        generic_layer(W + net.w_off[l], W + net.b_off[l], net.width[l], net.width[l + 1], l + 2 < L, cur, nxt, lane);
        __syncwarp();
        float *t = cur; cur = nxt; nxt = t;
      }
      const float o0 = cur[0], o1 = cur[1], o2 = cur[2], o3 = cur[3];
      __syncwarp();  // everybody has the outputs before the next step's inputs overwrite a buffer
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
    // sequential Euler prefix of x, y over the block (the reference's order); lanes beyond the end of the horizon hold
    // r_vx = r_vy = 0 and contribute exact zeros
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
  // ---- running mean of the step costs (PI/mppi_controller.cu:162-165) = their arithmetic mean: summed in double over the
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

// net_structure = {6, hidden..., 4}; the packed transposed parameters (per layer Wt[k][j] then b[j]) are p.theta_t
cudaError_t launch_rollout_generic(const RolloutParams &p, cudaStream_t st, const int *net_structure, int num_layers) {
  if (num_layers < 2 || num_layers > GEN_MAX_LAYERS) return cudaErrorInvalidValue;
  GenericNet net{};
  net.num_layers = num_layers;
  int off = 0;
  for (int l = 0; l < num_layers; l++) {
    if (net_structure[l] < 1 || net_structure[l] > GEN_MAX_WIDTH) return cudaErrorInvalidValue;
    net.width[l] = net_structure[l];
    if (l + 1 < num_layers) {
      net.w_off[l] = off; off += net_structure[l] * net_structure[l + 1];
      net.b_off[l] = off; off += net_structure[l + 1];
    }
  }
  net.nparams = off;
  const long long total = (long long)p.B * p.n_local;  // multiple of 64
  // small networks: 4 rollouts per CTA (the CTAs spread over the SMs); large ones: 16 rollouts share one copy of the weights
  const size_t wbytes = (size_t)((off + 3) & ~3) * sizeof(float);
  int nwarps = wbytes <= 16 * 1024 ? 4 : 16;
  const size_t per_warp = (size_t)(2 * GEN_ACT + ((p.T + 3) & ~3)) * sizeof(float);
  size_t smem = 0;
  for (;;) {  // long horizons: fewer warps per CTA, then weights through the read-only path
    net.weights_in_smem = (wbytes + nwarps * per_warp <= 200 * 1024) ? 1 : 0;
    smem = (net.weights_in_smem ? wbytes : 0) + nwarps * per_warp;
    if (smem <= 227 * 1024 || nwarps == 1) break;
    nwarps >>= 1;
  }
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((total + nwarps - 1) / nwarps);
  if (smem > 48 * 1024) {
    cudaError_t e = net.weights_in_smem ? cudaFuncSetAttribute(rollout_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                        : cudaFuncSetAttribute(rollout_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (net.weights_in_smem) rollout_generic_kernel<true><<<grid, nwarps * 32, smem, st>>>(p, net);
  else rollout_generic_kernel<false><<<grid, nwarps * 32, smem, st>>>(p, net);
  return cudaGetLastError();
}

}  // namespace mppi
