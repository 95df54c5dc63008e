// weighting.cuh -- the Philox control-noise sampler, the importance-weighting reduction and the
// finalisation (control update, Savitzky-Golay smoothing, nominal trajectory).
//
// Replaces curandGenerateNormal (PI/mppi_controller.cu:612), the host min / normaliser loops
// (:627-652), normExpKernel (:193-203), weightedReductionKernel (:219-267), savitskyGolay
// (:468-499) and computeNominalTraj (:501-519).
#pragma once
#include "device_common.cuh"
#include "dynamics.cuh"
#include "philox.cuh"
#include "warp_mlp.cuh"

namespace mppi {

// ------------------------------------------------------------------------------ sampler ----
// Philox4x32-10 + Box-Muller live in philox.cuh (shared with the rollout kernels that generate their noise in place).
// One thread writes one aligned float4 (2 timesteps x 2 controls), a warp writes 512 contiguous bytes.
// Division of an index below 2^31 by a run-time constant as one multiply-high and a shift (the two divisions by Q and
// n_local were a quarter of the sampler's instructions): with s = ceil(log2 d), M = ceil(2^(31+s) / d) < 2^32 and
// n (M d - 2^(31+s)) < 2^(31+s) for every n < 2^31, so (n M) >> (31 + s) is exact.
struct FastDiv {
  uint32_t mul, shift;  // mul == 0: divisor 1
  __host__ static FastDiv make(uint32_t d) {
    FastDiv f{0u, 0u};
    if (d <= 1u) return f;
    uint32_t s = 0;
    while ((1ull << s) < d) s++;
    f.mul = (uint32_t)(((1ull << (31 + s)) + d - 1) / d);
    f.shift = s - 1;  // after the implicit >> 32 of the multiply-high
    return f;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return mul ? (__umulhi(n, mul) >> shift) : n; }
};

// `call_ptr` is the device-resident compute-call counter (advanced by finalize_kernel), so a captured
// CUDA graph replays with a fresh Philox offset every launch.
//
// Side job (zero-copy inbox): when inbox_src != nullptr, CTA 0 copies the call's inputs (state, history, U; 848 B at
// T = 100) from mapped pinned host memory into the device inbox, which replaces a separate H2D copy node in front of
// the pipeline; the rollout kernel reads the inbox only after this grid has completed.
__global__ void __launch_bounds__(256) sample_noise_kernel(float *__restrict__ du, int n_local, int r_begin, int T,
                                                            int B, int b_begin, uint32_t seed_lo, uint32_t seed_hi,
                                                            const uint32_t *__restrict__ call_ptr,
                                                            const float4 *__restrict__ inbox_src, float4 *__restrict__ inbox_dst,
                                                            int inbox_float4s, FastDiv divQ, FastDiv divN) {
  pdl_trigger();  // the rollout kernel may start fetching its weights now
  if (inbox_src != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < inbox_float4s; i += blockDim.x) inbox_dst[i] = inbox_src[i];
  const uint32_t call = *call_ptr;
  const int Q = (T + 1) >> 1;
  const long long total = (long long)B * n_local * Q;
  // 32-bit index arithmetic whenever it fits (it does up to 2^31 float4 = 32 GiB of noise): the 64-bit division and
  // modulo by run-time values were most of this kernel's instructions (166 per float4 at 1M rollouts)
  if (total < (1LL << 31)) {
    const unsigned utotal = (unsigned)total, uQ = (unsigned)Q, un = (unsigned)n_local, stride = gridDim.x * blockDim.x;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < utotal; idx += stride) {
      const unsigned g = divQ.div(idx), q = idx - g * uQ;
      const unsigned b = (B == 1) ? 0u : divN.div(g), lr = g - b * un;
      const float4 z = philox_normal4(q, (uint32_t)r_begin + lr, call, (uint32_t)b_begin + b, seed_lo, seed_hi);
      float *dst = du + ((size_t)g * T + 2 * q) * C_DIM;
      if ((T & 1) == 0) {
        *reinterpret_cast<float4 *>(dst) = z;
      } else {
        dst[0] = z.x; dst[1] = z.y;
        if (2 * q + 1 < (unsigned)T) { dst[2] = z.z; dst[3] = z.w; }
      }
    }
    return;
  }
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % Q);
    const long long g = idx / Q;
    const int lr = (int)(g % n_local), b = (int)(g / n_local);
    const float4 z = philox_normal4((uint32_t)q, (uint32_t)(r_begin + lr), call, (uint32_t)(b_begin + b), seed_lo, seed_hi);
    float *dst = du + ((size_t)g * T + 2 * q) * C_DIM;
    if ((T & 1) == 0) {
      *reinterpret_cast<float4 *>(dst) = z;
    } else {
      dst[0] = z.x; dst[1] = z.y;
      if (2 * q + 1 < T) { dst[2] = z.z; dst[3] = z.w; }
    }
  }
}

// -------------------------------------------------------------------- weighting reduction ----
// Shard partial record per controller: [0] baseline b, [1] Z = sum exp(-gamma (c - b)),
// [2] Q = sum exp(..)^2, [3] unused, [4 .. 4+2T) W[t][j] = sum exp(..) * V[r][t][j].
constexpr int SHARD_HDR = 4;
constexpr int COMBINE_GROUP = 32;  // CTAs whose partial records one group combiner adds up (weight_reduce_kernel)

struct WeightParams {
  const float *costs;            // [B][n_local]
  const float2 *V;               // [B][n_local][T]
  const unsigned int *baseline;  // [B]
  float *block_partials;         // [B][nblk][shard_floats]
  float *group_partials;         // [B][ngroups][shard_floats]: the per-CTA records of COMBINE_GROUP consecutive CTAs, summed
  float *shard;                  // [B][shard_floats]
  unsigned int *done_counter;    // [B][1 + ngroups]: tickets of the group combiners, then of the CTAs of every group
  int ngroups;
  int n_local, T, nblk, rows_per_blk, shard_floats;
  float gamma;
  int partials_only;  // 1: stop after the per-CTA partial records; finalize_kernel adds them up itself (combine_partials)
  // Peer-memory exchange (multi-GPU, mppi_p2p_init): the CTA that finishes a controller's shard record also stores it
  // into the mailbox of every GPU -- over NVLink for the peers -- and then raises that GPU's flag for (rank, controller)
  // to `seq`.  peer_mailbox[g] = GPU g's mailbox [2][G][B][shard_floats], peer_flags[g] = its flags [2][G][B].
  float *const *peer_mailbox;
  unsigned int *const *peer_flags;
  int G, rank, B;
  unsigned int seq;  // call sequence number (> 0); parity seq & 1 selects the mailbox half
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long P2P_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Entry k of the per-CTA partial records summed over the CTAs in a fixed order (bitwise reproducible run to run), with 16
// independent L2 loads in flight per thread (the partials were just written by other SMs).
__device__ __forceinline__ float sum_partials_fixed_order(const float *parts, int nblk, int shard_floats, int k) {
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
  int j = 0;
  for (; j + 15 < nblk; j += 16) {
    float v[16];
#pragma unroll
    for (int m = 0; m < 16; m++) v[m] = __ldcg(parts + (size_t)(j + m) * shard_floats + k);
#pragma unroll
    for (int m = 0; m < 16; m += 4) { a0 += v[m]; a1 += v[m + 1]; a2 += v[m + 2]; a3 += v[m + 3]; }
  }
  for (; j + 3 < nblk; j += 4) {
    a0 += __ldcg(parts + (size_t)j * shard_floats + k);
    a1 += __ldcg(parts + (size_t)(j + 1) * shard_floats + k);
    a2 += __ldcg(parts + (size_t)(j + 2) * shard_floats + k);
    a3 += __ldcg(parts + (size_t)(j + 3) * shard_floats + k);
  }
  for (; j < nblk; j++) a0 += __ldcg(parts + (size_t)j * shard_floats + k);
  return (a0 + a1) + (a2 + a3);
}

// ROWS_IN_FLIGHT independent row loads per thread: 8 keeps the kernel at 32 registers (8 CTAs per SM, one wave of 148 x 8
// CTAs for the filled GPU); the small grids of the latency configurations use 16 (one L2 round trip per 32-row CTA).
template <int ROWS_IN_FLIGHT, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) weight_reduce_kernel(const __grid_constant__ WeightParams p) {
  extern __shared__ float sm[];
  float *w = sm;                               // [rows_per_blk]
  float2 *colsum = reinterpret_cast<float2 *>(sm + ((p.rows_per_blk + 3) & ~3));  // [nrl][T]
  __shared__ float red[2][8];
  __shared__ bool is_last;
  const int b = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();  // rollout costs, sampled controls and the baseline come from the rollout kernel
  const int r0 = blk * p.rows_per_blk;
  const int nrows = min(p.rows_per_blk, p.n_local - r0);
  const float base = ordered_to_float(p.baseline[b]);
  const float *costs = p.costs + (size_t)b * p.n_local + r0;
  float z = 0.0f, q = 0.0f;
  for (int i = tid; i < nrows; i += 256) {
    const float wi = expf(-p.gamma * (costs[i] - base));  // normExpKernel (:200-201)
    w[i] = wi;
    z += wi;
    q = fmaf(wi, wi, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    z += __shfl_xor_sync(0xffffffffu, z, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = z; red[1][tid >> 5] = q; }
  __syncthreads();
  // weighted column sums: thread (rl, c) walks rows rl, rl+nrl, ... of column c
  const int T = p.T;
  const int nrl = max(1, 256 / T);
  const float2 *V = p.V + ((size_t)b * p.n_local + r0) * T;
  float *out = p.block_partials + ((size_t)b * p.nblk + blk) * p.shard_floats;
  for (int c0 = 0; c0 < T; c0 += 256) {  // one pass when T <= 256
    const int rl = (T <= 256) ? tid / T : 0;
    const int c = (T <= 256) ? tid - rl * T : c0 + tid;
    float2 acc = make_float2(0.0f, 0.0f);
    if (rl < nrl && c < T) {
      int i = rl;
      for (; i + (ROWS_IN_FLIGHT - 1) * nrl < nrows; i += ROWS_IN_FLIGHT * nrl) {  // independent row loads in flight, accumulated in row order
        float2 v[ROWS_IN_FLIGHT];
#pragma unroll
        for (int m = 0; m < ROWS_IN_FLIGHT; m++) v[m] = V[(size_t)(i + m * nrl) * T + c];
#pragma unroll
        for (int m = 0; m < ROWS_IN_FLIGHT; m++) { acc.x = fmaf(w[i + m * nrl], v[m].x, acc.x); acc.y = fmaf(w[i + m * nrl], v[m].y, acc.y); }
      }
      if (i < nrows) {  // the remaining rows as one predicated batch (a serial tail would be one L2 round trip per row)
        float2 v[ROWS_IN_FLIGHT];
#pragma unroll
        for (int m = 0; m < ROWS_IN_FLIGHT; m++) v[m] = (i + m * nrl < nrows) ? V[(size_t)(i + m * nrl) * T + c] : make_float2(0.0f, 0.0f);
#pragma unroll
        for (int m = 0; m < ROWS_IN_FLIGHT; m++)
          if (i + m * nrl < nrows) { acc.x = fmaf(w[i + m * nrl], v[m].x, acc.x); acc.y = fmaf(w[i + m * nrl], v[m].y, acc.y); }
      }
      colsum[rl * T + c] = acc;
    }
    if (T > 256) {
      if (c < T) reinterpret_cast<float2 *>(p.block_partials + ((size_t)b * p.nblk + blk) * p.shard_floats + SHARD_HDR)[c] = acc;
    }
  }
  __syncthreads();
  if (T <= 256 && tid < T) {
    float2 acc = colsum[tid];
    for (int rl = 1; rl < nrl; rl++) { acc.x += colsum[rl * T + tid].x; acc.y += colsum[rl * T + tid].y; }
    reinterpret_cast<float2 *>(out + SHARD_HDR)[tid] = acc;
  }
  if (tid == 0) {
    float zz = 0.0f, qq = 0.0f;
    for (int i = 0; i < 8; i++) { zz += red[0][i]; qq += red[1][i]; }
    out[0] = base; out[1] = zz; out[2] = qq; out[3] = 0.0f;
  }
  if (p.partials_only) return;
  // ---- two-level combine in fixed order: the last CTA of every group of COMBINE_GROUP CTAs adds the group's records up, the
  //      last of those adds the group records up.  (One last CTA walking all 1184 records of a filled GPU was a chain of 74
  //      dependent L2 round trips per thread, ~30 us of this kernel's 178 us at 1 M rollouts.) ----
  const int grp = blk / COMBINE_GROUP, g0 = grp * COMBINE_GROUP, gsize = min(COMBINE_GROUP, p.nblk - g0);
  unsigned int *tickets = p.done_counter + (size_t)b * (1 + p.ngroups);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(tickets + 1 + grp, 1u);
    is_last = (ticket == (unsigned int)gsize - 1);
    if (is_last) tickets[1 + grp] = 0;  // re-arm for the next launch
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float *shard = p.shard + (size_t)b * p.shard_floats;
  {
    const float *parts = p.block_partials + ((size_t)b * p.nblk + g0) * p.shard_floats;
    float *dst = p.ngroups == 1 ? shard : p.group_partials + ((size_t)b * p.ngroups + grp) * p.shard_floats;
    for (int k = tid; k < p.shard_floats; k += 256) {
      if (k == 0) { dst[0] = base; continue; }
      if (k == 3 || k >= SHARD_HDR + 2 * T) { dst[k] = 0.0f; continue; }
      dst[k] = sum_partials_fixed_order(parts, gsize, p.shard_floats, k);
    }
  }
  if (p.ngroups > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int ticket = atomicAdd(tickets, 1u);
      is_last = (ticket == (unsigned int)p.ngroups - 1);
      if (is_last) tickets[0] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const float *parts = p.group_partials + (size_t)b * p.ngroups * p.shard_floats;
    for (int k = tid; k < p.shard_floats; k += 256) {
      if (k == 0) { shard[0] = base; continue; }
      if (k == 3 || k >= SHARD_HDR + 2 * T) { shard[k] = 0.0f; continue; }
      shard[k] = sum_partials_fixed_order(parts, p.ngroups, p.shard_floats, k);
    }
  }
  if (p.G > 1) {
    // fused exchange: this CTA's record goes straight into every GPU's mailbox (peer stores over NVLink), then the flags
    __syncthreads();
    const int par = (int)(p.seq & 1u);
    const size_t slot = (((size_t)par * p.G + p.rank) * p.B + b) * p.shard_floats;
    for (int g = 0; g < p.G; g++) {
      float *dst = p.peer_mailbox[g] + slot;
      for (int k = tid; k < p.shard_floats; k += 256) dst[k] = shard[k];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.G) st_release_sys(p.peer_flags[tid] + ((size_t)par * p.G + p.rank) * p.B + b, p.seq);
  }
}

// ------------------------------------------------------------------------------ finalize ----
struct FinalizeParams {
  const float *gathered;  // [G][B][shard_floats]
  // combine_partials > 0: `gathered` is the weighting kernel's per-CTA partial records [combine_partials][shard_floats] of
  // ONE controller (B == 1, G == 1); this kernel adds them up in the fixed order of sum_partials_fixed_order, which
  // spares the weighting kernel its fence / atomic-ticket / last-CTA pass (three dependent L2 round trips)
  int combine_partials;
  int is_nn64;  // 6-64-64-64-64-4 with 256 threads: CtaMlp<64, 4> (weights in registers)
  float *inbox;           // [B][inbox_stride]  (U is rewritten for the next iteration / resident step)
  float *outbox;          // [B][outbox_stride]: result[4] | U_smoothed[2T] | U_new[2T] | state_sol[7T] | ctrl_sol[2T]
  const float *theta_t;   // NN: transposed packed weights; BF: theta transposed [25][4]
  const float *theta_fold;  // 6-32-32-4: folded weights for WarpMlp32 (see RolloutParams)
  const int *net_structure;
  int num_layers;         // NN: entries of net_structure; 0 = basis-function model
  int is_nn32;            // the 6-32-32-4 network: nominal trajectory on the WarpMlp32 fast path
  int G, B, T, shard_floats, inbox_stride, outbox_stride;
  float gamma, dt, lo0, hi0, lo1, hi1;
  int negate_yaw;
  int last_iter;          // 1: smooth + nominal trajectory; 0: U <- U_new only (num_iters > 1)
  int feed_back;          // 1: write the smoothed U back into the inbox (resident stepping)
  unsigned int *baseline; // [B] re-armed (0xffffffff) for the next rollout launch
  uint32_t *call_counter; // advanced once per launch (Philox offset of the next sampler launch)
  // Peer-memory exchange: wait until every rank's flag for this controller reached `p2p_seq`, then read the mailbox half
  // `p2p_seq & 1` (gathered points at it).  A rank that never arrives trips the timeout and *p2p_error is set.
  const unsigned int *p2p_flags;  // this GPU's flags [2][G][B]; nullptr = no peer exchange
  unsigned int p2p_seq;
  unsigned int *p2p_error;
  // 0: the whole kernel.  Device-resident stepping splits it so that the nominal trajectory -- a 100-step dependent chain on
  // one warp that only produces outputs -- runs on a second stream beside the NEXT step's rollouts:
  // 1: combine + control update + smoothing (+ feed-back) only; 2: nominal trajectory only (reads the smoothed U from the outbox).
  int phase;
};

constexpr int FIN_MAX_WIDTH = 128;

// Host twin of the basis functions for the nominal trajectory (single lane; 100 steps).
__device__ __forceinline__ void car_basis_host_twin(const float *theta, const float *s, float u0, float u1, float *out4) {
  float in[6][1] = {{s[3]}, {s[4]}, {s[5]}, {s[6]}, {u0}, {u1}};
  float o[4][1];
  CarBasisDyn::deriv(theta, nullptr, in, o);
  for (int j = 0; j < 4; j++) out4[j] = o[j][0];
}

// Nominal trajectory for the 6-32-32-4 network on ONE warp (computeNominalTraj, PI/mppi_controller.cu:501-519):
// WarpMlp32 (warp_mlp.cuh) advances the recursive state; x / y need sincosf(yaw) only for the OUTPUT, so they are
// filled in 32 steps at a time by all lanes in parallel (sequential FMA prefix, the reference's Euler order).
// Within the 1e-4 parity tolerance of the host twin (FMA contraction, partial sums and tanh_fast differ from
// Eigen/libm in the last bits).
__device__ __forceinline__ void nominal_traj_nn32(const WarpMlp32 &net, const float *__restrict__ inbox,
                                                  const float *__restrict__ Usm, int T, float dt, int negate_yaw, float lo0,
                                                  float hi0, float lo1, float hi1, float *__restrict__ ssol,
                                                  float *__restrict__ csol, float *__restrict__ act, int lane) {
  const unsigned full = 0xffffffffu;
  float x = inbox[INBOX_STATE + 0], y = inbox[INBOX_STATE + 1], yaw = inbox[INBOX_STATE + 2];
  float roll = inbox[INBOX_STATE + 3], vx = inbox[INBOX_STATE + 4], vy = inbox[INBOX_STATE + 5], wz = inbox[INBOX_STATE + 6];
  for (int i0 = 0; i0 < T; i0 += 32) {
    const int nb = min(32, T - i0);
    // this lane's timestep: clamp the control (enforceConstraints) and publish it
    float u0m = 0.0f, u1m = 0.0f;
    if (lane < nb) {
      u0m = Usm[2 * (i0 + lane)]; u1m = Usm[2 * (i0 + lane) + 1];
      u0m = u0m < lo0 ? lo0 : (u0m > hi0 ? hi0 : u0m);
      u1m = u1m < lo1 ? lo1 : (u1m > hi1 ? hi1 : u1m);
      csol[2 * (i0 + lane)] = u0m; csol[2 * (i0 + lane) + 1] = u1m;
    }
    float r_yaw = 0.0f, r_roll = 0.0f, r_vx = 0.0f, r_vy = 0.0f, r_wz = 0.0f;
    // the controls of a timestep are fetched (shuffles) one timestep ahead: they do not depend on the state
    float u0n = __shfl_sync(full, u0m, 0), u1n = __shfl_sync(full, u1m, 0);
    for (int ii = 0; ii < nb; ii++) {
      const float u0 = u0n, u1 = u1n;
      u0n = __shfl_sync(full, u0m, (ii + 1) & 31); u1n = __shfl_sync(full, u1m, (ii + 1) & 31);
      if (lane == ii) { r_yaw = yaw; r_roll = roll; r_vx = vx; r_vy = vy; r_wz = wz; }
      float o0, o1, o2, o3;
      net.forward(roll, vx, vy, wz, u0, u1, act, ii & 1, lane, o0, o1, o2, o3);
      yaw = fmaf(negate_yaw ? -wz : wz, dt, yaw);
      roll = fmaf(o0, dt, roll); vx = fmaf(o1, dt, vx); vy = fmaf(o2, dt, vy); wz = fmaf(o3, dt, wz);
    }
    float sn, cs;
    sincosf(r_yaw, &sn, &cs);
    const float d0 = fmaf(cs, r_vx, -__fmul_rn(sn, r_vy)), d1 = fmaf(sn, r_vx, __fmul_rn(cs, r_vy));
    float myx = 0.0f, myy = 0.0f;
    // fully unrolled so the 64 shuffles are in flight together; lanes beyond the horizon hold zeros (exact no-ops)
#pragma unroll
    for (int j = 0; j < 32; j++) {
      if (lane == j) { myx = x; myy = y; }
      x = fmaf(__shfl_sync(full, d0, j), dt, x);
      y = fmaf(__shfl_sync(full, d1, j), dt, y);
    }
    if (lane < nb) {
      float *o = ssol + (size_t)(i0 + lane) * S_DIM;
      o[0] = myx; o[1] = myy; o[2] = r_yaw; o[3] = r_roll; o[4] = r_vx; o[5] = r_vy; o[6] = r_wz;
    }
  }
}

// Nominal trajectory for a 6 -> W x NH (tanh) -> 4 network of compile-time shape (the fork's 6-64-64-64-64-4) on a
// 256-thread CTA with the weights in REGISTERS: thread (j, g) = (tid mod W, tid / W) holds, of every hidden layer, the
// weights of neuron j for the g-th quarter of its inputs (W/4 per layer) and sums that quarter with four interleaved FMA
// chains; the quarters meet in shared memory, threads 0..W-1 add them in a fixed order, add the bias and apply tanh_fast.
// Per layer the dependent chain is 4 broadcast LDS.128, W/16 FMAs deep, and two CTA barriers.  Only roll, u_x, u_y and the
// yaw rate feed the network, so every thread integrates those four (identical arithmetic, no broadcast) while thread 0 alone
// integrates x, y, yaw (precise sinf / cosf) and writes the solution.  Same arithmetic class as WarpMlp32.
template <int W, int NH>
struct CtaMlp {
  static_assert(W == 64, "one neuron per thread of a 64-thread group, four input quarters over 256 threads");
  static constexpr int Q = W / 4;
  static constexpr int TH_B1 = 6 * W;
  __host__ __device__ static constexpr int th_w(int h) { return 7 * W + (h - 1) * (W * W + W); }
  static constexpr int TH_WL = 7 * W + (NH - 1) * (W * W + W), TH_BL = TH_WL + 4 * W;
  float w1[6], wh[NH - 1][Q], wl[Q], b[NH], bl[4];

  __device__ __forceinline__ void load(const float *__restrict__ th, int j, int g) {
#pragma unroll
    for (int k = 0; k < 6; k++) w1[k] = th[k * W + j];
    b[0] = th[TH_B1 + j];
#pragma unroll
    for (int h = 1; h < NH; h++) {
#pragma unroll
      for (int i = 0; i < Q; i++) wh[h - 1][i] = th[th_w(h) + (g * Q + i) * W + j];
      b[h] = th[th_w(h) + W * W + j];
    }
#pragma unroll
    for (int i = 0; i < Q; i++) wl[i] = th[TH_WL + (g * Q + i) * 4 + (j & 3)];
#pragma unroll
    for (int k = 0; k < 4; k++) bl[k] = th[TH_BL + k];
  }

  __device__ __forceinline__ static float dot_quarter(const float (&w)[Q], const float *__restrict__ x) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int i = 0; i < Q; i += 4) {
      const float4 v = *reinterpret_cast<const float4 *>(x + i);
      a0 = fmaf(w[i], v.x, a0); a1 = fmaf(w[i + 1], v.y, a1); a2 = fmaf(w[i + 2], v.z, a2); a3 = fmaf(w[i + 3], v.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
  }

  // act: [2][W] activations, part: [4][W] partial sums (shared memory).  All 256 threads call this.
  __device__ __forceinline__ void trajectory(const float *__restrict__ inbox, const float *__restrict__ Usm, int T, float dt, int negate_yaw,
                                             float lo0, float hi0, float lo1, float hi1, float *__restrict__ ssol,
                                             float *__restrict__ csol, float *__restrict__ act, float *__restrict__ part, int tid) const {
    const int j = tid % W, g = tid / W;
    float x = inbox[INBOX_STATE + 0], y = inbox[INBOX_STATE + 1], yaw = inbox[INBOX_STATE + 2];
    float roll = inbox[INBOX_STATE + 3], vx = inbox[INBOX_STATE + 4], vy = inbox[INBOX_STATE + 5], wz = inbox[INBOX_STATE + 6];
    for (int i = 0; i < T; i++) {
      float u0 = Usm[2 * i], u1 = Usm[2 * i + 1];
      u0 = u0 < lo0 ? lo0 : (u0 > hi0 ? hi0 : u0);
      u1 = u1 < lo1 ? lo1 : (u1 > hi1 ? hi1 : u1);
      if (tid == 0) {
        float *o = ssol + (size_t)i * S_DIM;
        o[0] = x; o[1] = y; o[2] = yaw; o[3] = roll; o[4] = vx; o[5] = vy; o[6] = wz;
        csol[2 * i] = u0; csol[2 * i + 1] = u1;
        float sn, cs;
        sincosf(yaw, &sn, &cs);
        x = fmaf(fmaf(cs, vx, -__fmul_rn(sn, vy)), dt, x);
        y = fmaf(fmaf(sn, vx, __fmul_rn(cs, vy)), dt, y);
      }
      yaw = fmaf(negate_yaw ? -wz : wz, dt, yaw);
      // layer 1 (6 inputs, straight from registers; k ascending, bias last)
      if (g == 0) {
        float t = w1[0] * roll;
        t = fmaf(w1[1], vx, t); t = fmaf(w1[2], vy, t); t = fmaf(w1[3], wz, t); t = fmaf(w1[4], u0, t); t = fmaf(w1[5], u1, t);
        act[j] = tanh_fast(t + b[0]);
      }
      __syncthreads();
#pragma unroll
      for (int h = 1; h < NH; h++) {
        const float *cur = act + ((h - 1) & 1) * W;
        float *nxt = act + (h & 1) * W;
        part[g * W + j] = dot_quarter(wh[h - 1], cur + g * Q);
        __syncthreads();
        if (g == 0) nxt[j] = tanh_fast((((part[j] + part[W + j]) + (part[2 * W + j] + part[3 * W + j]))) + b[h]);
        __syncthreads();
      }
      // output layer: threads j < 4 of every quarter, then every thread adds the quarters itself
      const float *last = act + ((NH - 1) & 1) * W;
      if (j < 4) part[g * W + j] = dot_quarter(wl, last + g * Q);
      __syncthreads();
      const float o0 = ((part[0] + part[W + 0]) + (part[2 * W + 0] + part[3 * W + 0])) + bl[0];
      const float o1 = ((part[1] + part[W + 1]) + (part[2 * W + 1] + part[3 * W + 1])) + bl[1];
      const float o2 = ((part[2] + part[W + 2]) + (part[2 * W + 2] + part[3 * W + 2])) + bl[2];
      const float o3 = ((part[3] + part[W + 3]) + (part[2 * W + 3] + part[3 * W + 3])) + bl[3];
      roll = fmaf(o0, dt, roll); vx = fmaf(o1, dt, vx); vy = fmaf(o2, dt, vy); wz = fmaf(o3, dt, wz);
      __syncthreads();  // the partial sums are consumed before the next step overwrites them
    }
  }
};

// Nominal trajectory for a runtime-layer network (the fork's 6-64-64-64-64-4) on the WHOLE CTA: thread (j, g) = (tid mod
// 64, tid / 64) sums the g-th slice of the inputs of neuron j (four interleaved FMA partial sums), the slices meet in
// shared memory and threads 0..63 add them in a fixed order, add the bias and apply tanh_fast.  One warp alone issued the
// 2 480 instructions of a timestep back to back (3.2 us per step, 0.32 ms per trajectory); spread over 8 warps the chain
// per layer is 16 FMAs and two CTA barriers.  Every thread integrates its own copy of the state (identical arithmetic),
// so nothing but the activations crosses threads.  Same arithmetic class as WarpMlp32 (FMA, partial sums, tanh_fast).
__device__ __forceinline__ void nominal_traj_mlp_cta(const float *__restrict__ sw, const int *__restrict__ ns, int num_layers,
                                                     const float *__restrict__ inbox, const float *__restrict__ Usm, int T, float dt,
                                                     int negate_yaw, float lo0, float hi0, float lo1, float hi1,
                                                     float *__restrict__ ssol, float *__restrict__ csol, float *__restrict__ act,
                                                     float *__restrict__ part /* [4][64] */, int tid, int nthr) {
  const int jl = tid & 63, g = tid >> 6, ng = nthr >= 256 ? 4 : 1, lg = nthr >= 256 ? 2 : 0;
  const bool worker = tid < ng * 64;
  float s[S_DIM];
  for (int k = 0; k < S_DIM; k++) s[k] = inbox[INBOX_STATE + k];
  for (int i = 0; i < T; i++) {
    float u0 = Usm[2 * i], u1 = Usm[2 * i + 1];
    u0 = u0 < lo0 ? lo0 : (u0 > hi0 ? hi0 : u0);
    u1 = u1 < lo1 ? lo1 : (u1 > hi1 ? hi1 : u1);
    float *cur = act, *nxt = act + FIN_MAX_WIDTH;
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < S_DIM; k++) ssol[i * S_DIM + k] = s[k];
      csol[2 * i] = u0; csol[2 * i + 1] = u1;
      cur[0] = s[3]; cur[1] = s[4]; cur[2] = s[5]; cur[3] = s[6]; cur[4] = u0; cur[5] = u1;
    }
    __syncthreads();
    const float *W = sw;
    for (int l = 0; l + 1 < num_layers; l++) {
      const int nin = ns[l], nout = ns[l + 1];
      const int per = (nin + ng - 1) >> lg, k0 = g * per, k1 = min(nin, k0 + per);  // ng is 1 or 4: no integer division
      for (int j0 = 0; j0 < nout; j0 += 64) {
        const int j = j0 + jl;
        if (worker && j < nout) {
          const float *wp = W + j + k0 * nout, *cp = cur + k0;  // walked with pointer increments (no per-load multiplies)
          float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
          int n = k1 - k0;
          for (; n >= 4; n -= 4, wp += 4 * nout, cp += 4) {
            const float w0 = wp[0], w1 = wp[nout], w2 = wp[2 * nout], w3 = wp[3 * nout];
            a0 = fmaf(w0, cp[0], a0); a1 = fmaf(w1, cp[1], a1); a2 = fmaf(w2, cp[2], a2); a3 = fmaf(w3, cp[3], a3);
          }
          for (; n > 0; n--, wp += nout, cp++) a0 = fmaf(wp[0], cp[0], a0);
          part[g * 64 + jl] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        if (tid < 64 && j < nout) {
          float t = part[jl];
          for (int q = 1; q < ng; q++) t += part[q * 64 + jl];
          t += W[nin * nout + j];
          if (l + 2 < num_layers) t = tanh_fast(t);
          nxt[j] = t;
        }
        __syncthreads();
      }
      W += (nin + 1) * nout;
      float *tmp = cur; cur = nxt; nxt = tmp;
    }
    const float o0 = cur[0], o1 = cur[1], o2 = cur[2], o3 = cur[3];
    const float cs = cosf(s[2]), sn = sinf(s[2]);
    const float d0 = __fsub_rn(__fmul_rn(cs, s[4]), __fmul_rn(sn, s[5]));
    const float d1 = __fadd_rn(__fmul_rn(sn, s[4]), __fmul_rn(cs, s[5]));
    const float d2 = negate_yaw ? -s[6] : s[6];
    s[0] = __fadd_rn(s[0], __fmul_rn(d0, dt));
    s[1] = __fadd_rn(s[1], __fmul_rn(d1, dt));
    s[2] = __fadd_rn(s[2], __fmul_rn(d2, dt));
    s[3] = __fadd_rn(s[3], __fmul_rn(o0, dt)); s[4] = __fadd_rn(s[4], __fmul_rn(o1, dt));
    s[5] = __fadd_rn(s[5], __fmul_rn(o2, dt)); s[6] = __fadd_rn(s[6], __fmul_rn(o3, dt));
    __syncthreads();  // everybody has read the outputs before thread 0 overwrites the input slots of the next step
  }
}

// grid B, block 256 (64 when many controllers are batched: only warp 0 does the long part).  Combines the G shard records (log-sum-exp rescale),
// U_new = W / Z (control update :663-667), Savitzky-Golay (:468-499), then warp 0 integrates the
// nominal trajectory with the host-twin arithmetic (separate multiply and add, precise tanhf/sinf/cosf).
__global__ void bump_counter_kernel(uint32_t *counter) { *counter += 1u; }

// NET: 32 = NeuralNetModel<7,2,3,6,32,32,4> (WarpMlp32), 64 = 6-64-64-64-64-4 on 256 threads (CtaMlp<64,4>), 0 = anything else
// (basis functions, runtime-layer networks).  One instantiation per kind: each carries only its own weights in registers.
template <int NET>
__global__ void __launch_bounds__(256, NET == 0 ? 1 : 2) finalize_kernel(const __grid_constant__ FinalizeParams p) {
  extern __shared__ float fsm[];
  const int T = p.T, tid = threadIdx.x, b = blockIdx.x, nthr = blockDim.x;
  float *Unew = fsm;                 // [2T]
  float *Usm = Unew + 2 * T;         // [2T]
  float *act = Usm + 2 * T;          // [2][FIN_MAX_WIDTH]
  float *sw = act + 2 * FIN_MAX_WIDTH;  // staged parameters
  __shared__ float rec_sm[4 + 2 * 256 + 4];  // the combined record (combine_partials; T <= 256)
  __shared__ int ns_sm[16];                 // layer widths of the runtime-layer network
  if (tid < 16 && tid < p.num_layers) ns_sm[tid] = p.net_structure[tid];
  __shared__ float scale[64];
  __shared__ float hdr[4];
  float *inbox = p.inbox + (size_t)b * p.inbox_stride;
  float *outbox = p.outbox + (size_t)b * p.outbox_stride;
  // warp 0 fetches its slices of the network while the other warps combine the shard records
  WarpMlp32 net;
  if (NET == 32 && p.last_iter && p.phase != 1 && tid < 32) net.load(p.theta_fold, tid);
  CtaMlp<64, 4> net64;
  if (NET == 64 && p.last_iter && p.phase != 1) net64.load(p.theta_t, tid & 63, tid >> 6);
  pdl_trigger();  // the next step's first kernel may run its prologue (weights -> shared memory, tensor-memory allocation) now
  pdl_wait();  // the shard records come from the weighting kernel (or the exchange)
  if (p.phase == 2) {  // nominal trajectory only: the smoothed controls were left in the outbox by the phase-1 launch
    for (int k = tid; k < 2 * T; k += nthr) Usm[k] = outbox[4 + k];
  }
  if (p.phase != 2) {
  if (p.p2p_flags != nullptr) {
    if (tid < p.G) {
      const unsigned int *flag = p.p2p_flags + ((size_t)(p.p2p_seq & 1u) * p.G + tid) * p.B + b;
      // bounded wait on the wall clock: ranks may be seconds apart on their first call; a dead peer must not hang the GPU
      const unsigned long long t0 = globaltimer_ns();
      while (ld_acquire_sys(flag) != p.p2p_seq) {
        __nanosleep(100);
        if (globaltimer_ns() - t0 > P2P_TIMEOUT_NS) { atomicExch(p.p2p_error, 1u); break; }
      }
    }
    __syncthreads();
  }

  const float *gathered = p.gathered;
  if (p.combine_partials > 0) {
    for (int k = tid; k < p.shard_floats; k += nthr) {
      float v;
      if (k == 0) v = __ldcg(p.gathered);                                  // every CTA used the same global baseline
      else if (k == 3 || k >= SHARD_HDR + 2 * T) v = 0.0f;
      else v = sum_partials_fixed_order(p.gathered, p.combine_partials, p.shard_floats, k);
      rec_sm[k] = v;
    }
    __syncthreads();
    gathered = rec_sm;
  }

  if (tid == 0) {
    float base = gathered[((size_t)0 * p.B + b) * p.shard_floats];
    for (int g = 1; g < p.G; g++) base = fminf(base, gathered[((size_t)g * p.B + b) * p.shard_floats]);
    float Z = 0.0f, Q = 0.0f;
    for (int g = 0; g < p.G; g++) {
      const float *rec = gathered + ((size_t)g * p.B + b) * p.shard_floats;
      const float sg = (p.G == 1) ? 1.0f : expf(-p.gamma * (rec[0] - base));
      scale[g] = sg;
      Z = fmaf(sg, rec[1], Z);
      Q = fmaf(sg * sg, rec[2], Q);
    }
    hdr[0] = base; hdr[1] = Z; hdr[2] = Q / Z; hdr[3] = 0.0f;
  }
  __syncthreads();
  const float Z = hdr[1];
  for (int k = tid; k < 2 * T; k += nthr) {
    float wsum = 0.0f;
    for (int g = 0; g < p.G; g++) wsum = fmaf(scale[g], gathered[((size_t)g * p.B + b) * p.shard_floats + SHARD_HDR + k], wsum);
    Unew[k] = wsum / Z;
  }
  if (tid < 4) outbox[tid] = hdr[tid];
  if (tid == 0) {
    p.baseline[b] = 0xffffffffu;
    if (b == 0) *p.call_counter += 1u;
  }
  __syncthreads();
  for (int k = tid; k < 2 * T; k += nthr) outbox[4 + 2 * T + k] = Unew[k];
  if (!p.last_iter) {
    for (int k = tid; k < 2 * T; k += nthr) inbox[INBOX_U + k] = Unew[k];
    return;
  }
  // Savitzky-Golay: P = [hist0, hist1, U_0 .. U_{T-1}, U_{T-1}, U_{T-1}], taps [-3 12 17 12 -3]/35
  const float f0 = -3.0f / 35.0f, f1 = 12.0f / 35.0f, f2 = 17.0f / 35.0f;
  for (int k = tid; k < 2 * T; k += nthr) {
    const int i = k >> 1, j = k & 1;
    float pv[5];
#pragma unroll
    for (int m = 0; m < 5; m++) {
      const int ii = i + m;  // index into P
      pv[m] = ii < 2 ? inbox[INBOX_HIST + 2 * ii + j] : (ii < T + 2 ? Unew[2 * (ii - 2) + j] : Unew[2 * (T - 1) + j]);
    }
    float acc = __fmul_rn(f0, pv[0]);
    acc = __fadd_rn(acc, __fmul_rn(f1, pv[1]));
    acc = __fadd_rn(acc, __fmul_rn(f2, pv[2]));
    acc = __fadd_rn(acc, __fmul_rn(f1, pv[3]));
    acc = __fadd_rn(acc, __fmul_rn(f0, pv[4]));
    Usm[k] = acc;
    outbox[4 + k] = acc;
  }
  }  // phase != 2
  if (p.phase == 1) {
    __syncthreads();
    if (p.feed_back)
      for (int k = tid; k < 2 * T; k += nthr) inbox[INBOX_U + k] = Usm[k];
    return;
  }
  // stage the model parameters for the generic nominal trajectory (the 6-32-32-4 path holds them in registers)
  if (NET == 0) {
    int nparams = 100;
    if (p.num_layers > 0) {
      nparams = 0;
      for (int l = 0; l + 1 < p.num_layers; l++) nparams += (p.net_structure[l] + 1) * p.net_structure[l + 1];
    }
    for (int k = tid; k < nparams; k += nthr) sw[k] = p.theta_t[k];
  }
  __syncthreads();
  if (p.feed_back && p.phase == 0)
    for (int k = tid; k < 2 * T; k += nthr) inbox[INBOX_U + k] = Usm[k];
  if (NET == 64) {
    __shared__ __align__(16) float act64[2 * 64];
    __shared__ float part64[4 * 64];
    net64.trajectory(inbox, Usm, T, p.dt, p.negate_yaw, p.lo0, p.hi0, p.lo1, p.hi1, outbox + 4 + 4 * T, outbox + 4 + 4 * T + S_DIM * T,
                     act64, part64, tid);
    return;
  }
  if (NET == 0 && p.num_layers > 0) {  // runtime-layer network: the whole CTA works on the trajectory
    __shared__ float part_sm[4 * 64];
    nominal_traj_mlp_cta(sw, ns_sm, p.num_layers, inbox, Usm, T, p.dt, p.negate_yaw, p.lo0, p.hi0, p.lo1, p.hi1,
                         outbox + 4 + 4 * T, outbox + 4 + 4 * T + S_DIM * T, act, part_sm, tid, nthr);
    return;
  }
  if (tid >= 32) return;
  // ---- nominal trajectory (computeNominalTraj :501-519 -> host updateState) on warp 0 ----
  const int lane = tid;
  float *ssol = outbox + 4 + 4 * T;
  float *csol = ssol + S_DIM * T;
  if (NET == 32) {
    nominal_traj_nn32(net, inbox, Usm, T, p.dt, p.negate_yaw, p.lo0, p.hi0, p.lo1, p.hi1, ssol, csol, act, lane);
    return;
  }
  float s[S_DIM];
  for (int k = 0; k < S_DIM; k++) s[k] = inbox[INBOX_STATE + k];
  for (int i = 0; i < T; i++) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < S_DIM; k++) ssol[i * S_DIM + k] = s[k];
    }
    float u0 = Usm[2 * i], u1 = Usm[2 * i + 1];
    u0 = u0 < p.lo0 ? p.lo0 : (u0 > p.hi0 ? p.hi0 : u0);
    u1 = u1 < p.lo1 ? p.lo1 : (u1 > p.hi1 ? p.hi1 : u1);
    if (lane == 0) { csol[2 * i] = u0; csol[2 * i + 1] = u1; }
    float dyn[4];
    if (p.num_layers > 0) {
      // runtime-layer MLP (the 6-64-64-64-64-4 network): lane j owns neurons j and j + 32 of each layer, four interleaved
      // partial sums each (FMA, tanh_fast: the arithmetic of WarpMlp32, inside the 1e-4 tolerance of the host twin);
      // activations ping-pong through shared memory, weights are staged there once per launch
      float *cur = act, *nxt = act + FIN_MAX_WIDTH;
      if (lane == 0) { cur[0] = s[3]; cur[1] = s[4]; cur[2] = s[5]; cur[3] = s[6]; cur[4] = u0; cur[5] = u1; }
      __syncwarp();
      const float *W = sw;
      for (int l = 0; l + 1 < p.num_layers; l++) {
        const int nin = ns_sm[l], nout = ns_sm[l + 1];
        for (int j0 = lane; j0 < nout; j0 += 64) {
          const int j1 = j0 + 32;
          const bool two = j1 < nout;
          const float *w0 = W + j0, *w1 = W + (two ? j1 : j0);
          float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
          int k = 0;
#pragma unroll 4
          for (; k + 3 < nin; k += 4) {  // 4 iterations = 36 shared-memory loads in flight ahead of their FMAs
            const float4 x = *reinterpret_cast<const float4 *>(cur + k);
            a0 = fmaf(w0[(k + 0) * nout], x.x, a0); a1 = fmaf(w0[(k + 1) * nout], x.y, a1);
            a2 = fmaf(w0[(k + 2) * nout], x.z, a2); a3 = fmaf(w0[(k + 3) * nout], x.w, a3);
            c0 = fmaf(w1[(k + 0) * nout], x.x, c0); c1 = fmaf(w1[(k + 1) * nout], x.y, c1);
            c2 = fmaf(w1[(k + 2) * nout], x.z, c2); c3 = fmaf(w1[(k + 3) * nout], x.w, c3);
          }
          for (; k < nin; k++) { a0 = fmaf(w0[k * nout], cur[k], a0); c0 = fmaf(w1[k * nout], cur[k], c0); }
          float t0 = ((a0 + a1) + (a2 + a3)) + w0[nin * nout];
          float t1 = ((c0 + c1) + (c2 + c3)) + w1[nin * nout];
          if (l + 2 < p.num_layers) { t0 = tanh_fast(t0); t1 = tanh_fast(t1); }
          nxt[j0] = t0;
          if (two) nxt[j1] = t1;
        }
        __syncwarp();
        W += (nin + 1) * nout;
        float *tmp = cur; cur = nxt; nxt = tmp;
      }
      for (int k = 0; k < 4; k++) dyn[k] = cur[k];
      __syncwarp();
    } else {
      car_basis_host_twin(sw, s, u0, u1, dyn);
    }
    const float cs = cosf(s[2]), sn = sinf(s[2]);
    const float d0 = __fsub_rn(__fmul_rn(cs, s[4]), __fmul_rn(sn, s[5]));
    const float d1 = __fadd_rn(__fmul_rn(sn, s[4]), __fmul_rn(cs, s[5]));
    const float d2 = (p.num_layers == 0 || p.negate_yaw) ? -s[6] : s[6];
    s[0] = __fadd_rn(s[0], __fmul_rn(d0, p.dt));
    s[1] = __fadd_rn(s[1], __fmul_rn(d1, p.dt));
    s[2] = __fadd_rn(s[2], __fmul_rn(d2, p.dt));
    for (int k = 0; k < 4; k++) s[3 + k] = __fadd_rn(s[3 + k], __fmul_rn(dyn[k], p.dt));
  }
}

}  // namespace mppi
