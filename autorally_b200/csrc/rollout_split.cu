// rollout_split.cu -- latency variant of the fused rollout kernel for NeuralNetModel<7,2,3,6,32,32,4>:
// ONE ROLLOUT ACROSS 8 LANES (the role BDIM_Y plays in the reference, PI/mppi_controller.cu:275-278).
//
// At the reference's own size (1920 rollouts) one thread per rollout gives 60 warps -- 10% of the 592
// SM sub-partitions, each walking a ~2000-instruction serial chain per timestep.  Here a warp carries 4
// rollouts: lane l of a group owns hidden neurons 4l..4l+3 of both hidden layers, so the 32x32 layer is
// 128 FFMA per lane instead of 1024, and 1920 rollouts become 480 warps spread over 120 SMs.
//   layer 1: 4 neurons x 6 inputs per lane, weights in registers (state is replicated in the 8 lanes);
//   exchange: one STS.128 + __syncwarp + 8 LDS.128 through a double-buffered 128-byte slot per rollout;
//   layer 2: 4 neurons x 32 inputs per lane, weights LDS.128 from the transposed table (the 8 lanes of a
//            group read 128 contiguous bytes; the 4 groups of a warp broadcast), k ascending as in the
//            reference (PI/neural_net_model.cu:388-399);
//   layer 3: each lane multiplies its own 4 hidden values into the 4 outputs, 3 xor-shuffles add the 8
//            partials (every lane ends with bit-identical sums, so the replicated state stays coherent).
// Noise is read and the sampled controls are written back 8 timesteps at a time (lane l holds step
// i0+l: one coalesced 64-byte access per rollout instead of 8 dependent 8-byte ones).
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {

namespace {
constexpr int kNP = 1412;          // 6-32-32-4 packed transposed parameters
constexpr int kW1 = 0, kB1 = 192, kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;
constexpr int kBlock = 128;        // 4 warps = 16 rollouts per CTA, one warp per SM sub-partition
}  // namespace

__global__ void __launch_bounds__(kBlock) rollout_split8_kernel(const __grid_constant__ RolloutParams p) {
  __shared__ float4 sw4[kNP / 4];
  __shared__ float4 xch4[(kBlock / 8) * 2 * 8];  // per rollout: 2 buffers x 32 floats
  __shared__ float2 sU[512];
  const int tid = threadIdx.x, lane = tid & 31, l = lane & 7;
  for (int i = tid; i < kNP / 4; i += kBlock) sw4[i] = reinterpret_cast<const float4 *>(p.theta_t)[i];
  const long long gro = (long long)blockIdx.x * (kBlock / 8) + (tid >> 3);  // flat rollout index over B * n_local
  const int ctrl = (int)(gro / p.n_local);
  const int lr = (int)(gro - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  // a CTA's 16 rollouts belong to one controller (n_local % 64 == 0): stage its nominal controls once
  const int T = p.T;
  const bool u_in_smem = T <= 512;
  if (u_in_smem)
    for (int i = tid; i < T; i += kBlock) sU[i] = reinterpret_cast<const float2 *>(inbox + INBOX_U)[i];
  __syncthreads();
  const float *sw = reinterpret_cast<const float *>(sw4);
  const float2 *Ug = reinterpret_cast<const float2 *>(inbox + INBOX_U);

  // per-lane weight slices kept in registers for all T steps
  float4 w1[6], w3[4];
#pragma unroll
  for (int k = 0; k < 6; k++) w1[k] = sw4[(kW1 + k * 32) / 4 + l];
  const float4 b1 = sw4[kB1 / 4 + l], b2 = sw4[kB2 / 4 + l], b3 = sw4[kB3 / 4];
#pragma unroll
  for (int kk = 0; kk < 4; kk++) w3[kk] = sw4[kW3 / 4 + 4 * l + kk];  // Wt3[k][0..3], k = 4l + kk
  const float4 *w2 = sw4 + kW2 / 4 + l;                                 // Wt2[k][4l..4l+3] at w2[8 k]
  float4 *xbuf = xch4 + (tid >> 3) * 16;                               // this rollout's two 8-float4 buffers

  float s[S_DIM];
#pragma unroll
  for (int k = 0; k < S_DIM; k++) s[k] = inbox[INBOX_STATE + k];
  float running = 0.0f;
  int crash = 0;
  const int rg = p.r_begin + lr;
  const bool noise_free = (rg == 0), pure_noise = (rg >= p.pure_noise_from);
  float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)gro * T;
  const unsigned full = 0xffffffffu;
  const int group_base = lane & ~7;

  for (int i0 = 0; i0 < T; i0 += 8) {
    const bool mine = i0 + l < T;
    const float2 e_mine = mine ? row[i0 + l] : make_float2(0.0f, 0.0f);
    float2 wb = make_float2(0.0f, 0.0f);
    const int nb = min(8, T - i0);
    for (int ii = 0; ii < nb; ii++) {
      const int i = i0 + ii;
      const float ex = __shfl_sync(full, e_mine.x, group_base | ii);
      const float ey = __shfl_sync(full, e_mine.y, group_base | ii);
      const float2 Ui = u_in_smem ? sU[i] : Ug[i];
      float du0, du1, u0, u1;
      if (noise_free || i < p.opt_delay) {
        du0 = 0.0f; du1 = 0.0f; u0 = Ui.x; u1 = Ui.y;
      } else if (pure_noise) {
        du0 = __fmul_rn(ex, p.nu0); du1 = __fmul_rn(ey, p.nu1); u0 = du0; u1 = du1;
      } else {
        du0 = __fmul_rn(ex, p.nu0); du1 = __fmul_rn(ey, p.nu1);
        u0 = __fadd_rn(Ui.x, du0); u1 = __fadd_rn(Ui.y, du1);
      }
      if (l == ii) wb = make_float2(u0, u1);  // un-clamped write-back (PI/mppi_controller.cu:153)
      u0 = u0 < p.lo0 ? p.lo0 : (u0 > p.hi0 ? p.hi0 : u0);
      u1 = u1 < p.lo1 ? p.lo1 : (u1 > p.hi1 ? p.hi1 : u1);
      if (i > 0) {
        const float c = running_cost_step(p.cp, p.tex, s, u0, u1, du0, du1, p.nu0, p.nu1, crash);
        running = (float)((double)running + (double)__fsub_rn(c, running) * p.inv_step[i]);
      }
      // ---- layer 1: neurons 4l..4l+3 ----
      const float a0[6] = {s[3], s[4], s[5], s[6], u0, u1};
      float4 h = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < 6; k++) {
        h.x = fmaf(w1[k].x, a0[k], h.x); h.y = fmaf(w1[k].y, a0[k], h.y);
        h.z = fmaf(w1[k].z, a0[k], h.z); h.w = fmaf(w1[k].w, a0[k], h.w);
      }
      h.x = tanh_fast(h.x + b1.x); h.y = tanh_fast(h.y + b1.y);
      h.z = tanh_fast(h.z + b1.z); h.w = tanh_fast(h.w + b1.w);
      float4 *buf = xbuf + (i & 1) * 8;
      buf[l] = h;
      __syncwarp();
      // ---- layer 2: neurons 4l..4l+3 over all 32 inputs, k ascending ----
      float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
      for (int k4 = 0; k4 < 8; k4++) {
        const float4 hv = buf[k4];
        const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const float4 w = w2[8 * (4 * k4 + kk)];
          acc.x = fmaf(w.x, hk[kk], acc.x); acc.y = fmaf(w.y, hk[kk], acc.y);
          acc.z = fmaf(w.z, hk[kk], acc.z); acc.w = fmaf(w.w, hk[kk], acc.w);
        }
      }
      const float g0 = tanh_fast(acc.x + b2.x), g1 = tanh_fast(acc.y + b2.y);
      const float g2 = tanh_fast(acc.z + b2.z), g3 = tanh_fast(acc.w + b2.w);
      // ---- layer 3: partial sums over this lane's 4 hidden units, then xor-butterfly over the group ----
      float4 o;
      o.x = w3[0].x * g0; o.y = w3[0].y * g0; o.z = w3[0].z * g0; o.w = w3[0].w * g0;
      o.x = fmaf(w3[1].x, g1, o.x); o.y = fmaf(w3[1].y, g1, o.y); o.z = fmaf(w3[1].z, g1, o.z); o.w = fmaf(w3[1].w, g1, o.w);
      o.x = fmaf(w3[2].x, g2, o.x); o.y = fmaf(w3[2].y, g2, o.y); o.z = fmaf(w3[2].z, g2, o.z); o.w = fmaf(w3[2].w, g2, o.w);
      o.x = fmaf(w3[3].x, g3, o.x); o.y = fmaf(w3[3].y, g3, o.y); o.z = fmaf(w3[3].z, g3, o.z); o.w = fmaf(w3[3].w, g3, o.w);
#pragma unroll
      for (int m = 1; m < 8; m <<= 1) {
        o.x += __shfl_xor_sync(full, o.x, m); o.y += __shfl_xor_sync(full, o.y, m);
        o.z += __shfl_xor_sync(full, o.z, m); o.w += __shfl_xor_sync(full, o.w, m);
      }
      o.x += b3.x; o.y += b3.y; o.z += b3.z; o.w += b3.w;
      // ---- kinematics + Euler step + crash check (replicated in the 8 lanes) ----
      float sn, cs;
      sincosf(s[2], &sn, &cs);
      const float d0 = fmaf(cs, s[4], -__fmul_rn(sn, s[5]));
      const float d1 = fmaf(sn, s[4], __fmul_rn(cs, s[5]));
      const float d2 = p.negate_yaw ? -s[6] : s[6];
      s[0] = fmaf(d0, p.dt, s[0]); s[1] = fmaf(d1, p.dt, s[1]); s[2] = fmaf(d2, p.dt, s[2]);
      s[3] = fmaf(o.x, p.dt, s[3]); s[4] = fmaf(o.y, p.dt, s[4]); s[5] = fmaf(o.z, p.dt, s[5]); s[6] = fmaf(o.w, p.dt, s[6]);
      if (fabsf(s[3]) >= 1.57f) crash = 1;
    }
    if (mine) row[i0 + l] = wb;
  }
  if (l == 0) {
    p.costs[gro] = running;
    p.crash[gro] = (unsigned char)crash;
  }
  const unsigned int wbest = __reduce_min_sync(full, float_to_ordered(running));
  if (lane == 0) atomicMin(p.baseline + ctrl, wbest);
}

cudaError_t launch_rollout_nn32_split8(const RolloutParams &p, cudaStream_t st) {
  const long long total = (long long)p.B * p.n_local;  // multiple of 64, so every CTA is full
  const unsigned grid = (unsigned)(total / (kBlock / 8));
  rollout_split8_kernel<<<grid, kBlock, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace mppi
