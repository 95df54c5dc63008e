// warp_mlp.cuh -- the 6-32-32-4 tanh MLP of NeuralNetModel<7,2,3,6,32,32,4> evaluated by ONE WARP for ONE input
// vector with a short dependent chain: the nominal trajectory of finalize_kernel (weighting.cuh).
//
// Lane l owns hidden neuron l of both hidden layers and output (l & 3) of the last layer, with its weight slices in
// registers for the whole trajectory (72 registers).  Activations cross lanes through a double-buffered 128-byte
// shared-memory slot per layer (one __syncwarp per exchange).  Layers 2 and 3 accumulate their 32 products in four
// interleaved partial sums (k mod 4), so the FMA chain is 8 deep instead of the reference's 32
// (PI/neural_net_model.cu:388-399 sums k ascending): the result differs from the sequential order in the last bits
// (~1e-7 relative), far inside the 1e-4 parity tolerance; the throughput kernels (dynamics.cuh) keep the sequential order.
#pragma once
#include "device_common.cuh"

namespace mppi {

struct WarpMlp32 {
  static constexpr int kW1 = 0, kB1 = 192, kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;  // packed transposed layout
  float w1[6], w2[32], w3[32];
  float b1, b2, b3;

  // theta_t: [W1t 6x32 | b1 32 | W2t 32x32 | b2 32 | W3t 32x4 | b3 4] (global or shared memory) in its FOLDED form (fold_nn32,
  // mppi_b200.cu: tanh scale in layers 1 and 2, the map 1 - 2 r in layers 2 and 3), so the activations that cross lanes are
  // r = 1 / (2^x + 1) and each activation is three dependent instructions; loads are coalesced
  __device__ __forceinline__ void load(const float *__restrict__ theta_t, int lane) {
#pragma unroll
    for (int k = 0; k < 6; k++) w1[k] = theta_t[kW1 + k * 32 + lane];
#pragma unroll
    for (int k = 0; k < 32; k++) w2[k] = theta_t[kW2 + k * 32 + lane];
    const int jo = lane & 3;
#pragma unroll
    for (int k = 0; k < 32; k++) w3[k] = theta_t[kW3 + k * 4 + jo];
    b1 = theta_t[kB1 + lane];
    b2 = theta_t[kB2 + lane];
    b3 = theta_t[kB3 + jo];
  }

  // xbuf: 128 floats of shared memory owned by this warp (h1[2][32], h2[2][32]); parity = step & 1.
  // Inputs are replicated in all lanes; the four outputs come back replicated in all lanes.
  __device__ __forceinline__ void forward(float roll, float vx, float vy, float wz, float u0, float u1, float *xbuf,
                                          int parity, int lane, float &o0, float &o1, float &o2, float &o3) const {
    const unsigned full = 0xffffffffu;
    // layer 1 (k ascending, bias last: the reference's order)
    float t = w1[0] * roll;
    t = fmaf(w1[1], vx, t); t = fmaf(w1[2], vy, t); t = fmaf(w1[3], wz, t); t = fmaf(w1[4], u0, t); t = fmaf(w1[5], u1, t);
    float *h1 = xbuf + parity * 32;
    h1[lane] = recip_core(t + b1);
    __syncwarp();
    // layer 2
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < 8; k4++) {
      const float4 hv = reinterpret_cast<const float4 *>(h1)[k4];
      a0 = fmaf(w2[4 * k4 + 0], hv.x, a0); a1 = fmaf(w2[4 * k4 + 1], hv.y, a1);
      a2 = fmaf(w2[4 * k4 + 2], hv.z, a2); a3 = fmaf(w2[4 * k4 + 3], hv.w, a3);
    }
    float *h2 = xbuf + 64 + parity * 32;
    h2[lane] = recip_core(((a0 + a1) + (a2 + a3)) + b2);
    __syncwarp();
    // layer 3: every lane sums all 32 products of output (lane & 3); lanes 0..3 publish the four outputs
    a0 = 0.0f; a1 = 0.0f; a2 = 0.0f; a3 = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < 8; k4++) {
      const float4 gv = reinterpret_cast<const float4 *>(h2)[k4];
      a0 = fmaf(w3[4 * k4 + 0], gv.x, a0); a1 = fmaf(w3[4 * k4 + 1], gv.y, a1);
      a2 = fmaf(w3[4 * k4 + 2], gv.z, a2); a3 = fmaf(w3[4 * k4 + 3], gv.w, a3);
    }
    const float part = ((a0 + a1) + (a2 + a3)) + b3;
    o0 = __shfl_sync(full, part, 0); o1 = __shfl_sync(full, part, 1);
    o2 = __shfl_sync(full, part, 2); o3 = __shfl_sync(full, part, 3);
  }
};

}  // namespace mppi
