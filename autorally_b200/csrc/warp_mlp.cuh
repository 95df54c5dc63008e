// warp_mlp.cuh -- the 6-32-32-4 tanh MLP of NeuralNetModel<7,2,3,6,32,32,4> evaluated by ONE WARP for ONE input
// vector with a short dependent chain: the nominal trajectory of finalize_kernel (weighting.cuh).
//
// Lane l owns hidden neuron l of both hidden layers and, for l < 16, output (l & 3) of the last layer over the quarter
// (l >> 2) of its inputs, with its weight slices in registers for the whole trajectory (46 registers).  Activations cross lanes through a double-buffered 128-byte
// shared-memory slot per layer (one __syncwarp per exchange).  Layers 2 and 3 accumulate their 32 products in four
// interleaved partial sums (k mod 4), so the FMA chain is 8 deep instead of the reference's 32
// (PI/neural_net_model.cu:388-399 sums k ascending): the result differs from the sequential order in the last bits
// (~1e-7 relative), far inside the 1e-4 parity tolerance; the throughput kernels (dynamics.cuh) keep the sequential order.
#pragma once
#include "device_common.cuh"

namespace mppi {

struct WarpMlp32 {
  static constexpr int kW1 = 0, kB1 = 192, kW2 = 224, kB2 = 1248, kW3 = 1280, kB3 = 1408;  // packed transposed layout
  float w1[6], w2[32], w3[8];
  float b1, b2, b3;

  // theta_t: [W1t 6x32 | b1 32 | W2t 32x32 | b2 32 | W3t 32x4 | b3 4] (global or shared memory) in its FOLDED form (fold_nn32,
  // mppi_b200.cu: tanh scale in layers 1 and 2, the map 1 - 2 r in layers 2 and 3), so the activations that cross lanes are
  // r = 1 / (2^x + 1) and each activation is three dependent instructions; loads are coalesced
  __device__ __forceinline__ void load(const float *__restrict__ theta_t, int lane) {
#pragma unroll
    for (int k = 0; k < 6; k++) w1[k] = theta_t[kW1 + k * 32 + lane];
#pragma unroll
    for (int k = 0; k < 32; k++) w2[k] = theta_t[kW2 + k * 32 + lane];
    // layer 3 on lanes 0..15: output (lane & 3) over the quarter (lane >> 2) of k; the bias opens the first quarter's sum
    const int jo = lane & 3, q = (lane >> 2) & 3;
#pragma unroll
    for (int m = 0; m < 8; m++) w3[m] = theta_t[kW3 + (8 * q + m) * 4 + jo];
    b1 = theta_t[kB1 + lane];
    b2 = theta_t[kB2 + lane];
    b3 = q == 0 ? theta_t[kB3 + jo] : 0.0f;
  }

  // xbuf: 144 floats of shared memory owned by this warp (h1[2][32], h2[2][32], 16 partial sums); parity = step & 1.
  // Inputs are replicated in all lanes; the four outputs come back replicated in all lanes.
  __device__ __forceinline__ void forward(float roll, float vx, float vy, float wz, float u0, float u1, float *xbuf,
                                          int parity, int lane, float &o0, float &o1, float &o2, float &o3) const {
    // layer 1: two interleaved partial sums, the bias opens one of them, the controls (which arrive through shuffles) last
    float t = fmaf(w1[0], roll, b1), tb = w1[1] * vx;
    t = fmaf(w1[2], vy, t); tb = fmaf(w1[3], wz, tb); t = fmaf(w1[4], u0, t); tb = fmaf(w1[5], u1, tb);
    float *h1 = xbuf + parity * 32;
    h1[lane] = recip_core(t + tb);
    __syncwarp();
    // layer 2
    float a0 = b2, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;  // the bias opens the first partial sum
#pragma unroll
    for (int k4 = 0; k4 < 8; k4++) {
      const float4 hv = reinterpret_cast<const float4 *>(h1)[k4];
      a0 = fmaf(w2[4 * k4 + 0], hv.x, a0); a1 = fmaf(w2[4 * k4 + 1], hv.y, a1);
      a2 = fmaf(w2[4 * k4 + 2], hv.z, a2); a3 = fmaf(w2[4 * k4 + 3], hv.w, a3);
    }
    float *h2 = xbuf + 64 + parity * 32;
    h2[lane] = recip_core((a0 + a1) + (a2 + a3));
    __syncwarp();
    // layer 3: lanes 0..15 sum the 8 products of (output, quarter of k); the 16 partial sums meet in shared memory (one store,
    // four broadcast loads, two levels of adds: fewer instructions and a shorter chain than 32 products per lane + 4 shuffles)
    float *pb = xbuf + 128;
    if (lane < 16) {
      const float4 g0 = reinterpret_cast<const float4 *>(h2 + 8 * (lane >> 2))[0], g1 = reinterpret_cast<const float4 *>(h2 + 8 * (lane >> 2))[1];
      float c0 = fmaf(w3[0], g0.x, b3), c1 = w3[1] * g0.y;
      c0 = fmaf(w3[2], g0.z, c0); c1 = fmaf(w3[3], g0.w, c1);
      c0 = fmaf(w3[4], g1.x, c0); c1 = fmaf(w3[5], g1.y, c1);
      c0 = fmaf(w3[6], g1.z, c0); c1 = fmaf(w3[7], g1.w, c1);
      pb[lane] = c0 + c1;  // pb[4 q + jo]
    }
    __syncwarp();
    const float4 q0 = reinterpret_cast<const float4 *>(pb)[0], q1 = reinterpret_cast<const float4 *>(pb)[1];
    const float4 q2 = reinterpret_cast<const float4 *>(pb)[2], q3 = reinterpret_cast<const float4 *>(pb)[3];
    o0 = (q0.x + q1.x) + (q2.x + q3.x); o1 = (q0.y + q1.y) + (q2.y + q3.y);
    o2 = (q0.z + q1.z) + (q2.z + q3.z); o3 = (q0.w + q1.w) + (q2.w + q3.w);
  }
};

}  // namespace mppi
