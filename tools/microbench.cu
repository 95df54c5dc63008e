// microbench.cu -- design-question microbenchmarks for the rollout kernel (run under gpurun):
//   FFMA / FFMA2 issue peaks, the 32x32 layer contraction with weights from shared memory (broadcast
//   LDS.128) for R = 1, 2, 4 rollouts per thread and from the constant bank, and the MUFU-based tanh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../autorally_b200/csrc/dynamics.cuh"

using namespace mppi;

__constant__ float c_w[32 * 32 + 32];

template <int CH>
__global__ void __launch_bounds__(256) k_ffma(float *out, float a, float b, int iters) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float *out, float a, float b, int iters) {
  unsigned long long acc[8], av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int i = 0; i < 8; i++) { float x = (float)(threadIdx.x + i); asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x)); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
  if (s == 123.456f) out[0] = s;
}

// 32x32 layer, weights from smem (transposed, float4 broadcast), R rollouts per thread
template <int R, bool ACT>
__global__ void __launch_bounds__(128) k_layer_smem(float *out, const float *w, int iters) {
  __shared__ float4 sw4[(32 * 32 + 32) / 4];
  for (int i = threadIdx.x; i < (32 * 32 + 32) / 4; i += 128) sw4[i] = reinterpret_cast<const float4 *>(w)[i];
  __syncthreads();
  const float *sw = reinterpret_cast<const float *>(sw4);
  float a[32][R], o[32][R];
#pragma unroll
  for (int k = 0; k < 32; k++)
#pragma unroll
    for (int r = 0; r < R; r++) a[k][r] = 0.001f * (threadIdx.x + k + r);
  for (int it = 0; it < iters; it++) {
    asm volatile("" ::: "memory");  // keep the (loop-invariant) weight loads inside the loop
    dense_layer<R, 32, 32, ACT>(sw, a, o);
#pragma unroll
    for (int k = 0; k < 32; k++)
#pragma unroll
      for (int r = 0; r < R; r++) a[k][r] = o[k][r];
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 32; k++)
#pragma unroll
    for (int r = 0; r < R; r++) s += a[k][r];
  if (s == 123.456f) out[0] = s;
}

// same contraction, weights as constant-bank operands
template <int R>
__global__ void __launch_bounds__(128) k_layer_const(float *out, int iters) {
  float a[32][R], o[32][R];
#pragma unroll
  for (int k = 0; k < 32; k++)
#pragma unroll
    for (int r = 0; r < R; r++) a[k][r] = 0.001f * (threadIdx.x + k + r);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 32; j++)
#pragma unroll
      for (int r = 0; r < R; r++) o[j][r] = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k++)
#pragma unroll
      for (int j = 0; j < 32; j++)
#pragma unroll
        for (int r = 0; r < R; r++) o[j][r] = fmaf(c_w[k * 32 + j], a[k][r], o[j][r]);
#pragma unroll
    for (int k = 0; k < 32; k++)
#pragma unroll
      for (int r = 0; r < R; r++) a[k][r] = o[k][r] + c_w[1024 + k];
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 32; k++)
#pragma unroll
    for (int r = 0; r < R; r++) s += a[k][r];
  if (s == 123.456f) out[0] = s;
}

template <int WSRC, bool ACT>
__global__ void __launch_bounds__(128) k_layer_p2(float *out, const float *w, int iters) {
  __shared__ float4 sw4[(32 * 32 + 32) / 4];
  for (int i = threadIdx.x; i < (32 * 32 + 32) / 4; i += 128) sw4[i] = reinterpret_cast<const float4 *>(w)[i];
  __syncthreads();
  const float *sw = reinterpret_cast<const float *>(sw4);
  float2 a[32], o[32];
#pragma unroll
  for (int k = 0; k < 32; k++) a[k] = make_float2(0.001f * (threadIdx.x + k), 0.002f * (threadIdx.x + k));
  for (int it = 0; it < iters; it++) {
    asm volatile("" ::: "memory");
    dense_layer_p2<32, 32, ACT, WSRC>(sw, 0, a, o);
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = o[k];
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 32; k++) s = __fadd2_rn(s, a[k]);
  if (s.x == 123.456f) out[0] = s.y;
}

__global__ void __launch_bounds__(256) k_tanh(float *out, float a, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = a * (threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = tanh_fast(x[i] + a);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  if (s == 123.456f) out[0] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  float *d, *w;
  cudaMalloc(&d, 64);
  cudaMalloc(&w, (32 * 32 + 32) * 4);
  float hw[32 * 32 + 32];
  for (int i = 0; i < 32 * 32 + 32; i++) hw[i] = 0.03f * ((i * 7919) % 61 - 30) / 30.f;
  cudaMemcpy(w, hw, sizeof(hw), cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(c_w, hw, sizeof(hw));
  const int blocks = sms * 8;
  {
    const int it = 8192;
    float ms = time_ms([&] { k_ffma<16><<<blocks, 256>>>(d, 0.999f, 0.001f, it); });
    printf("FFMA  16 chains      : %7.2f TFLOP/s\n", 2.0 * 16 * it * 256.0 * blocks / ms / 1e9);
    ms = time_ms([&] { k_ffma<8><<<blocks, 256>>>(d, 0.999f, 0.001f, it); });
    printf("FFMA   8 chains      : %7.2f TFLOP/s\n", 2.0 * 8 * it * 256.0 * blocks / ms / 1e9);
    ms = time_ms([&] { k_ffma2<<<blocks, 256>>>(d, 0.999f, 0.001f, it); });
    printf("FFMA2  8 chains x2   : %7.2f TFLOP/s\n", 2.0 * 16 * it * 256.0 * blocks / ms / 1e9);
    ms = time_ms([&] { k_tanh<<<blocks, 256>>>(d, 0.001f, it); });
    printf("tanh_fast            : %7.2f Gtanh/s  (%.2f per clk per SM at %.0f MHz nominal)\n", 8.0 * it * 256.0 * blocks / ms / 1e6,
           8.0 * it * 256.0 * blocks / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1e3);
  }
  {
    const int it = 2000;
    const double fl = 2.0 * 1024 * it;
    for (int bps = 2; bps <= 8; bps += 2) {
      const int nb = sms * bps;
      float ms = time_ms([&] { k_layer_smem<1, false><<<nb, 128>>>(d, w, it); });
      printf("layer32x32 smem R=1 (%d CTA/SM of 128): %7.2f TFLOP/s\n", bps, fl * 1 * 128.0 * nb / ms / 1e9);
      ms = time_ms([&] { k_layer_smem<2, false><<<nb, 128>>>(d, w, it); });
      printf("layer32x32 smem R=2 (%d CTA/SM of 128): %7.2f TFLOP/s\n", bps, fl * 2 * 128.0 * nb / ms / 1e9);
      ms = time_ms([&] { k_layer_const<1><<<nb, 128>>>(d, it); });
      printf("layer32x32 const R=1 (%d CTA/SM of 128): %7.2f TFLOP/s\n", bps, fl * 1 * 128.0 * nb / ms / 1e9);
      ms = time_ms([&] { k_layer_const<2><<<nb, 128>>>(d, it); });
      printf("layer32x32 const R=2 (%d CTA/SM of 128): %7.2f TFLOP/s\n", bps, fl * 2 * 128.0 * nb / ms / 1e9);
    }
    cudaMemcpyToSymbol(c_theta, hw, sizeof(hw));
    for (int bps = 3; bps <= 6; bps += 3) {
      const int nbp = sms * bps;
      float msp = time_ms([&] { k_layer_p2<0, false><<<nbp, 128>>>(d, w, it); });
      printf("layer32x32 FFMA2 smem  (%d CTA/SM of 128): %7.2f TFLOP/s\n", bps, fl * 2 * 128.0 * nbp / msp / 1e9);
      msp = time_ms([&] { k_layer_p2<1, false><<<nbp, 128>>>(d, w, it); });
      printf("layer32x32 FFMA2 const (%d CTA/SM of 128): %7.2f TFLOP/s\n", bps, fl * 2 * 128.0 * nbp / msp / 1e9);
      msp = time_ms([&] { k_layer_p2<0, true><<<nbp, 128>>>(d, w, it); });
      printf("layer32x32 FFMA2 smem +tanh2 (%d CTA/SM): %7.2f TFLOP/s (FMA flops only)\n", bps, fl * 2 * 128.0 * nbp / msp / 1e9);
      msp = time_ms([&] { k_layer_p2<1, true><<<nbp, 128>>>(d, w, it); });
      printf("layer32x32 FFMA2 const+tanh2 (%d CTA/SM): %7.2f TFLOP/s (FMA flops only)\n", bps, fl * 2 * 128.0 * nbp / msp / 1e9);
    }
    const int nb = sms * 4;
    float ms = time_ms([&] { k_layer_smem<4, false><<<nb, 128>>>(d, w, it); });
    printf("layer32x32 smem R=4 (4 CTA/SM of 128): %7.2f TFLOP/s\n", fl * 4 * 128.0 * nb / ms / 1e9);
    ms = time_ms([&] { k_layer_smem<1, true><<<nb, 128>>>(d, w, it); });
    printf("layer32x32+tanh smem R=1 (4 CTA/SM)  : %7.2f TFLOP/s (FMA flops only)\n", fl * 1 * 128.0 * nb / ms / 1e9);
    ms = time_ms([&] { k_layer_smem<2, true><<<nb, 128>>>(d, w, it); });
    printf("layer32x32+tanh smem R=2 (4 CTA/SM)  : %7.2f TFLOP/s (FMA flops only)\n", fl * 2 * 128.0 * nb / ms / 1e9);
  }
  return 0;
}
