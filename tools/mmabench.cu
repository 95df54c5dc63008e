// mmabench.cu -- throughput of the legacy tensor path (mma.sync) on sm_100a for the shapes a tensor-core MLP variant
// would use: m16n8k8 TF32 and m16n8k16 BF16, independent accumulator chains, 4..16 warps per SM sub-partition group.
// Evidence for the DESIGN.md decision on an mma variant of the 32x32 layer (run under gpurun).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mmabench tools/mmabench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(256) k_tf32(float *out, int iters) {
  float c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_bf16(float *out, int iters) {
  float c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  float *d; cudaMalloc(&d, 64);
  const int it = 4096;
  for (int bps = 1; bps <= 4; bps *= 2) {
    const int nb = sms * bps;
    const double warps = 8.0 * nb;
    float ms = time_ms([&] { k_tf32<8><<<nb, 256>>>(d, it); });
    const double n_tf = warps * it * 8;
    printf("mma.sync m16n8k8 tf32, 8 chains, %2d warps/SM : %7.1f TFLOP/s  (%.2f mma / clk / SM at %.0f MHz)\n", 8 * bps,
           n_tf * 2048 / ms / 1e9, n_tf / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1e3);
    ms = time_ms([&] { k_bf16<8><<<nb, 256>>>(d, it); });
    printf("mma.sync m16n8k16 bf16, 8 chains, %2d warps/SM: %7.1f TFLOP/s  (%.2f mma / clk / SM)\n", 8 * bps,
           n_tf * 4096 / ms / 1e9, n_tf / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
  }
  float ms = time_ms([&] { k_tf32<1><<<sms, 32>>>(d, it); });
  printf("mma.sync m16n8k8 tf32 dependent chain latency: %.1f cycles\n", ms * 1e-3 * prop.clockRate * 1e3 / it);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
