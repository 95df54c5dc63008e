// latbench.cu -- dependent-chain latencies of the instructions on the latency path, measured with clock64() by ONE
// warp on an otherwise idle SM (run under gpurun).  These numbers drive the design of rollout_half.cu / warp_mlp.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/latbench tools/latbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 256

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float *out, long long *cyc, float a, float b) {
  __shared__ float sm[64];
  const int lane = threadIdx.x;
  float x = a + lane * 1e-3f, y = b;
  unsigned long long x2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(x), "f"(y));
  unsigned long long a2, b2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
  double d = x;
  sm[lane] = x; sm[32 + lane] = y;
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) {
    if (MODE == 0) x = fmaf(x, a, b);                                                     // FFMA
    if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x2) : "l"(a2), "l"(b2));  // FFMA2
    if (MODE == 2) x = ex2a(x);                                                           // MUFU.EX2
    if (MODE == 3) x = rcpa(x);                                                           // MUFU.RCP
    if (MODE == 4) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);                       // SHFL.IDX
    if (MODE == 5) x = __shfl_xor_sync(0xffffffffu, x, 4);                                 // SHFL.BFLY
    if (MODE == 6) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; __syncwarp(); }    // STS -> sync -> LDS -> sync
    if (MODE == 7) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; }                  // STS -> sync -> LDS (double buffered in real code)
    if (MODE == 8) x = fmaf(-2.0f, rcpa(ex2a(x * 2.885390f) + 1.0f), 1.0f);               // tanh_fast
    if (MODE == 9) x = (float)((double)x + (double)(y - x) * 0.25);                         // running-mean update (F2F, DFMA, F2F)
    if (MODE == 10) d = fma(d, 0.999, 0.001);                                             // DFMA
    if (MODE == 11) { float4 v = reinterpret_cast<float4 *>(sm)[(lane + i) & 7]; x = fmaf(v.x, a, x); }  // LDS.128 feeding an FMA (address independent)
    if (MODE == 12) { int idx = __float_as_int(x) & 31; x = sm[idx]; }                     // LDS with dependent address (pure LDS latency)
    if (MODE == 13) x = x + y;                                                            // FADD
  }
  long long t1 = clock64();
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x2));
  out[lane] = x + lo + hi + (float)d;
  if (lane == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char *name, float *out, long long *cyc) {
  long long h = 0;
  for (int rep = 0; rep < 3; rep++) {
    k<MODE><<<1, 32>>>(out, cyc, 0.999f, 0.25f);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  }
  printf("%-58s %7.1f cycles per dependent op\n", name, (double)h / N);
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  run<0>("FFMA", out, cyc);
  run<1>("FFMA2 (fma.rn.f32x2)", out, cyc);
  run<13>("FADD", out, cyc);
  run<2>("MUFU.EX2", out, cyc);
  run<3>("MUFU.RCP", out, cyc);
  run<8>("tanh_fast (FMUL, EX2, FADD, RCP, FFMA)", out, cyc);
  run<4>("SHFL.IDX", out, cyc);
  run<5>("SHFL.BFLY", out, cyc);
  run<12>("LDS (dependent address)", out, cyc);
  run<7>("STS -> __syncwarp -> LDS", out, cyc);
  run<6>("STS -> __syncwarp -> LDS -> __syncwarp", out, cyc);
  run<9>("running mean update (FSUB, F2F.F64, DFMA, F2F.F32)", out, cyc);
  run<10>("DFMA", out, cyc);
  run<11>("LDS.128 + FFMA (independent address)", out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
