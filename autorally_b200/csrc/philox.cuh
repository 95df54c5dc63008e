// philox.cuh -- the control-noise stream: Philox4x32-10 + Box-Muller, shared by the stand-alone sampler kernel
// (weighting.cuh) and the rollout kernels that generate their noise in place (RolloutParams::fused_noise).
//
// Replaces curandGenerateNormal (PI/mppi_controller.cu:330-331,612).  Stream definition: counter = (q, r, call, b) with q
// the timestep pair (t = 2q, 2q+1), r the GLOBAL rollout index, call the compute-call counter and b the GLOBAL controller
// index (mppi_config.controller_begin + local controller); key = seed.  The 4 outputs become eps[r][2q..2q+1][0..1] by
// Box-Muller.  Both users call the same inline functions below with the same counters, so a rollout kernel that
// generates its own noise sees bit-identical values to one that reads the sampler kernel's buffer
// (tests/test_parity_gpu.py::test_fused_noise_is_bitwise_the_sampler_kernel).
#pragma once
#include <stdint.h>

namespace mppi {

// Philox4x32-10 (Salmon et al., SC'11), checked against the Random123 known-answer vectors (tests/test_oracle_golden.py).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float2 box_muller(uint32_t xa, uint32_t xb) {
  // radius = sqrt(-2 ln((xa + .5) 2^-32)); angle = 2 pi (xb + .5) 2^-32 - pi
  const float ua = fmaf((float)xa, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float ang = fmaf((float)xb, 1.4629180792671596e-09f, -3.14159265358979f + 7.3145903963357981e-10f);
  // MUFU.SQRT: the IEEE sqrtf sequence (Newton step + slow-path call) was a seventh of the sampler's instructions.  ua can
  // round to exactly 1 (xa >= 2^32 - 128): the radicand is clamped so that an approximate logarithm that came out a hair
  // positive could never become a NaN sample.
  float rad;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(fmaxf(-2.0f * __logf(ua), 0.0f)));
  float sn, cs;
  __sincosf(ang, &sn, &cs);
  return make_float2(__fmul_rn(rad, cs), __fmul_rn(rad, sn));
}

// eps[r][2q][0], eps[r][2q][1], eps[r][2q+1][0], eps[r][2q+1][1]
__device__ __forceinline__ float4 philox_normal4(uint32_t q, uint32_t r_global, uint32_t call, uint32_t b_global, uint32_t seed_lo,
                                                 uint32_t seed_hi) {
  uint32_t x[4];
  philox4x32_10(q, r_global, call, b_global, seed_lo, seed_hi, x);
  const float2 z0 = box_muller(x[0], x[1]), z1 = box_muller(x[2], x[3]);
  return make_float4(z0.x, z0.y, z1.x, z1.y);
}

// Per-rollout noise source of a rollout kernel that walks the timesteps in order: one Philox call per timestep pair.
struct FusedNoise {
  uint32_t r_global, b_global, call, seed_lo, seed_hi;
  float2 held;  // eps of the odd timestep of the last pair
  __device__ __forceinline__ float2 step(int i) {
    if ((i & 1) == 0) {
      const float4 z = philox_normal4((uint32_t)(i >> 1), r_global, call, b_global, seed_lo, seed_hi);
      held = make_float2(z.z, z.w);
      return make_float2(z.x, z.y);
    }
    return held;
  }
};

}  // namespace mppi
