// rollout_tc.cu -- the filled-GPU rollout kernel with the MLP contraction on the 5th-generation tensor cores.
//
// Replaces rolloutKernel (PI/mppi_controller.cu:72-184) + NeuralNetModel<7,2,3,6,32,32,4>::computeDynamics
// (PI/neural_net_model.cu:357-410) for large rollout counts.  One thread owns one rollout (bookkeeping, clamp, costs,
// kinematics, Euler step: exactly the code of rollout.cuh); the three layer contractions of a 128-rollout tile are
// tcgen05.mma instructions (M = 128 rollouts, N = 32 / 32 / 16 neurons):
//
//   * A (the activations) lives in TENSOR MEMORY: thread r owns TMEM lane r, writes its own row with tcgen05.st and the
//     MMA reads it from there -- no shared-memory round trip, no swizzled layouts, no cross-thread traffic at all.
//   * B (the weights) is staged once per CTA in shared memory in the canonical K-major no-swizzle layout.
//   * D (the pre-activations) accumulates in TMEM in FP32 and comes back with tcgen05.ld, one row per thread.
//
// FP32 accuracy on FP16 tensor cores: a single FP16 / TF32 pass (11-bit significands) misses the 1e-4 parity bar by an
// order of magnitude after 100 recurrent steps, so every operand is split into two FP16 halves,
// x = x_hi + x_lo (|x_lo| <= 2^-11 |x|; the split is exact in FP32), and each layer is three accumulating passes
//   D = A_hi B_hi + A_lo B_hi + A_hi B_lo          (dropped term A_lo B_lo ~ 2^-22 relative),
// issued as K = 16 chunks.  FP16 products are exact in the FP32 accumulator.
//
// The tanh epilogue is what remains on the CUDA cores, and the MUFU unit (16 results/clk/SM) is its bottleneck, so it
// is written to need as few MUFU results as possible.  The hidden layers' weights are pre-multiplied by 2 log2(e), and
// the bias is folded into the exponential: 2^(x' + b') = 2^x' * 2^b' with 2^b' a kernel constant, so
//   tanh(x + b) = 1 - 2 r,  r = 1 / (2^x' 2^b' + 1):   MUFU.EX2, one FFMA for "* 2^b' + 1", a reciprocal;
// the affine map 1 - 2r is folded into the NEXT layer (W y + b = (b + W 1) - 2 W r: the next layer's B matrix is -2 W,
// its bias b + rowsum(W)), so the activations that travel through tensor memory are the r themselves;
// and the reciprocals of FOUR neurons share ONE MUFU.RCP (1/(d0 d1 d2 d3) and five packed multiplies; every d is
// clamped to 2^30, where tanh is 1 to the last bit, so the product cannot overflow): 5 MUFU per 4 neurons, not 8.
//
// Tensor work per tile and timestep: 2 + 6 + 6 MMAs = 16 + 96 + 48 = 160 tensor-pipe cycles for 128 rollout-steps
// (1.25 cycles per rollout-step, B300_MICROARCH.md: M=128 costs N/2 cycles per K chunk); the FFMA2 kernel spends 10.5
// FMA-pipe cycles per rollout-step on the same contraction.
//
// Synchronisation per layer: tcgen05.st -> wait::st -> fence::before_thread_sync -> bar.sync -> one elected thread issues
// the MMAs and a tcgen05.commit onto an mbarrier -> everybody waits on the mbarrier -> fence::after_thread_sync ->
// tcgen05.ld.  Eight CTAs (tiles) per SM, 64 TMEM columns and 64 registers each, hide that round trip behind each other's
// epilogues.
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include "rollout.cuh"
#include "rollout_launch.h"

namespace mppi {
namespace tc {

#ifndef TC_EXP
#define TC_EXP 0  // timing experiments only (tools/exp_tc_parts.sh); 0 = the product kernel
#endif
#ifndef TC_MINCTAS
#define TC_MINCTAS 8  // resident tiles per SM of the 32-wide kernel (registers per thread = 65536 / 128 / TC_MINCTAS)
#endif
#ifndef TC_PAD
#define TC_PAD 0      // 1: pad the dynamic shared memory so that at most MIN_CTAS CTAs are resident (experiment)
#endif

constexpr int TILE = 128;         // rollouts per CTA = TMEM lanes
constexpr float TANH_SCALE = 2.88539008177792681472f;  // 2 log2(e)

// Geometry of one instantiation: 6 -> HID x NHID (tanh) -> 4.  NeuralNetModel<7,2,3,6,32,32,4> is <32, 2>, the
// wider_deeper network 6-64-64-64-64-4 the fork ships (SRC/params/models/wider_deeper_network_08_20_2020.npz) <64, 4>.
template <int HID, int NHID, int TPC_ = 1, int CS_ = 1>
struct Geo {
  static_assert(HID == 32 || HID == 64, "hidden width 32 or 64");
  static_assert(NHID >= 2, "at least two hidden layers");
  static constexpr int NCH = HID / 16;                    // K = 16 chunks per hidden layer = 16-column epilogue chunks
  static constexpr int COL_D = 0;                         // accumulator, HID columns
  static constexpr int COL_A = HID;                       // activations: per chunk [hi(16 neurons) | lo(16 neurons)], 8 + 8 columns
  static constexpr int TMEM_COLS = 2 * HID;               // per tile, power of two: 64 or 128
  // Tiles (128-thread groups) per CTA.  The 56 KB of FP16 weights of the 64-wide network allow only 3 one-tile CTAs per SM;
  // two tiles sharing one copy of the weights give 2 CTAs = 4 tiles per SM, all of the tensor memory (1 M rollouts: 14.4 ->
  // 13.6 ms, 65536: 1.32 -> 1.06 ms).  Small problems keep one tile per CTA so that the tiles spread over more SMs (1920
  // rollouts = 15 tiles: 0.54 ms on 15 SMs, 0.73 ms on 8).
  static constexpr int TPC = TPC_;
  // Column slices: threads per rollout in the tanh epilogues.  One tile on an SM (a controller-sized problem: 1920 rollouts
  // = 15 tiles) is a latency chain in which a single warp per scheduler works through HID tanh + FP16 splits per layer; with
  // CS slices, warp 4 s + w (the same tensor-memory lane quarter as warp w) takes the 16-column chunks c = s (mod CS) of its
  // 32 rollouts.  Slice 0 owns the rollouts (controls, costs, kinematics, the state); the others only run epilogues.
  static constexpr int CS = CS_;
  static_assert(CS == 1 || (TPC == 1 && (HID / 16) % CS == 0), "column slices: one tile per CTA, chunks divide evenly");
  static constexpr int THREADS = TILE * TPC * CS;
  // shared-memory B matrices (FP16, canonical K-major no-swizzle: 8 rows x 16 bytes core matrices)
  static constexpr int SZ_B1 = HID * 16 * 2;              // N = HID, K = 16
  static constexpr int SZ_BH = HID * HID * 2;             // N = HID, K = HID
  static constexpr int SZ_BL = 16 * HID * 2;              // N = 16 (4 real output rows), K = HID
  static constexpr int OFF_B1A = 0;                       // W1_hi twice (against [a_hi | a_lo])
  static constexpr int OFF_B1B = SZ_B1;                   // W1_lo, 0
  static constexpr int OFF_BH = 2 * SZ_B1;                // hidden layer h = 1 .. NHID-1: hi at OFF_BH + (h-1) 2 SZ_BH, lo after it
  static constexpr int OFF_BL = OFF_BH + (NHID - 1) * 2 * SZ_BH;  // last layer: hi, lo
  static constexpr int B_BYTES = OFF_BL + 2 * SZ_BL;
  // HID = 32: the pad keeps residency at 8 CTAs per SM (8 x 64 TMEM columns): a ninth CTA would only spin in
  // tcgen05.alloc.  HID = 64: 56 KB of weights, 3 CTAs per SM (4 x 128 columns would fit, shared memory does not).
  // Resident tiles per SM: 8 x 64 TMEM columns at HID = 32 (the 64 registers per thread this allows fill the register
  // file exactly, so a ninth CTA, which could only spin in tcgen05.alloc, never becomes resident); 3 at HID = 64 (56 KB of
  // weights each).  No shared-memory padding: padding the allocation to fence off extra CTAs costs L1 / texture cache
  // (unified with shared memory) and was measured 8 % slower at 1 M rollouts (profiles/exp_tc_cfg_r01.txt).  The launcher
  // falls back to padding only if a build ever uses so few registers that one more CTA would fit.
  static constexpr int MIN_CTAS = CS > 1 ? 1 : HID == 32 ? TC_MINCTAS : (TPC == 2 ? 2 : 3);
  static constexpr int PAD_BYTES = (227 / MIN_CTAS - 2) * 1024;
  static constexpr int SMEM_BYTES = (TC_PAD && B_BYTES < PAD_BYTES) ? PAD_BYTES : B_BYTES;
  // packed transposed parameters: per layer Wt[k][j] then b[j]
  static constexpr int TH_W1 = 0, TH_B1 = 6 * HID;
  __host__ __device__ static constexpr int th_w(int h) { return 7 * HID + (h - 1) * (HID * HID + HID); }       // hidden layer h >= 1
  __host__ __device__ static constexpr int th_b(int h) { return th_w(h) + HID * HID; }
  static constexpr int TH_WL = 7 * HID + (NHID - 1) * (HID * HID + HID), TH_BL = TH_WL + HID * 4;
  static constexpr int NPARAMS = TH_BL + 4;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (n, k) of an N x K FP16 matrix: core matrices ordered [k / 8][n / 8]
__device__ __forceinline__ int b_off(int N, int n, int k) { return ((k >> 3) * (N >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2; }

// SM100 shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 [0,14), LBO >> 4 [16,30) = byte
// distance between the two 16-byte K halves, SBO >> 4 [32,46) = byte distance between 8-row groups, version 1 [46,48),
// layout type 0 = no swizzle [61,64).  `base_lo` = (shared address of the B block) >> 4; everything else is a constant.
__device__ __forceinline__ uint64_t chunk_desc(uint32_t base_lo, int off, int N, int chunk) {
  const uint32_t lbo = (uint32_t)(N >> 3) * 128u, sbo = 128u;
  const uint32_t lo = base_lo + (((uint32_t)off + (uint32_t)chunk * 2u * lbo) >> 4) + ((lbo >> 4) << 16);  // one K = 16 chunk = two core-matrix columns
  const uint32_t hi = (sbo >> 4) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

// instruction descriptor: D = F32 [4,6), A = B = F16 (0), both K-major, N >> 3 [17,23), M >> 4 [24,29)
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(id), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a converged warp (elect.sync): the form of single-thread predicate the compiler recognises -- under a plain
// `tid == 0` it wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop over the active lanes (~65 cycles per MMA issued)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bounded: a mis-programmed MMA must end as wrong numbers (caught by the parity tests), never as a hung GPU.
// SPIN: poll with test_wait (one tile per SM: nobody else wants the issue slots and the wake-up from a suspended try_wait is
// part of the latency chain); otherwise try_wait with a suspend-time hint, which sleeps in hardware.
template <bool SPIN>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (TC_EXP == 3 || TC_EXP == 4) return true;
  uint32_t done = 0;
  if (SPIN) {
#pragma unroll 1
    for (int spin = 0; spin < (1 << 26); spin++) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      if (done) return true;
    }
    return false;
  }
#pragma unroll 1
  for (int spin = 0; spin < (1 << 18); spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)  // suspend-time hint (ns): sleep in hardware instead of spinning on issue slots
        : "memory");
    if (done) return true;
  }
  return false;
}

__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}

// x0, x1 -> packed FP16 pair of the high parts (x0 in the low half = the smaller k) and of the residuals x - hi.
// The residual is one mixed-precision FMA per element (fma.rn.f32.f16: hi * (-1) + x, SASS FHFMA; exact, because x - hi is
// representable): 4 instructions per pair instead of 5 with a conversion back to FP32 and a packed subtract.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  float l0, l1;
  asm("{\n\t.reg .b16 h0, h1, m1;\n\tmov.b32 {h0, h1}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
      "fma.rn.f32.f16 %0, h0, m1, %3;\n\tfma.rn.f32.f16 %1, h1, m1, %4;\n\t}\n"
      : "=f"(l0), "=f"(l1)
      : "r"(hi), "f"(x0), "f"(x1));
  const __half2 l = __floats2half2_rn(l0, l1);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// Epilogue constants, in the kernel parameter block (constant bank): e^(2 (b + rowsum W)) of hidden layers 2 .. NHID
// (eb[1 ..]; the first layer's bias goes through its MMA), b + rowsum W of the output layer.
template <int HID, int NHID>
struct TcEpilogue {
  float eb[NHID][HID];
  float b_last[4];
};

// r = 1 / (2^x eb + 1) for four neurons (tanh(.) = 1 - 2r); x already scaled by 2 log2(e), eb = 2^(scaled bias).
// One MUFU.RCP for all four.
__device__ __forceinline__ void recip4(float x0, float x1, float x2, float x3, float2 eb01, float2 eb23, float2 &r01, float2 &r23) {
  const float2 one = make_float2(1.0f, 1.0f);
#if TC_EXP == 2
  float2 d01 = __ffma2_rn(make_float2(fmaf(x0, x0, 1.0f), fmaf(x1, x1, 1.0f)), eb01, one);
  float2 d23 = __ffma2_rn(make_float2(fmaf(x2, x2, 1.0f), fmaf(x3, x3, 1.0f)), eb23, one);
#else
  float2 d01 = __ffma2_rn(make_float2(ex2_approx(x0), ex2_approx(x1)), eb01, one);
  float2 d23 = __ffma2_rn(make_float2(ex2_approx(x2), ex2_approx(x3)), eb23, one);
#endif
  const float cap = 1073741824.0f;  // 2^30: 1 - 2/d is 1 to the last bit beyond it; keeps d0 d1 d2 d3 finite
  d01.x = fminf(d01.x, cap); d01.y = fminf(d01.y, cap);
  d23.x = fminf(d23.x, cap); d23.y = fminf(d23.y, cap);
  const float2 q = __fmul2_rn(d01, d23);                   // (d0 d2, d1 d3)
  const float r = rcp_approx(__fmul_rn(q.x, q.y));         // 1 / (d0 d1 d2 d3)
  const float2 iq = __fmul2_rn(make_float2(r, r), make_float2(q.y, q.x));  // (1 / (d0 d2), 1 / (d1 d3))
  r01 = __fmul2_rn(iq, d23);                               // (1 / d0, 1 / d1)
  r23 = __fmul2_rn(iq, d01);                               // (1 / d2, 1 / d3)
}

__device__ __forceinline__ void split2v(float2 y, uint32_t &hi, uint32_t &lo) { split2(y.x, y.y, hi, lo); }

// 16 pre-activations -> 8 columns of hi pairs, 8 columns of lo pairs.  BIAS_IN_ACC: the bias arrived through the MMA
// (layer 1 carries it in its K padding), so 2^x needs no factor: d = 2^x + 1 and no constants are fetched.
template <bool BIAS_IN_ACC>
__device__ __forceinline__ void activate16(const float (&v)[16], const float *eb, uint32_t (&out)[16]) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 y01, y23;
    const float2 e01 = BIAS_IN_ACC ? make_float2(1.0f, 1.0f) : make_float2(eb[4 * j], eb[4 * j + 1]);
    const float2 e23 = BIAS_IN_ACC ? make_float2(1.0f, 1.0f) : make_float2(eb[4 * j + 2], eb[4 * j + 3]);
    recip4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3], e01, e23, y01, y23);
    split2v(y01, out[2 * j], out[8 + 2 * j]);
    split2v(y23, out[2 * j + 1], out[8 + 2 * j + 1]);
  }
}

__device__ __forceinline__ void put_split(unsigned char *hi_base, unsigned char *lo_base, int N, int n, int k, float x) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(__fsub_rn(x, __half2float(h)));
  *reinterpret_cast<__half *>(hi_base + b_off(N, n, k)) = h;
  if (lo_base) *reinterpret_cast<__half *>(lo_base + b_off(N, n, k)) = l;
}

// FUSED: the noise of a timestep pair is drawn in place from the Philox stream (philox.cuh) inside the wait for the output
// layer's MMAs instead of being read from `du` (800-byte stride, written by a separate sampler launch).
template <int HID, int NHID, int TPC, bool FUSED, int CS>
__global__ void __launch_bounds__(Geo<HID, NHID, TPC, CS>::THREADS, Geo<HID, NHID, TPC, CS>::MIN_CTAS)
rollout_tc_kernel(const __grid_constant__ RolloutParams p, const __grid_constant__ TcEpilogue<HID, NHID> ep) {
  using G = Geo<HID, NHID, TPC, CS>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mma_bar_sm[G::TPC];
  __shared__ uint32_t tmem_base_slot;
  // a CTA holds TPC independent tiles that share the weights in shared memory; `tid`, `warp` are relative to the tile
  // with column slices the CTA is one tile and `slice` says which chunks of the epilogues this thread takes
  const int cta_tid = threadIdx.x, grp = cta_tid / TILE, tid = cta_tid - grp * TILE, warp = tid >> 5;
  const int tile = CS > 1 ? 0 : grp, slice = CS > 1 ? grp : 0;
  const bool owner = (slice == 0);
  auto tile_sync = [&]() {  // barrier among the 128 threads of this tile (named barrier 1 + tile)
    if (G::TPC == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + tile), "n"(TILE) : "memory");
  };

  // ---- prologue: weights -> FP16 hi / lo B matrices in shared memory (theta_t: per layer Wt[k][j], then b[j]) ----
  for (int i = cta_tid; i < G::B_BYTES / 16; i += G::THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  {
    const float *th = p.theta_t;
    for (int i = cta_tid; i < 6 * HID; i += G::THREADS) {  // layer 1 weights, scaled
      const int k = i / HID, n = i - k * HID;
      const float w = __fmul_rn(th[G::TH_W1 + i], TANH_SCALE);
      put_split(smem + G::OFF_B1A, smem + G::OFF_B1B, HID, n, k, w);
      put_split(smem + G::OFF_B1A, nullptr, HID, n, 8 + k, w);
    }
    for (int n = cta_tid; n < HID; n += G::THREADS)  // b1 rides in the K padding: the operand carries 1.0 at k = 6
      put_split(smem + G::OFF_B1A, smem + G::OFF_B1B, HID, n, 6, __fmul_rn(th[G::TH_B1 + n], TANH_SCALE));
#pragma unroll
    for (int h = 1; h < NHID; h++) {  // hidden layer h acts on r of layer h - 1: -2 W (and the tanh scale)
      unsigned char *hi = smem + G::OFF_BH + (h - 1) * 2 * G::SZ_BH;
      for (int i = cta_tid; i < HID * HID; i += G::THREADS) {
        const int k = i / HID, n = i - k * HID;
        put_split(hi, hi + G::SZ_BH, HID, n, k, __fmul_rn(th[G::th_w(h) + i], -2.0f * TANH_SCALE));
      }
    }
    for (int i = cta_tid; i < HID * 4; i += G::THREADS) {  // output layer acts on r of the last hidden layer: -2 W (linear, no tanh scale)
      const int k = i >> 2, n = i & 3;
      put_split(smem + G::OFF_BL, smem + G::OFF_BL + G::SZ_BL, 16, n, k, -2.0f * th[G::TH_WL + i]);
    }
  }
  const uint32_t bar = smem_u32(&mma_bar_sm[tile]);
  if (owner && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (cta_tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(G::TMEM_COLS * G::TPC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy weight stores -> visible to the tensor core
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_cta = tmem_base_slot;
  const uint32_t tmem = tmem_cta + (uint32_t)(tile * G::TMEM_COLS);  // this tile's columns
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes

  const uint32_t sb = (smem_u32(smem) & 0x3FFFFu) >> 4;
  constexpr uint32_t IDH = idesc(128, HID), ID16 = idesc(128, 16);
  // Programmatic dependent launch: everything above (weights -> FP16 B matrices, tensor-memory allocation) reads model
  // parameters only and overlaps the tail of the previous kernel of the stream (the finalize kernel of the previous step);
  // the inbox, the call counter and the noise buffer are read after the wait.  No early launch_dependents here: the weighting
  // kernel's CTAs would become resident at once (a one-wave rollout grid leaves them room), and 4 x 256 threads per SM polling
  // in griddepcontrol.wait for the whole rollout kernel cost it 11 % (131072 rollouts: 0.543 -> 0.606 ms).
  pdl_wait();

  // ---- rollout bookkeeping (rollout.cuh, R = 1) ----
  const long long total = (long long)p.B * p.n_local;
  const long long g0 = ((long long)blockIdx.x * G::TPC + tile) * TILE + tid;
  const bool valid = owner && g0 < total;
  const long long gc = valid ? g0 : 0;  // idle threads shadow rollout 0 (they must take part in every barrier) and store nothing
  const int ctrl = (int)(gc / p.n_local);
  const int lr0 = (int)(gc - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  const float2 *U = reinterpret_cast<const float2 *>(inbox + INBOX_U);
  float s[S_DIM];
#pragma unroll
  for (int k = 0; k < S_DIM; k++) s[k] = inbox[INBOX_STATE + k];
  float running = 0.0f;
  int crash = 0;
  const int rg = p.r_begin + lr0;
  const bool noise_free = (rg == 0), pure_noise = (rg >= p.pure_noise_from);
  float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)gc * p.T;
  uint32_t phase = 0;
  bool ok = true;

  // The controls of a step do not depend on the state, so they are prepared one step ahead, inside the wait for the last
  // layer of the previous step (PI/mppi_controller.cu:130-155 + enforceConstraints).  The noise row has an 800-byte
  // stride (every fetch is its own DRAM sector) and is requested a whole step before it is used.
  float2 e_next = FUSED ? make_float2(0.0f, 0.0f) : row[0];
  FusedNoise fz;
  fz.r_global = (uint32_t)rg; fz.b_global = (uint32_t)(p.b_begin + ctrl); fz.call = FUSED ? *p.call_ptr : 0u;
  fz.seed_lo = p.seed_lo; fz.seed_hi = p.seed_hi; fz.held = make_float2(0.0f, 0.0f);
  float control_cost;    // control cost of the prepared step (PI/costs.cu:307-313): a function of the controls only
  uint32_t ua_hi, ua_lo;  // FP16 hi / lo pairs of the clamped controls, ready for the layer-1 operand
  auto prepare_controls = [&](int i) {
    const float2 Ui = U[i];
    float2 e;
    if (FUSED) {
      e = fz.step(i);
    } else {
      e = e_next;
      if (i + 1 < p.T) e_next = row[i + 1];
    }
    float du0, du1, u0, u1;
    if (noise_free || i < p.opt_delay) {
      du0 = 0.0f; du1 = 0.0f; u0 = Ui.x; u1 = Ui.y;
    } else if (pure_noise) {
      du0 = __fmul_rn(e.x, p.nu0); du1 = __fmul_rn(e.y, p.nu1); u0 = du0; u1 = du1;
    } else {
      du0 = __fmul_rn(e.x, p.nu0); du1 = __fmul_rn(e.y, p.nu1);
      u0 = __fadd_rn(Ui.x, du0); u1 = __fadd_rn(Ui.y, du1);
    }
    if (valid) row[i] = make_float2(u0, u1);  // un-clamped write-back (:153)
    u0 = u0 < p.lo0 ? p.lo0 : (u0 > p.hi0 ? p.hi0 : u0);  // enforceConstraints
    u1 = u1 < p.lo1 ? p.lo1 : (u1 > p.hi1 ? p.hi1 : u1);
    control_cost = cost_control_part(p.cp, u0, u1, du0, du1, p.nu0, p.nu1);
    split2(u0, u1, ua_hi, ua_lo);
  };
  if (owner) prepare_controls(0);
  float front = 0.0f, back = 0.0f;  // costmap texels under the state of the current step (requested one step ahead as well)

#if TC_EXP == 9  // timing experiment: cycle stamps of step 50 of rollout 0, written over its row of sampled controls
  unsigned stamps[48];
#pragma unroll
  for (int k = 0; k < 48; k++) stamps[k] = 0u;
#define TC_STAMP(k) do { if (i == 50 && cta_tid == 0 && blockIdx.x == 0) stamps[k] = clock(); } while (0)
#else
#define TC_STAMP(k) do { } while (0)
#endif
  for (int i = 0; i < p.T; i++) {
    TC_STAMP(0);
    // ---- layer 1: a = [roll, u_x, u_y, yaw rate, steering, throttle, 1 (bias), 0] as [a_hi | a_lo], one K = 16 chunk ----
    if (owner) {
      uint32_t a[8];
      split2(s[3], s[4], a[0], a[4]);
      split2(s[5], s[6], a[1], a[5]);
      a[2] = ua_hi; a[6] = ua_lo;
      a[3] = 0x00003C00u;  // (1.0, 0): multiplies the bias row of B
      a[7] = 0u;
      tmem_st8(lane_base + G::COL_A, a);
    }
    wait_st();
    fence_before();
    TC_STAMP(1);
    if (TC_EXP != 4) tile_sync();
    TC_STAMP(2);
    if (owner && warp == 0 && TC_EXP != 3 && TC_EXP != 4) {
      if (elect_one()) {
        fence_after();
        mma_ts(tmem + G::COL_D, tmem + G::COL_A, chunk_desc(sb, G::OFF_B1A, HID, 0), IDH, 0u);
        mma_ts(tmem + G::COL_D, tmem + G::COL_A, chunk_desc(sb, G::OFF_B1B, HID, 0), IDH, 1u);
        mma_commit(bar);
      }
      __syncwarp();
    }
    // The running cost of this step (state before the dynamics, PI/mppi_controller.cu:162-165) is spread over the waits
    // for the three layers.  Here: control + speed + crash + track, in the reference's summation order (PI/costs.cu:396-409).
    TC_STAMP(3);
    float cost_acc = 0.0f;
    if (owner && i > 0 && TC_EXP != 1) {
      bool boundary;
      const float track = cost_track_part(p.cp, front, back, boundary);
      if (boundary) crash = 1;
      const float pre = cost_pre_part(p.cp, control_cost, s[4]);
      cost_acc = __fadd_rn(__fadd_rn(pre, crash > 0 ? p.cp.crash_cost_on : 0.0f), track);
    }
    TC_STAMP(4);
    ok = ok && mbar_wait<(CS > 1)>(bar, phase);  // after one time-out nothing waits again: a broken launch ends in seconds
    phase ^= 1u;
    fence_after();
    TC_STAMP(5);

    // ---- following layers: tanh epilogue -> hi / lo activations back into TMEM -> 3 MMAs per K = 16 chunk ----
#pragma unroll
    for (int layer = 1; layer <= NHID; layer++) {  // layer = index of the layer whose MMAs are issued here (NHID = output layer)
#pragma unroll
      for (int c = 0; c < G::NCH; c++) {
        if (CS > 1 && (c % CS) != slice) continue;  // warp-uniform: this slice's chunks only
        float v[16];
        uint32_t h[16];
        tmem_ld16(lane_base + G::COL_D + 16 * c, v);
        wait_ld();
        if (c == 0) TC_STAMP(6 * layer + 0);
        if (layer == 1) activate16<true>(v, nullptr, h);
        else activate16<false>(v, &ep.eb[layer - 1][16 * c], h);
        if (c == 0) TC_STAMP(6 * layer + 1);
        tmem_st16(lane_base + G::COL_A + 16 * c, h);
      }
      wait_st();
      fence_before();
      TC_STAMP(6 * layer + 2);
      if (TC_EXP != 4) tile_sync();
      TC_STAMP(6 * layer + 3);
      if (owner && warp == 0 && TC_EXP != 3 && TC_EXP != 4) {
        if (elect_one()) {
          fence_after();
          const bool last = (layer == NHID);
          const int N = last ? 16 : HID;
          const uint32_t id = last ? ID16 : IDH;
          const int bh = last ? G::OFF_BL : G::OFF_BH + (layer - 1) * 2 * G::SZ_BH, bl = bh + (last ? G::SZ_BL : G::SZ_BH);
#pragma unroll
          for (int c = 0; c < G::NCH; c++) mma_ts(tmem + G::COL_D, tmem + G::COL_A + 16 * c + 8, chunk_desc(sb, bh, N, c), id, c > 0 ? 1u : 0u);  // lo x W_hi
#pragma unroll
          for (int c = 0; c < G::NCH; c++) mma_ts(tmem + G::COL_D, tmem + G::COL_A + 16 * c, chunk_desc(sb, bl, N, c), id, 1u);                  // hi x W_lo
#pragma unroll
          for (int c = 0; c < G::NCH; c++) mma_ts(tmem + G::COL_D, tmem + G::COL_A + 16 * c, chunk_desc(sb, bh, N, c), id, 1u);                  // hi x W_hi
          mma_commit(bar);
        }
        __syncwarp();
      }
      TC_STAMP(6 * layer + 4);
      if (layer == NHID) break;
      // second wait: the stabilizing cost (atan of the slip angle), the NaN / 1e12 clamp and the running mean (:162-165)
      if (owner && layer == 1 && i > 0 && TC_EXP != 1) {
        float c = __fadd_rn(cost_acc, cost_stab_part(p.cp, s[4], s[5]));
        if (c > 1e12f || isnan(c)) c = 1e12f;
        running = (float)((double)running + (double)__fsub_rn(c, running) * p.inv_step[i]);
      }
      ok = ok && mbar_wait<(CS > 1)>(bar, phase);  // after one time-out nothing waits again: a broken launch ends in seconds
      phase ^= 1u;
      fence_after();
      TC_STAMP(6 * layer + 5);
    }
    // third wait: kinematics (PI/neural_net_model.cu:346-355, precise sinf / cosf); x, y, yaw of the next state do not
    // depend on the network, so they are advanced now and the next step's costmap texels and controls are requested
    if (owner) {
      float sn, cs;
      if (TC_EXP == 5) __sincosf(s[2], &sn, &cs); else sincosf(s[2], &sn, &cs);
      const float d0 = fmaf(cs, s[4], -__fmul_rn(sn, s[5]));
      const float d1 = fmaf(sn, s[4], __fmul_rn(cs, s[5]));
      const float d2 = p.negate_yaw ? -s[6] : s[6];
      s[0] = fmaf(d0, p.dt, s[0]);  // incrementState, PI/neural_net_model.cu:334-344
      s[1] = fmaf(d1, p.dt, s[1]);
      s[2] = fmaf(d2, p.dt, s[2]);
    }
    TC_STAMP(36);
    if (owner && i + 1 < p.T) {
      if (TC_EXP != 1) track_lookups(p.cp, p.tex, s[0], s[1], s[2], front, back);
      prepare_controls(i + 1);
    }
    TC_STAMP(37);
    ok = ok && mbar_wait<(CS > 1)>(bar, phase);  // after one time-out nothing waits again: a broken launch ends in seconds
    phase ^= 1u;
    fence_after();
    TC_STAMP(38);
    if (owner) {
      float o[4];
      tmem_ld4(lane_base + G::COL_D, o);
      wait_ld();
#pragma unroll
      for (int k = 0; k < 4; k++) s[3 + k] = fmaf(__fadd_rn(o[k], ep.b_last[k]), p.dt, s[3 + k]);
      if (fabsf(s[3]) >= 1.57f) crash = 1;  // getCrash, PI/costs.cu:301-305
    }
    TC_STAMP(39);
  }
#if TC_EXP == 9
  if (cta_tid == 0 && blockIdx.x == 0)
    for (int k = 0; k < 40; k++) row[k] = make_float2((float)(stamps[k] - stamps[0]), (float)k);
#endif

  // ---- epilogue: costs, crash flags, min-cost baseline; release the tensor memory ----
  if (!ok) running = __int_as_float(0x7fc00000);  // the MMA never signalled: poison the result (tests catch it)
  unsigned int best = 0xffffffffu;
  if (valid) {
    p.costs[g0] = running;
    p.crash[g0] = (unsigned char)crash;
    best = float_to_ordered(running);
  }
  const unsigned int wbest = __reduce_min_sync(0xffffffffu, best);
  const int wctrl = __shfl_sync(0xffffffffu, ctrl, 0);
  if ((tid & 31) == 0 && wbest != 0xffffffffu) atomicMin(p.baseline + wctrl, wbest);
  fence_before();
  __syncthreads();
  if (cta_tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_cta), "r"(G::TMEM_COLS * G::TPC) : "memory");
}

}  // namespace tc

template <int HID, int NHID, int TPC, bool FUSED, int CS = 1>
static cudaError_t launch_tc_f(const RolloutParams &p, cudaStream_t st, const float *host_theta_t, bool pdl) {
  using G = tc::Geo<HID, NHID, TPC, CS>;
  const long long total = (long long)p.B * p.n_local;
  const unsigned grid = (unsigned)((total + tc::TILE * TPC - 1) / (tc::TILE * TPC));
  // the folded biases are b + rowsum(W) (tanh = 1 - 2r travels as r); theta_t holds Wt[k][j] per layer, then b[j]
  tc::TcEpilogue<HID, NHID> ep;
  for (int j = 0; j < HID; j++) ep.eb[0][j] = 1.0f;  // unused: the first layer's bias rides in the K padding of its MMA
  for (int h = 1; h < NHID; h++)
    for (int j = 0; j < HID; j++) {
      double sum = host_theta_t[G::th_b(h) + j];
      for (int k = 0; k < HID; k++) sum += (double)host_theta_t[G::th_w(h) + k * HID + j];
      ep.eb[h][j] = (float)std::exp(2.0 * sum);
    }
  for (int j = 0; j < 4; j++) {
    double sum = host_theta_t[G::TH_BL + j];
    for (int k = 0; k < HID; k++) sum += (double)host_theta_t[G::TH_WL + k * 4 + j];
    ep.b_last[j] = (float)sum;
  }
  // decided once per instantiation AND device: cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute
  // (a second context on another GPU of the same process would otherwise launch without the opt-in)
  static int smem_bytes_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int &smem_bytes = smem_bytes_dev[dev & 63];
  if (smem_bytes == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, tc::rollout_tc_kernel<HID, NHID, TPC, FUSED, CS>);
    if (e != cudaSuccess) return e;
    int bytes = G::SMEM_BYTES;
    const long long tmem_ctas = 512 / (G::TMEM_COLS * G::TPC);  // CTAs per SM the tensor memory admits
    const bool regs_admit_more = (long long)fa.numRegs * G::THREADS * (tmem_ctas + 1) <= 65536;
    const bool smem_admits_more = (long long)(bytes + 2048) * (tmem_ctas + 1) <= 228 * 1024;
    if (regs_admit_more && smem_admits_more) bytes = ((228 / (int)(tmem_ctas + 1)) - 1) * 1024;  // keep TMEM the only limiter
    if (bytes > 48 * 1024) {
      e = cudaFuncSetAttribute(tc::rollout_tc_kernel<HID, NHID, TPC, FUSED, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e != cudaSuccess) return e;
    }
    smem_bytes = bytes;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(G::THREADS); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, tc::rollout_tc_kernel<HID, NHID, TPC, FUSED, CS>, p, ep);
}

template <int HID, int NHID, int TPC, int CS = 1>
static cudaError_t launch_tc(const RolloutParams &p, cudaStream_t st, const float *host_theta_t, bool pdl) {
  return p.fused_noise ? launch_tc_f<HID, NHID, TPC, true, CS>(p, st, host_theta_t, pdl) : launch_tc_f<HID, NHID, TPC, false, CS>(p, st, host_theta_t, pdl);
}

cudaError_t launch_rollout_nn32_tc(const RolloutParams &p, cudaStream_t st, const float *host_theta_t, bool pdl) { return launch_tc<32, 2, 1>(p, st, host_theta_t, pdl); }
cudaError_t launch_rollout_nn64_tc(const RolloutParams &p, cudaStream_t st, const float *host_theta_t, bool pdl) {
  // up to two one-tile CTAs per SM: spread the tiles; beyond that, two tiles per CTA share the weights (4 tiles per SM)
  const long long tiles = ((long long)p.B * p.n_local + tc::TILE - 1) / tc::TILE;
  // one tile per SM at most (controller-sized problems): four threads per rollout share the tanh epilogues
  if (tiles <= 148 && !getenv("MPPI_TC_NO_SLICES")) return launch_tc<64, 4, 1, 4>(p, st, host_theta_t, pdl);
  return tiles <= 2 * 148 ? launch_tc<64, 4, 1>(p, st, host_theta_t, pdl) : launch_tc<64, 4, 2>(p, st, host_theta_t, pdl);
}

// True when every folded bias of the network keeps e^(2 b) inside the FP32 range (|b| < 40); otherwise the caller uses
// the FP32 kernels.  widths = {6, HID x NHID, 4}.
bool tc_biases_in_range(const float *host_theta_t, int hid, int nhid) {
  const int th_b1 = 6 * hid;
  for (int j = 0; j < hid; j++)
    if (!(std::fabs(host_theta_t[th_b1 + j]) < 40.0f)) return false;
  for (int h = 1; h < nhid; h++) {
    const int w = 7 * hid + (h - 1) * (hid * hid + hid), b = w + hid * hid;
    for (int j = 0; j < hid; j++) {
      double sum = host_theta_t[b + j];
      for (int k = 0; k < hid; k++) sum += (double)host_theta_t[w + k * hid + j];
      if (!(std::fabs(sum) < 40.0)) return false;
    }
  }
  return true;
}

}  // namespace mppi
