// GeneralizedLinear<CarBasisFuncs,7,2,25,CarKinematics,3> (PI/generalized_linear.cu:168-245, PI/car_bfs.cuh:44-120).
//
// Two kernels:
//  * rollout_kernel<CarBasisDyn>: one rollout per thread (rollout.cuh), the throughput shape.
//  * rollout_bf_split_kernel: the latency shape for a few thousand rollouts.  The recursion of the basis-function model
//    only involves roll, u_x, u_y and the yaw rate: it never reads the position, the yaw or the cost.  So the work of one
//    rollout is split over TWO WARPS of a 64-thread CTA that owns 32 rollouts: warp 0 runs the dynamics chain (25 basis
//    functions with tanf / atanf / sinf, the 4 x 25 contraction, Euler step) and publishes the state of every timestep in a
//    shared-memory ring; warp 1 follows behind and does everything else for the same 32 rollouts (noise, un-clamped
//    write-back, kinematics, costmap texels, all cost terms, running mean, crash flag).  The two dependent chains of about
//    0.5 us per timestep run concurrently instead of back to back.  Every variable sees the same operations in the same
//    order as in the one-thread kernel, so the results are bit-identical.
#include "rollout_launch_impl.cuh"
#include <cstdlib>

namespace mppi {

namespace {

constexpr int RING = 8;  // timesteps the dynamics warp may run ahead

struct BfRing {
  float s[RING][4][32];  // (roll, u_x, u_y, yaw rate) BEFORE step i, slot i % RING
  volatile int produced; // states 0 .. produced-1 are published
  volatile int consumed; // timesteps the cost warp has finished
  volatile int failed;   // a bounded wait ran out (never in a working build): results are poisoned, the GPU does not hang
};

__device__ __forceinline__ bool wait_at_least(volatile int *counter, int want, volatile int *failed) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 24); spin++) {
    if (*counter >= want) return true;
    if (*failed) return false;
    __nanosleep(20);
  }
  *failed = 1;
  return false;
}

__global__ void __launch_bounds__(64) rollout_bf_split_kernel(const __grid_constant__ RolloutParams p) {
  __shared__ BfRing ring;
  __shared__ __align__(16) float sw[100];
  const int tid = threadIdx.x, lane = tid & 31, role = tid >> 5;
  for (int i = tid; i < 100; i += 64) sw[i] = p.theta_t[i];
  if (tid == 0) { ring.produced = 0; ring.consumed = 0; ring.failed = 0; }
  __syncthreads();

  const long long g0 = (long long)blockIdx.x * 32 + lane;  // n_local is a multiple of 64: every lane is a valid rollout
  const int ctrl = (int)(g0 / p.n_local);
  const int lr0 = (int)(g0 - (long long)ctrl * p.n_local);
  const float *inbox = p.inbox + (size_t)ctrl * p.inbox_stride;
  const float2 *U = reinterpret_cast<const float2 *>(inbox + INBOX_U);
  const int rg = p.r_begin + lr0;
  const bool noise_free = (rg == 0), pure_noise = (rg >= p.pure_noise_from);
  float2 *row = reinterpret_cast<float2 *>(p.du) + (size_t)g0 * p.T;

  // the clamped control of step i (PI/mppi_controller.cu:130-155 + enforceConstraints); both warps evaluate it
  auto controls = [&](int i, float2 e, float &du0, float &du1, float &uu0, float &uu1, float &u0, float &u1) {
    const float2 Ui = U[i];
    if (noise_free || i < p.opt_delay) {
      du0 = 0.0f; du1 = 0.0f; uu0 = Ui.x; uu1 = Ui.y;
    } else if (pure_noise) {
      du0 = __fmul_rn(e.x, p.nu0); du1 = __fmul_rn(e.y, p.nu1); uu0 = du0; uu1 = du1;
    } else {
      du0 = __fmul_rn(e.x, p.nu0); du1 = __fmul_rn(e.y, p.nu1);
      uu0 = __fadd_rn(Ui.x, du0); uu1 = __fadd_rn(Ui.y, du1);
    }
    u0 = uu0 < p.lo0 ? p.lo0 : (uu0 > p.hi0 ? p.hi0 : uu0);
    u1 = uu1 < p.lo1 ? p.lo1 : (uu1 > p.hi1 ? p.hi1 : uu1);
  };

  if (role == 0) {
    // ---------------- dynamics warp ----------------
    float s3 = inbox[INBOX_STATE + 3], s4 = inbox[INBOX_STATE + 4], s5 = inbox[INBOX_STATE + 5], s6 = inbox[INBOX_STATE + 6];
    float2 e_next = row[0];
    for (int i = 0; i <= p.T; i++) {
      // slot i % RING was last read by the cost warp in steps i - RING (as its state) and i - RING - 1 (as its successor)
      if (i >= RING && !wait_at_least(&ring.consumed, i - RING + 1, &ring.failed)) break;
      ring.s[i % RING][0][lane] = s3; ring.s[i % RING][1][lane] = s4; ring.s[i % RING][2][lane] = s5; ring.s[i % RING][3][lane] = s6;
      __threadfence_block();
      __syncwarp();
      if (lane == 0) ring.produced = i + 1;
      if (i == p.T) break;
      const float2 e = e_next;
      if (i + 1 < p.T) e_next = row[i + 1];  // consumed one iteration later, before the state the cost warp waits for is published
      float du0, du1, uu0, uu1, u0, u1;
      controls(i, e, du0, du1, uu0, uu1, u0, u1);
      const float in[6][1] = {{s3}, {s4}, {s5}, {s6}, {u0}, {u1}};
      float out[4][1];
      CarBasisDyn::deriv(sw, nullptr, in, out);
      s3 = fmaf(out[0][0], p.dt, s3); s4 = fmaf(out[1][0], p.dt, s4); s5 = fmaf(out[2][0], p.dt, s5); s6 = fmaf(out[3][0], p.dt, s6);
    }
    return;
  }
  // ---------------- cost warp ----------------
  float x = inbox[INBOX_STATE + 0], y = inbox[INBOX_STATE + 1], yaw = inbox[INBOX_STATE + 2];
  float running = 0.0f, front = 0.0f, back = 0.0f;
  int crash = 0;
  bool ok = true;
  for (int i = 0; i < p.T; i++) {
    if (!wait_at_least(&ring.produced, i + 2, &ring.failed)) { ok = false; break; }
    __threadfence_block();
    const float s3 = ring.s[i % RING][0][lane], s4 = ring.s[i % RING][1][lane], s5 = ring.s[i % RING][2][lane], s6 = ring.s[i % RING][3][lane];
    const float roll_next = ring.s[(i + 1) % RING][0][lane];
    // The noise of step i is consumed by the dynamics warp at least one published state earlier (it reads row[i] before it
    // publishes state i + 1, and this warp needs state i + 1 before it gets here), so overwriting row[i] is safe.
    const float2 e = row[i];
    float du0, du1, uu0, uu1, u0, u1;
    controls(i, e, du0, du1, uu0, uu1, u0, u1);
    row[i] = make_float2(uu0, uu1);  // un-clamped write-back (:153)
    if (i > 0) {
      const float c = running_cost_from_parts(p.cp, step_cost_from_lookups(p.cp, front, back, s4, s5, u0, u1, du0, du1, p.nu0, p.nu1), crash);
      running = (float)((double)running + (double)__fsub_rn(c, running) * p.inv_step[i]);
    }
    float sn, cs;
    sincosf(yaw, &sn, &cs);
    const float d0 = fmaf(cs, s4, -__fmul_rn(sn, s5));
    const float d1 = fmaf(sn, s4, __fmul_rn(cs, s5));
    const float d2 = p.negate_yaw ? -s6 : s6;
    x = fmaf(d0, p.dt, x); y = fmaf(d1, p.dt, y); yaw = fmaf(d2, p.dt, yaw);
    if (i + 1 < p.T) track_lookups(p.cp, p.tex, x, y, yaw, front, back);
    if (fabsf(roll_next) >= 1.57f) crash = 1;  // getCrash after the state update, PI/costs.cu:301-305
    (void)s3;
    __syncwarp();
    if (lane == 0) ring.consumed = i + 1;
  }
  if (!ok || ring.failed) running = __int_as_float(0x7fc00000);
  p.costs[g0] = running;
  p.crash[g0] = (unsigned char)crash;
  const unsigned int wbest = __reduce_min_sync(0xffffffffu, float_to_ordered(running));
  if (lane == 0) atomicMin(p.baseline + ctrl, wbest);
}

}  // namespace

// true when launch_rollout_bf would pick the two-warp latency kernel (which reads its noise from the buffer in both warps)
bool rollout_bf_is_split(long long total) {
  static const char *force = std::getenv("MPPI_BF_SPLIT");
  return force ? std::atoi(force) != 0 : total <= 148LL * 32 * 4;
}

cudaError_t launch_rollout_bf(const RolloutParams &p, cudaStream_t st, bool small) {
  const long long total = (long long)p.B * p.n_local;
  // latency regime (under four 32-rollout CTAs per SM): the two-warp split; MPPI_BF_SPLIT=0/1 forces either kernel
  static const char *force = std::getenv("MPPI_BF_SPLIT");
  const bool split = !p.fused_noise && (force ? std::atoi(force) != 0 : total <= 148LL * 32 * 4);
  if (split) {
    rollout_bf_split_kernel<<<(unsigned)(total / 32), 64, 0, st>>>(p);
    return cudaGetLastError();
  }
  return small ? launch_rollout_t<CarBasisDyn, 32, 1, true>(p, st) : launch_rollout_t<CarBasisDyn, 128, 1, true>(p, st);
}

}  // namespace mppi
