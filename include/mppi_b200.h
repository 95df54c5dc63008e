/*
 * mppi_b200.h -- C ABI of libmppi_b200.so: the B200 (sm_100a) implementation of AutoRally's MPPI
 * controller hot path, MPPIController<DYNAMICS_T,COSTS_T,ROLLOUTS,BDIM_X,BDIM_Y>::computeControl.
 *
 * The reference has no FFI: its host C++ templates call CUDA directly.  This header is the seam a
 * maintainer binds instead; each entry point names the reference code it replaces ("PI/" =
 * autorally_control/include/autorally_control/path_integral/ in rdesc/autorally).  The drop-in C++
 * templates in include/autorally_control/path_integral/ are a thin host layer over these calls.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success, a positive
 * cudaError_t value for a CUDA failure, or a negative MPPI_ERR_* code.  Nothing throws and nothing
 * calls exit() (the reference prints and continues, PI/gpu_err_chk.h:49).  There is no CPU fallback:
 * without a CUDA device mppi_create fails.
 *
 * Layouts (float32, identical to the reference):
 *   state   [7]      x, y, yaw, roll, u_x, u_y, yaw_rate          (PI/run_control_loop.cuh:147)
 *   U       [T][2]   steering, throttle                            (PI/mppi_controller.cuh:207)
 *   noise / sampled controls  [rollout][t][2], index 2*T*r + 2*t + j (PI/mppi_controller.cu:133)
 * In batched mode (num_controllers = B > 1) every per-controller array gains a leading [B] axis.
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_STATE_DIM 7
#define MPPI_CONTROL_DIM 2

enum {
  MPPI_OK = 0,
  MPPI_ERR_INVALID_ARG = -1,   /* null pointer, bad size, rollout count not a multiple of 64 ... */
  MPPI_ERR_UNSUPPORTED = -2,   /* e.g. an MLP structure no kernel is instantiated for */
  MPPI_ERR_NOT_READY = -3,     /* compute called before weights / costmap / cost params were set */
  MPPI_ERR_NO_DEVICE = -4,     /* no CUDA device: there is no CPU fallback */
  MPPI_ERR_ALLOC = -5,
  MPPI_ERR_COMM = -6           /* libnccl.so.2 not loadable, or an NCCL call failed */
};

enum { MPPI_DYNAMICS_NN = 0, MPPI_DYNAMICS_BF = 1 };

/* Rollout-kernel variants (mppi_config.rollout_variant).  AUTO picks by network and problem size (rollouts of all
 * controllers of the context together): 6-32-32-4 up to 512 rollouts WARP32, up to 16384 HALF16, above TENSOR;
 * 6-64-64-64-64-4 up to 2368 rollouts LAYER_PIPE, above TENSOR; any other layer pack (widths <= 128) GENERIC; basis
 * functions THREAD1.  A network whose folded biases would leave the FP32 range in the tensor kernel's e^(2b) constants
 * (|b| >= 40) runs on the FP32 kernels instead.  The numeric values are stable (3-8 were experimental designs of round 1
 * and are retired). */
enum {
  MPPI_ROLLOUT_AUTO = 0,
  MPPI_ROLLOUT_THREAD1 = 1, /* one rollout per thread, weights broadcast from shared memory */
  MPPI_ROLLOUT_THREAD2 = 2, /* two rollouts per thread (register-tiled FFMA2), weights from shared memory */
  MPPI_ROLLOUT_HALF16 = 9,  /* one rollout per half-warp, FFMA2 over neuron pairs, deferred running mean: the latency default */
  MPPI_ROLLOUT_TENSOR = 10, /* one rollout per thread, layer contractions on tcgen05 (FP16 hi/lo split, A in tensor memory) */
  MPPI_ROLLOUT_GENERIC = 11, /* run-time layer pack (NeuralNetModel<7,2,3,6,...,4>, widths <= 128): one or two rollouts per warp, FP32 */
  MPPI_ROLLOUT_LAYER_PIPE = 12, /* 6-64-64-64-64-4 only: each hidden layer in the registers of one warp, rollouts flow through the warps: the
                                   latency default of that network up to 2368 rollouts (one wave) */
  MPPI_ROLLOUT_WARP32 = 13 /* 6-32-32-4: one rollout per warp, one neuron per lane, weights in registers: the latency default up to 512 rollouts */
};

/* Replaces the MPPIController template/ctor arguments (PI/mppi_controller.cuh:52-53,101-102) plus the
 * sharding description of SURVEY.md section 8(e). */
typedef struct mppi_config {
  int dynamics;         /* MPPI_DYNAMICS_NN | MPPI_DYNAMICS_BF */
  int num_rollouts;     /* global NUM_ROLLOUTS per controller; must be a multiple of 64 (:58-60) */
  int num_timesteps;    /* T (numTimesteps_) */
  int num_controllers;  /* B independent controllers sharing model and costs; 1 = reference */
  int rollout_begin;    /* this GPU's shard of the global rollout range ...            */
  int rollout_count;    /* ... [begin, begin+count); count 0 = all; multiple of 64      */
  int hz;               /* dt = 1/hz (SRC/path_integral_main.cu:100) */
  int optimization_stride; /* = opt_delay of rolloutKernel (PI/mppi_controller.cu:616) */
  float gamma;          /* softmax temperature (:195) */
  int num_iters;        /* optimisation iterations per computeControl (:609) */
  int bdim_x, bdim_y;   /* reference block shape; informational, kept for API parity */
  int device;           /* CUDA device ordinal, -1 = current device */
  int rollout_variant;  /* MPPI_ROLLOUT_* */
  int controller_begin; /* batched mode sharded over GPUs: global index of this context's controller 0 (Philox counter) */
  uint64_t seed;        /* Philox key; the reference seeds cuRAND with 1234 (:331) */
} mppi_config;

/* MPPICosts::CostParams (PI/costs.cuh:67-85), same field order, plus the class member l1_cost_. */
typedef struct mppi_cost_params {
  float desired_speed, speed_coeff, track_coeff, max_slip_ang, slip_penalty, track_slop, crash_coeff;
  float steering_coeff, throttle_coeff, boundary_threshold, discount;
  int num_timesteps, grid_res;
  float r_c1[3], r_c2[3], trs[3];
  int l1_cost;
} mppi_cost_params;

/* Scalars computeControl leaves in the controller (PI/mppi_controller.cu:627-652). */
typedef struct mppi_result {
  float baseline;        /* min rollout cost */
  float normalizer;      /* normalizer_ = sum_i exp(-gamma (c_i - baseline)) */
  float trajectory_cost; /* trajectory_cost_ = sum_i w_i^2 / normalizer_ (getComputedTrajectoryCost) */
  float reserved;
} mppi_result;

typedef struct mppi_ctx mppi_ctx;

const char *mppi_version(void);
const char *mppi_error_string(int code);
void mppi_config_default(mppi_config *cfg); /* launch/path_integral_nn.launch values, 1920 rollouts */

/* ctor + allocateCudaMem (PI/mppi_controller.cu:321-387) / deallocateCudaMem (:389-400). */
int mppi_create(const mppi_config *cfg, mppi_ctx **out);
int mppi_destroy(mppi_ctx *ctx);

/* NeuralNetModel::paramsToDevice (PI/neural_net_model.cu:120-150): theta packed [W1|b1|W2|b2|...],
 * row-major, net_structure e.g. {6,32,32,4}. */
int mppi_set_nn_params(mppi_ctx *ctx, const float *theta, const int *net_structure, int num_layers);
/* GeneralizedLinear::paramsToDevice (PI/generalized_linear.cu:110-116): theta 4x25 row-major. */
int mppi_set_bf_params(mppi_ctx *ctx, const float *theta_4x25);
/* control_rngs_ upload (PI/neural_net_model.cu:149): {steer_lo, steer_hi, throttle_lo, throttle_hi}. */
int mppi_set_control_ranges(mppi_ctx *ctx, const float lo_hi[4]);
/* negate_yaw_der member (PI/neural_net_model.cuh:75); ignored by the BF model, which always negates. */
int mppi_set_negate_yaw_der(mppi_ctx *ctx, int negate);
/* MPPICosts::paramsToDevice (PI/costs.cu:234-238). */
int mppi_set_cost_params(mppi_ctx *ctx, const mppi_cost_params *p);
/* MPPICosts::costmapToTexture (PI/costs.cu:99-154): row-major H x W texels, `channels` = 1 (channel0
 * only) or 4 (the reference's interleaved float4, of which only .x is used, PI/costs.cu:379-380). */
int mppi_set_costmap(mppi_ctx *ctx, const float *texels, int width, int height, int channels);
/* nu_ upload (PI/mppi_controller.cu:345,357).  This is what north_star calls updateControlNoise
 * (absent from the reference): exploration standard deviations {steering, throttle}. */
int mppi_set_exploration_std(mppi_ctx *ctx, const float std2[2]);
int mppi_set_gamma(mppi_ctx *ctx, float gamma);

/* Noise.  Default: every compute call samples N(0,1) with the Philox4x32-10 kernel (replaces
 * curandGenerateNormal, PI/mppi_controller.cu:612).  mppi_set_noise injects host noise
 * [num_iters][B][rollout_count][T][2] for parity runs and disables sampling until
 * mppi_use_sampler() is called. */
int mppi_set_noise(mppi_ctx *ctx, const float *eps, size_t count);
int mppi_use_sampler(mppi_ctx *ctx);
int mppi_seed(mppi_ctx *ctx, uint64_t seed, uint32_t call_counter);
/* Where the sampled noise is produced: 1 = inside the rollout kernel (no separate sampler launch, no noise round trip
 * through HBM), 0 = by the stand-alone sampler kernel into the [rollouts x T x 2] buffer, -1 = automatic (in place for the
 * throughput kernels, the sampler kernel for the latency kernels).  Both draw the same Philox stream: results are
 * bit-identical.  Ignored while injected noise is in use. */
int mppi_set_fused_noise(mppi_ctx *ctx, int mode);
/* Runs only the sampler into the noise buffer and copies it out (tests). */
int mppi_sample_noise(mppi_ctx *ctx, float *eps_out /* [B][rollout_count][T][2] or NULL */);

/* computeControl(state) (PI/mppi_controller.cu:600-675): noise -> rollouts -> importance weighting
 * -> control update -> Savitzky-Golay -> nominal trajectory.  All host pointers; [B] leading axis in
 * batched mode.  U is read and overwritten with the smoothed sequence.  control_hist is the
 * controller's control_hist_ (2 past controls, [4]).  state_solution [T][7] / control_solution [T][2]
 * / result may be NULL. */
int mppi_compute_control(mppi_ctx *ctx, const float *state, float *U, const float *control_hist,
                         float *state_solution, float *control_solution, mppi_result *result);

/* The two halves of mppi_compute_control: enqueue everything on the context's stream and return; wait and unpack.
 * Contexts own independent streams, so the two controllers of the reference's control loop (actual-state and
 * predicted-state, PI/run_control_loop.cuh:218-219), each of which fills a few percent of a B200, can run
 * concurrently: async on both, then wait on both. */
int mppi_compute_control_async(mppi_ctx *ctx, const float *state, const float *U, const float *control_hist);
int mppi_compute_control_wait(mppi_ctx *ctx, float *U, float *state_solution, float *control_solution, mppi_result *result);

/* Measurement helper: `reps` consecutive mppi_compute_control calls made from C with the given host buffers (U fed back
 * from call to call); latency_ms[reps] receives each call's host-observed duration (steady clock). */
int mppi_bench_compute_control(mppi_ctx *ctx, const float *state, float *U, const float *control_hist, int reps,
                               float *latency_ms);

/* Stage access for parity tests and multi-GPU plumbing (after a compute call): */
int mppi_get_rollout_costs(mppi_ctx *ctx, float *costs /* [B][rollout_count] */);
int mppi_get_rollout_crash(mppi_ctx *ctx, int *crash /* [B][rollout_count] */);
int mppi_get_sampled_controls(mppi_ctx *ctx, float *V /* [B][rollout_count][T][2] */);
int mppi_get_unsmoothed_controls(mppi_ctx *ctx, float *U_new /* [B][T][2], before Savitzky-Golay */);

/* Multi-GPU (nothing like it in the reference).  Each rank runs
 * mppi_shard_begin (noise, rollouts and the local weighting partials), exchanges the
 * mppi_shard_floats() floats per controller that mppi_shard_partials_device() points at (device
 * memory; all-gather over NCCL), and finishes with mppi_shard_finish on the gathered buffer
 * [num_shards][B][mppi_shard_floats].  One shard reproduces mppi_compute_control. */
int mppi_shard_floats(const mppi_ctx *ctx); /* 3 + 2T, padded to a multiple of 4 */
int mppi_shard_begin(mppi_ctx *ctx, const float *state, const float *U, const float *control_hist);
int mppi_shard_partials_device(mppi_ctx *ctx, float **dev_ptr);
int mppi_shard_finish(mppi_ctx *ctx, const float *gathered_dev, int num_shards, float *U,
                      float *state_solution, float *control_solution, mppi_result *result);

/* Asynchronous halves of the two calls above, for callers that order the exchange on a stream instead of the host
 * (bench.py, the in-library NCCL path).  mppi_shard_begin_async with state == NULL reuses the device-resident
 * state / U / history (the smoothed U of the previous mppi_shard_finish_async with feed_back != 0).  Nothing
 * synchronises until mppi_shard_result, which waits for the stream and unpacks the last finish. */
int mppi_shard_begin_async(mppi_ctx *ctx, const float *state, const float *U, const float *control_hist);
int mppi_shard_finish_async(mppi_ctx *ctx, const float *gathered_dev, int num_shards, int feed_back);
int mppi_shard_result(mppi_ctx *ctx, float *U, float *state_solution, float *control_solution, mppi_result *result);
/* Run on the caller's stream (e.g. the one its NCCL calls are ordered on) instead of the context's own. */
int mppi_set_stream(mppi_ctx *ctx, void *cuda_stream);

/* In-library exchange over NCCL (libnccl.so.2 is loaded at run time with dlopen; the library has no link-time
 * dependency on it).  One process per GPU: rank 0 calls mppi_comm_unique_id, ships the 128 bytes to the other
 * ranks out of band, and every rank calls mppi_comm_init.  mppi_compute_control_sharded is then computeControl
 * for this rank's shard of the rollouts: sampler -> rollouts -> local weighting record -> ONE ncclAllGather of
 * mppi_shard_floats() floats per controller -> finalize, all on the context's stream; every rank returns the same U. */
int mppi_comm_unique_id(void *id_out_128_bytes);
int mppi_comm_init(mppi_ctx *ctx, const void *id_128_bytes, int rank, int num_ranks);
int mppi_comm_destroy(mppi_ctx *ctx);
int mppi_compute_control_sharded(mppi_ctx *ctx, const float *state, float *U, const float *control_hist,
                                 float *state_solution, float *control_solution, mppi_result *result);
/* The same exchange WITHOUT a collective: over NVLink peer memory, fused into the kernels on either side of it.  The
 * CTA that completes a rank's record stores it directly into every GPU's mailbox and raises a flag; finalize_kernel
 * waits for the G flags of its controller (bounded: a missing peer yields MPPI_ERR_COMM, not a hang).  One process per
 * GPU on one NVSwitch box.  Set-up: every rank calls mppi_p2p_export (2 x 64-byte CUDA IPC handles), the handles of all
 * ranks are exchanged out of band in rank order, every rank calls mppi_p2p_init.  Afterwards
 * mppi_compute_control_sharded / mppi_run_resident_sharded use this path instead of NCCL. */
int mppi_p2p_export(mppi_ctx *ctx, int num_ranks, void *handles_out_128_bytes);
int mppi_p2p_init(mppi_ctx *ctx, const void *all_handles /* [num_ranks][128 bytes] */, int rank, int num_ranks);
int mppi_p2p_destroy(mppi_ctx *ctx);

/* Device-resident sharded stepping (throughput measurement): `steps` sharded pipelines back to back, no host copies. */
int mppi_run_resident_sharded(mppi_ctx *ctx, int steps, float *elapsed_ms);

/* Device-resident stepping for throughput measurement: runs `steps` complete pipelines back to back
 * with state/U/history already in HBM and no host copies (the smoothed U of one step warm-starts the
 * next).  Returns the elapsed device time (CUDA events on the context's stream) and, if non-NULL, the
 * time spent in the rollout kernel alone.  flush_l2 != 0 writes a 256 MiB scratch buffer before every
 * step, outside the timed intervals, and sums the per-step intervals instead. */
int mppi_run_resident(mppi_ctx *ctx, int steps, int flush_l2, float *elapsed_ms, float *rollout_kernel_ms);
/* Measurement helper: device time (ms, CUDA events, every stage serialised, mean of `reps` repetitions) of the four stages
 * of one pipeline on the resident inputs: {stand-alone sampler kernel (replaces curandGenerateNormal, PI/mppi_controller.cu:612;
 * timed even when the rollout kernel draws its noise in place), rollout kernel (:72-184), weighting kernel (host min /
 * normaliser loops + normExpKernel + weightedReductionKernel, :193-267, :627-652), finalize kernel (:468-519, :663-667)}. */
int mppi_time_stages(mppi_ctx *ctx, int reps, float stage_ms[4]);
int mppi_get_stream(mppi_ctx *ctx, void **cuda_stream);
int mppi_synchronize(mppi_ctx *ctx);
/* Number of kernels the last compute call launched (bench.py's gpu_launches). */
int mppi_last_launch_count(const mppi_ctx *ctx);
/* The variant AUTO resolved to (MPPI_ROLLOUT_*). */
int mppi_resolved_variant(const mppi_ctx *ctx);

/* Roofline denominators measured on the device this library runs on: dependent-free FFMA issue
 * (TFLOP/s, FP32 CUDA cores) and a device-to-device copy (GB/s, read+write bytes). */
int mppi_measure_fp32_peak(int device, float *tflops);
int mppi_measure_copy_bandwidth(int device, size_t bytes, float *gbps);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H_ */
