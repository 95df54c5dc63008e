// rollout_launch.h -- host-side launchers of the rollout-kernel instantiations.  Each lives in its own
// translation unit (rollout_*.cu) so the variants compile in parallel.
#pragma once
#include "device_common.cuh"

namespace mppi {

// `small` selects 32-thread CTAs so that a few thousand rollouts still spread over many SMs.
cudaError_t launch_rollout_nn32_r1(const RolloutParams &p, cudaStream_t st, bool small);
cudaError_t launch_rollout_nn32_r2(const RolloutParams &p, cudaStream_t st, bool small);
// pdl: launch with programmatic stream serialization (the kernel overlaps its prologue with its predecessor's tail)
cudaError_t launch_rollout_nn32_half(const RolloutParams &p, cudaStream_t st, bool pdl);
cudaError_t launch_rollout_nn32_warp(const RolloutParams &p, cudaStream_t st, bool pdl);  // one rollout per warp (<= 2368 rollouts)
// tcgen05 tensor-core MLP (FP16 hi/lo split, activations in tensor memory): the filled-GPU kernel
cudaError_t launch_rollout_nn32_tc(const RolloutParams &p, cudaStream_t st, const float *host_theta_t, bool pdl);
cudaError_t launch_rollout_nn64_tc(const RolloutParams &p, cudaStream_t st, const float *host_theta_t, bool pdl);  // 6-64-64-64-64-4
bool tc_biases_in_range(const float *host_theta_t, int hid, int nhid);
cudaError_t launch_rollout_nn64_r1(const RolloutParams &p, cudaStream_t st, bool small);
// 6-64-64-64-64-4 at controller sizes: hidden layers in the registers of one warp each, rollouts flowing through them
cudaError_t launch_rollout_nn64_pipe(const RolloutParams &p, cudaStream_t st);
bool rollout_pipe64_fits(int T);
cudaError_t launch_rollout_bf(const RolloutParams &p, cudaStream_t st, bool small);
bool rollout_bf_is_split(long long total);
// run-time layer pack (any NeuralNetModel<7,2,3,6,...,4> with widths <= 128): rollout_generic.cu
cudaError_t launch_rollout_generic(const RolloutParams &p, cudaStream_t st, const int *net_structure, int num_layers);

}  // namespace mppi
