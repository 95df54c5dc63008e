// managed.cuh -- Managed: the base class the dynamics and cost types derive from; it records which CUDA stream an
// object was bound to (API shape of PI/managed.cuh:30-45: public `stream_`, `bindToStream`).
//
// The reference calls cudaDeviceSynchronize() on every bind.  Here the parameters of models and costs stay on the host
// until a controller uploads them on its own context stream, so binding is pure bookkeeping: nothing to wait for.
#pragma once
#include <cuda_runtime.h>

namespace autorally_control {

struct Managed {
  cudaStream_t stream_ = nullptr;  // nullptr = "use the controller context's own stream"

  void bindToStream(cudaStream_t stream) noexcept { stream_ = stream; }
  cudaStream_t boundStream() const noexcept { return stream_; }
};

}  // namespace autorally_control
