// managed.cuh -- stream-binding base of the dynamics and cost classes (API of PI/managed.cuh:30-45).
#ifndef MPPI_MANAGED_CUH_
#define MPPI_MANAGED_CUH_
#include <cuda_runtime.h>

namespace autorally_control {

class Managed {
 public:
  cudaStream_t stream_ = 0;  ///< stream the object is bound to (0 = the library's own stream)

  // The reference synchronises the whole device here (PI/managed.cuh:42).  Parameters of this
  // implementation live on the host until a controller uploads them on its own stream, so
  // binding is pure bookkeeping and needs no device-wide synchronisation.
  void bindToStream(cudaStream_t stream) { stream_ = stream; }
};

}  // namespace autorally_control
#endif
