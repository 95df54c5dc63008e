// gpu_err_chk.h -- error convention of the reference (PI/gpu_err_chk.h:19-26,49): HANDLE_ERROR
// prints and CONTINUES.  Accepts both cudaError_t values and the library's negative MPPI_ERR_* codes.
#ifndef HANDLE_ERROR_H_
#define HANDLE_ERROR_H_
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../mppi_b200.h"

inline void gpuAssert(int code, const char *file, int line, bool abort = true) {
  if (code != 0) {
    fprintf(stderr, "GPUassert: %s %s %d\n", mppi_error_string(code), file, line);
    if (abort) exit(code);
  }
}

#define HANDLE_ERROR(ans) { gpuAssert((int)(ans), __FILE__, __LINE__, false); }
#endif
