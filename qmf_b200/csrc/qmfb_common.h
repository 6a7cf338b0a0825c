// Shared plumbing for the C-ABI translation units: error reporting and CUDA call checking.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/qmf_b200.h"

namespace qmfb {

int set_error(int code, const char* fmt, ...);

#define QMFB_CUDA(expr)                                                                                       \
  do {                                                                                                        \
    cudaError_t qmfb_e_ = (expr);                                                                             \
    if (qmfb_e_ != cudaSuccess) {                                                                             \
      return ::qmfb::set_error(QMFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(qmfb_e_), \
                               __FILE__, __LINE__);                                                           \
    }                                                                                                         \
  } while (0)

}  // namespace qmfb
