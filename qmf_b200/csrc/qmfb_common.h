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

// ---- helpers shared between translation units (qmfb_ingest.cu) ----
// out[r] = row_ptr[row_begin + r] - row_ptr[row_begin], r = 0 .. nrows (device arrays, asynchronous on st)
int rebase_row_ptr(cudaStream_t st, const int64_t* row_ptr, int64_t row_begin, int64_t nrows, int64_t* out);
// order = local rows longest first (stable), the order the persistent solve kernel deals rows in;
// device arrays, current device, synchronous
int longest_first_order(const int64_t* row_ptr_local, int64_t nrows, int64_t nnz, int32_t* order);
// out[0] = sum of v[0..n) in a fixed order (single block, fixed strides + fixed tree), asynchronous (qmfb_wals.cu)
int det_sum_launch(cudaStream_t st, const double* v, int64_t n, double* out);

// ---- ranking evaluation on factors that are already on a device (qmfb_eval.cu) ----
// host test-user arrays in, host counters out; U / V / bias are device pointers on `device` (any row stride:
// repacked to padded rows when needed).  label_ptr may be a slice of a larger test set (label_ptr[0] != 0):
// label_items is then the FULL array and cnt / pos_scores point at the slice's first entry.  Synchronous.
int eval_rank_resident(int device, const double* U, int64_t ldu, int64_t nusers, const double* V, int64_t ldv, int64_t nitems, int k,
                       const double* bias, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                       const int32_t* label_items, int32_t* cnt, double* pos_scores);
// the test users cut into ndev contiguous slices, slice d scored on devices[d] against replicas U[d], V[d]
int eval_rank_sharded(int ndev, const int* devices, const double* const* U, int64_t ldu, int64_t nusers, const double* const* V, int64_t ldv,
                      int64_t nitems, int k, const int32_t* test_users, int64_t nT, const int64_t* label_ptr, const int32_t* label_items,
                      int32_t* cnt, double* pos_scores);

}  // namespace qmfb
