// C-ABI (include/qmf_b200.h) for dataset ingest on the GPU: raw (user id, item id, value) cells ->
// dense indices + both CSR orientations + longest-first row order, without the host sorts.
// Replaces IdIndex (qmf/utils/IdIndex.h) + WALSEngine::groupSignals / sortDataset
// (qmf/wals/WALSEngine.cpp:130-163), SURVEY.md §8f rank 1:
//   * the dense idx of an id is its rank among the distinct ids (the reference assigns indices while
//     walking the dataset sorted by id: ascending id order)
//   * rows sorted by row id, cells inside a row by column id, duplicates kept (stable: file order)
// The radix sorts and the unique/compaction are CUB device primitives (library plumbing, not a hot
// path); index assignment, key packing and row-pointer construction are the kernels below.
#include "qmfb_common.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cstdint>

namespace qmfb {

// idx[p] = rank of id[p] in the ascending table of distinct ids
__global__ void rank_ids_kernel(const int64_t* __restrict__ id, int64_t n, const int64_t* __restrict__ table, int64_t m,
                                int32_t* __restrict__ idx) {
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v = id[p];
    int64_t lo = 0, hi = m;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (table[mid] < v) lo = mid + 1; else hi = mid;
    }
    idx[p] = int32_t(lo);
  }
}

__global__ void pack_keys_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t n,
                                 uint64_t* __restrict__ key) {
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    key[p] = (uint64_t(uint32_t(row[p])) << 32) | uint32_t(col[p]);
  }
}

__global__ void unpack_cols_kernel(const uint64_t* __restrict__ key, int64_t n, int32_t* __restrict__ col) {
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    col[p] = int32_t(uint32_t(key[p]));
  }
}

// row_ptr[r] = first position whose row is >= r (keys sorted), r in [0, nrows]; len[r] = row length
__global__ void row_ptr_kernel(const uint64_t* __restrict__ key, int64_t n, int64_t nrows, int64_t* __restrict__ row_ptr) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r <= nrows; r += int64_t(gridDim.x) * blockDim.x) {
    const uint64_t target = uint64_t(r) << 32;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (key[mid] < target) lo = mid + 1; else hi = mid;
    }
    row_ptr[r] = lo;
  }
}

__global__ void row_len_kernel(const int64_t* __restrict__ row_ptr, int64_t nrows, uint64_t* __restrict__ len, int32_t* __restrict__ iota) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < nrows; r += int64_t(gridDim.x) * blockDim.x) {
    len[r] = uint64_t(row_ptr[r + 1] - row_ptr[r]);
    iota[r] = int32_t(r);
  }
}

}  // namespace qmfb

using namespace qmfb;

struct qmfb_signals {
  int device = 0;
  int64_t nnz = 0;
  int64_t n[2] = {0, 0};
  int64_t* ids[2] = {nullptr, nullptr};
  int64_t* row_ptr[2] = {nullptr, nullptr};
  int32_t* col[2] = {nullptr, nullptr};
  double* val[2] = {nullptr, nullptr};
  int32_t* order[2] = {nullptr, nullptr};
};

namespace {

struct DevBuf {  // frees on scope exit
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

int dev_alloc(DevBuf& b, size_t bytes) {
  QMFB_CUDA(cudaMalloc(&b.p, std::max<size_t>(bytes, 16)));
  return QMFB_OK;
}

int grid_for(int64_t n) { return int(std::min<int64_t>(148 * 16, std::max<int64_t>(1, (n + 255) / 256))); }

int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
  int b = 1;
  while ((int64_t(1) << b) < n && b < 62) ++b;
  return b;
}

// distinct ascending ids of id[0..n): sorted copy + unique
int distinct_ids(const int64_t* d_id, int64_t n, int64_t** table, int64_t* count) {
  DevBuf sorted, tmp, num, uniq;
  if (int rc = dev_alloc(sorted, size_t(n) * 8)) return rc;
  if (int rc = dev_alloc(uniq, size_t(n) * 8)) return rc;
  if (int rc = dev_alloc(num, 8)) return rc;
  size_t bytes = 0;
  QMFB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_id, sorted.as<int64_t>(), n));
  if (int rc = dev_alloc(tmp, bytes)) return rc;
  QMFB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, d_id, sorted.as<int64_t>(), n));
  size_t bytes2 = 0;
  QMFB_CUDA(cub::DeviceSelect::Unique(nullptr, bytes2, sorted.as<int64_t>(), uniq.as<int64_t>(), num.as<int64_t>(), n));
  DevBuf tmp2;
  if (int rc = dev_alloc(tmp2, bytes2)) return rc;
  QMFB_CUDA(cub::DeviceSelect::Unique(tmp2.p, bytes2, sorted.as<int64_t>(), uniq.as<int64_t>(), num.as<int64_t>(), n));
  QMFB_CUDA(cudaMemcpy(count, num.p, 8, cudaMemcpyDeviceToHost));
  QMFB_CUDA(cudaMalloc(table, std::max<size_t>(size_t(*count) * 8, 16)));
  QMFB_CUDA(cudaMemcpy(*table, uniq.p, size_t(*count) * 8, cudaMemcpyDeviceToDevice));
  return QMFB_OK;
}

// one orientation: rows = ridx, cols = cidx
int build_side(qmfb_signals* s, int side, const int32_t* d_ridx, const int32_t* d_cidx, const double* d_val) {
  const int64_t n = s->nnz, nrows = s->n[side];
  DevBuf key_in, key_out, tmp;
  if (int rc = dev_alloc(key_in, size_t(n) * 8)) return rc;
  if (int rc = dev_alloc(key_out, size_t(n) * 8)) return rc;
  QMFB_CUDA(cudaMalloc(&s->val[side], std::max<size_t>(size_t(n) * 8, 16)));
  QMFB_CUDA(cudaMalloc(&s->col[side], std::max<size_t>(size_t(n) * 4, 16)));
  QMFB_CUDA(cudaMalloc(&s->row_ptr[side], size_t(nrows + 1) * 8));
  QMFB_CUDA(cudaMalloc(&s->order[side], std::max<size_t>(size_t(nrows) * 4, 16)));
  pack_keys_kernel<<<grid_for(n), 256>>>(d_ridx, d_cidx, n, key_in.as<uint64_t>());
  QMFB_CUDA(cudaGetLastError());
  const int end_bit = 32 + bits_for(nrows);
  size_t bytes = 0;
  QMFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.as<uint64_t>(), key_out.as<uint64_t>(), d_val, s->val[side], n, 0, end_bit));
  if (int rc = dev_alloc(tmp, bytes)) return rc;
  QMFB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_in.as<uint64_t>(), key_out.as<uint64_t>(), d_val, s->val[side], n, 0, end_bit));
  unpack_cols_kernel<<<grid_for(n), 256>>>(key_out.as<uint64_t>(), n, s->col[side]);
  QMFB_CUDA(cudaGetLastError());
  row_ptr_kernel<<<grid_for(nrows + 1), 256>>>(key_out.as<uint64_t>(), n, nrows, s->row_ptr[side]);
  QMFB_CUDA(cudaGetLastError());
  // longest rows first (stable: ties by ascending row), the order the persistent solve kernel deals rows in
  DevBuf len_in, len_out, iota, tmp2;
  if (int rc = dev_alloc(len_in, size_t(nrows) * 8)) return rc;
  if (int rc = dev_alloc(len_out, size_t(nrows) * 8)) return rc;
  if (int rc = dev_alloc(iota, size_t(nrows) * 4)) return rc;
  row_len_kernel<<<grid_for(nrows), 256>>>(s->row_ptr[side], nrows, len_in.as<uint64_t>(), iota.as<int32_t>());
  QMFB_CUDA(cudaGetLastError());
  size_t bytes2 = 0;
  QMFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes2, len_in.as<uint64_t>(), len_out.as<uint64_t>(), iota.as<int32_t>(),
                                                      s->order[side], nrows, 0, bits_for(n + 1)));
  if (int rc = dev_alloc(tmp2, bytes2)) return rc;
  QMFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp2.p, bytes2, len_in.as<uint64_t>(), len_out.as<uint64_t>(), iota.as<int32_t>(),
                                                      s->order[side], nrows, 0, bits_for(n + 1)));
  QMFB_CUDA(cudaDeviceSynchronize());
  return QMFB_OK;
}

}  // namespace

namespace qmfb {

__global__ void rebase_row_ptr_kernel(const int64_t* __restrict__ row_ptr, int64_t row_begin, int64_t nrows, int64_t* __restrict__ out) {
  const int64_t base = row_ptr[row_begin];
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r <= nrows; r += int64_t(gridDim.x) * blockDim.x) {
    out[r] = row_ptr[row_begin + r] - base;
  }
}

int rebase_row_ptr(cudaStream_t st, const int64_t* row_ptr, int64_t row_begin, int64_t nrows, int64_t* out) {
  rebase_row_ptr_kernel<<<grid_for(nrows + 1), 256, 0, st>>>(row_ptr, row_begin, nrows, out);
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

int longest_first_order(const int64_t* row_ptr_local, int64_t nrows, int64_t nnz, int32_t* order) {
  if (nrows < 1) return QMFB_OK;
  DevBuf len_in, len_out, iota, tmp;
  if (int rc = dev_alloc(len_in, size_t(nrows) * 8)) return rc;
  if (int rc = dev_alloc(len_out, size_t(nrows) * 8)) return rc;
  if (int rc = dev_alloc(iota, size_t(nrows) * 4)) return rc;
  row_len_kernel<<<grid_for(nrows), 256>>>(row_ptr_local, nrows, len_in.as<uint64_t>(), iota.as<int32_t>());
  QMFB_CUDA(cudaGetLastError());
  size_t bytes = 0;
  QMFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, len_in.as<uint64_t>(), len_out.as<uint64_t>(), iota.as<int32_t>(), order,
                                                      nrows, 0, bits_for(nnz + 1)));
  if (int rc = dev_alloc(tmp, bytes)) return rc;
  QMFB_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp.p, bytes, len_in.as<uint64_t>(), len_out.as<uint64_t>(), iota.as<int32_t>(), order,
                                                      nrows, 0, bits_for(nnz + 1)));
  QMFB_CUDA(cudaDeviceSynchronize());
  return QMFB_OK;
}

}  // namespace qmfb

extern "C" {

int qmfb_signals_destroy(qmfb_signals_t* s) {
  if (!s) return QMFB_OK;
  cudaSetDevice(s->device);
  for (int side = 0; side < 2; ++side) {
    cudaFree(s->ids[side]);
    cudaFree(s->row_ptr[side]);
    cudaFree(s->col[side]);
    cudaFree(s->val[side]);
    cudaFree(s->order[side]);
  }
  delete s;
  return QMFB_OK;
}

int qmfb_signals_build(int device, int64_t nnz, const int64_t* user_ids, const int64_t* item_ids, const double* values,
                       qmfb_signals_t** out) {
  if (!out || nnz < 1 || !user_ids || !item_ids || !values) return set_error(QMFB_ERR_INVALID, "qmfb_signals_build: bad argument");
  QMFB_CUDA(cudaSetDevice(device));
  auto* s = new qmfb_signals;
  s->device = device;
  s->nnz = nnz;
  auto fail = [&](int rc) {
    qmfb_signals_destroy(s);
    return rc;
  };
  DevBuf d_uid, d_iid, d_val, d_uidx, d_iidx;
  int rc;
  if ((rc = dev_alloc(d_uid, size_t(nnz) * 8)) || (rc = dev_alloc(d_iid, size_t(nnz) * 8)) || (rc = dev_alloc(d_val, size_t(nnz) * 8)) ||
      (rc = dev_alloc(d_uidx, size_t(nnz) * 4)) || (rc = dev_alloc(d_iidx, size_t(nnz) * 4))) {
    return fail(rc);
  }
  if (cudaMemcpy(d_uid.p, user_ids, size_t(nnz) * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_iid.p, item_ids, size_t(nnz) * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_val.p, values, size_t(nnz) * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
    return fail(set_error(QMFB_ERR_CUDA, "qmfb_signals_build: upload failed: %s", cudaGetErrorString(cudaGetLastError())));
  }
  if ((rc = distinct_ids(d_uid.as<int64_t>(), nnz, &s->ids[0], &s->n[0])) || (rc = distinct_ids(d_iid.as<int64_t>(), nnz, &s->ids[1], &s->n[1]))) {
    return fail(rc);
  }
  if (s->n[0] > INT32_MAX || s->n[1] > INT32_MAX) return fail(set_error(QMFB_ERR_UNSUPPORTED, "more than 2^31-1 distinct users or items"));
  rank_ids_kernel<<<grid_for(nnz), 256>>>(d_uid.as<int64_t>(), nnz, s->ids[0], s->n[0], d_uidx.as<int32_t>());
  rank_ids_kernel<<<grid_for(nnz), 256>>>(d_iid.as<int64_t>(), nnz, s->ids[1], s->n[1], d_iidx.as<int32_t>());
  if (cudaGetLastError() != cudaSuccess) return fail(set_error(QMFB_ERR_CUDA, "rank_ids_kernel launch failed"));
  cudaFree(d_uid.p); d_uid.p = nullptr;  // the raw ids are no longer needed: make room for the sort buffers
  cudaFree(d_iid.p); d_iid.p = nullptr;
  if ((rc = build_side(s, 0, d_uidx.as<int32_t>(), d_iidx.as<int32_t>(), d_val.as<double>())) ||
      (rc = build_side(s, 1, d_iidx.as<int32_t>(), d_uidx.as<int32_t>(), d_val.as<double>()))) {
    return fail(rc);
  }
  *out = s;
  return QMFB_OK;
}

int qmfb_signals_device_ordinal(const qmfb_signals_t* s) { return s ? s->device : -1; }

int qmfb_signals_dims(const qmfb_signals_t* s, int64_t* nusers, int64_t* nitems, int64_t* nnz) {
  if (!s) return set_error(QMFB_ERR_INVALID, "qmfb_signals_dims: null handle");
  if (nusers) *nusers = s->n[0];
  if (nitems) *nitems = s->n[1];
  if (nnz) *nnz = s->nnz;
  return QMFB_OK;
}

int qmfb_signals_ids(const qmfb_signals_t* s, int side, int64_t* ids_host) {
  if (!s || side < 0 || side > 1 || !ids_host) return set_error(QMFB_ERR_INVALID, "qmfb_signals_ids: bad argument");
  QMFB_CUDA(cudaSetDevice(s->device));
  QMFB_CUDA(cudaMemcpy(ids_host, s->ids[side], size_t(s->n[side]) * 8, cudaMemcpyDeviceToHost));
  return QMFB_OK;
}

int qmfb_signals_csr(const qmfb_signals_t* s, int side, int64_t* row_ptr, int32_t* col, double* val, int32_t* order) {
  if (!s || side < 0 || side > 1) return set_error(QMFB_ERR_INVALID, "qmfb_signals_csr: bad argument");
  QMFB_CUDA(cudaSetDevice(s->device));
  if (row_ptr) QMFB_CUDA(cudaMemcpy(row_ptr, s->row_ptr[side], size_t(s->n[side] + 1) * 8, cudaMemcpyDeviceToHost));
  if (col) QMFB_CUDA(cudaMemcpy(col, s->col[side], size_t(s->nnz) * 4, cudaMemcpyDeviceToHost));
  if (val) QMFB_CUDA(cudaMemcpy(val, s->val[side], size_t(s->nnz) * 8, cudaMemcpyDeviceToHost));
  if (order) QMFB_CUDA(cudaMemcpy(order, s->order[side], size_t(s->n[side]) * 4, cudaMemcpyDeviceToHost));
  return QMFB_OK;
}

int qmfb_signals_device(const qmfb_signals_t* s, int side, const int64_t** row_ptr, const int32_t** col, const double** val,
                        const int32_t** order) {
  if (!s || side < 0 || side > 1) return set_error(QMFB_ERR_INVALID, "qmfb_signals_device: bad argument");
  if (row_ptr) *row_ptr = s->row_ptr[side];
  if (col) *col = s->col[side];
  if (val) *val = s->val[side];
  if (order) *order = s->order[side];
  return QMFB_OK;
}

}  // extern "C"
