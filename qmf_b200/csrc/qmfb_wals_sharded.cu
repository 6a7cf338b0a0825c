// C-ABI (include/qmf_b200.h): WALS across the GPUs of ONE box from ONE process - the engine the C++
// `wals --ngpus N` binary binds (north_star: "the host side stays C++ ... WALS runs across the 8 GPUs").
// Replaces the loop of WALSEngine::iterate (qmf/wals/WALSEngine.cpp:165-218) for both sides.
//
// Sharding: users and items are each cut into ndev contiguous row ranges balanced by
// nnz; every device holds full replicas of both factor matrices, the CSR rows of its user range and
// the CSC rows of its item range.  One half-step "update side S" on device d:
//   1. wait for every device's previous solve (their peer stores filled d's replica of the fixed side)
//   2. partial Gram of the fixed side: device d sums Gram parts [P d/N, P (d+1)/N) of the P FIXED
//      row parts (gram_partial_kernel) from its own replica
//   3. every device sums all P parts in part order, reading the other devices' partials over NVLink
//      peer memory (gram_reduce_kernel): no library collective, and - because the parts and the order
//      do not depend on N - the Gram, hence every factor, is BIT-IDENTICAL to the single-GPU result
//   4. solve the rows of its shard; the kernel stores every solved row into all replicas
//      (qmfb_wals_solve_peers_dev: the all-gather is fused into the solve kernel as peer stores)
//   5. the per-row loss terms are gathered on device 0 and summed there in row order (same order as
//      one GPU -> same bits)
// Ordering between devices is CUDA events only (cudaStreamWaitEvent across devices); one host thread
// issues the work of all devices asynchronously.
#include "qmfb_common.h"

#include <algorithm>
#include <numeric>
#include <vector>

using namespace qmfb;

namespace {

struct Shard {
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  double* F[2] = {nullptr, nullptr};
  int64_t row_begin[2] = {0, 0}, nrows[2] = {0, 0}, nnz[2] = {0, 0};
  int64_t* row_ptr[2] = {nullptr, nullptr};
  int32_t* col[2] = {nullptr, nullptr};
  double* val[2] = {nullptr, nullptr};
  int32_t* order[2] = {nullptr, nullptr};
  double *gram_ws = nullptr, *gram_packed = nullptr, *row_loss = nullptr, *loss_sum = nullptr;
  int32_t* scratch = nullptr;
  cudaEvent_t ev_gram = nullptr, ev_solve = nullptr, ev_aux = nullptr, ev_copy = nullptr;
  int part_begin[2] = {0, 0}, part_end[2] = {0, 0};  // Gram parts of side s summed by this device
};

}  // namespace

struct qmfb_wals_sharded {
  int ndev = 0;
  std::vector<Shard> sh;
  int64_t n[2] = {0, 0};
  int k = 0, kp = 0;
  bool has_csr[2] = {false, false};
  double* row_loss_full = nullptr;  // device sh[0]: every row's loss term, global row order
  double* loss_total = nullptr;
  cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
  float gram_ms = 0.f, solve_ms = 0.f;
  int64_t launches = 0;
  int32_t* err_host = nullptr;  // pinned: 2 ints per device and half-step of the running call (a pageable
                                // destination would make every cudaMemcpyAsync block the issuing host thread)
};

namespace {

int sharded_release(qmfb_wals_sharded* h) {
  for (auto& s : h->sh) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.copy_stream) cudaStreamSynchronize(s.copy_stream);
  }
  for (auto& s : h->sh) {
    cudaSetDevice(s.device);
    for (int side = 0; side < 2; ++side) {
      cudaFree(s.F[side]);
      cudaFree(s.row_ptr[side]);
      cudaFree(s.col[side]);
      cudaFree(s.val[side]);
      cudaFree(s.order[side]);
    }
    cudaFree(s.gram_ws);
    cudaFree(s.gram_packed);
    cudaFree(s.row_loss);
    cudaFree(s.loss_sum);
    cudaFree(s.scratch);
    for (cudaEvent_t e : {s.ev_gram, s.ev_solve, s.ev_aux, s.ev_copy}) {
      if (e) cudaEventDestroy(e);
    }
    if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (!h->sh.empty()) {
    cudaSetDevice(h->sh[0].device);
    cudaFree(h->row_loss_full);
    cudaFree(h->loss_total);
    cudaFreeHost(h->err_host);
    for (auto& e : h->tev) {
      if (e) cudaEventDestroy(e);
    }
  }
  delete h;
  return QMFB_OK;
}

void free_side(Shard& s, int side) {
  cudaFree(s.row_ptr[side]);
  cudaFree(s.col[side]);
  cudaFree(s.val[side]);
  cudaFree(s.order[side]);
  s.row_ptr[side] = nullptr;
  s.col[side] = nullptr;
  s.val[side] = nullptr;
  s.order[side] = nullptr;
}

// contiguous row ranges with ~equal nnz: cut r = first row whose prefix reaches nnz * r / ndev
std::vector<int64_t> balanced_cuts(const int64_t* row_ptr, int64_t nrows, int ndev) {
  std::vector<int64_t> cuts(size_t(ndev) + 1, 0);
  const int64_t nnz = row_ptr[nrows];
  for (int r = 1; r < ndev; ++r) {
    const int64_t target = int64_t((__int128)nnz * r / ndev);
    const int64_t c = std::lower_bound(row_ptr, row_ptr + nrows + 1, target) - row_ptr;
    cuts[size_t(r)] = std::min(std::max(c, cuts[size_t(r) - 1]), nrows);
  }
  cuts[size_t(ndev)] = nrows;
  return cuts;
}

int alloc_side(Shard& s, int side, int64_t nrows, int64_t nnz) {
  QMFB_CUDA(cudaMalloc(&s.row_ptr[side], size_t(nrows + 1) * sizeof(int64_t)));
  QMFB_CUDA(cudaMalloc(&s.col[side], size_t(std::max<int64_t>(nnz, 1)) * sizeof(int32_t)));
  QMFB_CUDA(cudaMalloc(&s.val[side], size_t(std::max<int64_t>(nnz, 1)) * sizeof(double)));
  QMFB_CUDA(cudaMalloc(&s.order[side], size_t(std::max<int64_t>(nrows, 1)) * sizeof(int32_t)));
  return QMFB_OK;
}

// issue one half-step on every device (asynchronous); `slot` selects where the error flags of this
// half-step are copied in h->err_host
int half_step_async(qmfb_wals_sharded* h, int side, double alpha, double lambda, int slot) {
  const int other = 1 - side, N = h->ndev;
  if (!h->has_csr[side]) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded: no CSR uploaded for side %d", side);
  std::vector<const double*> ws(static_cast<size_t>(N), nullptr);
  std::vector<int> part_end(static_cast<size_t>(N), 0);
  for (int d = 0; d < N; ++d) {
    ws[size_t(d)] = h->sh[size_t(d)].gram_ws;
    part_end[size_t(d)] = h->sh[size_t(d)].part_end[other];
  }
  for (int d = 0; d < N; ++d) {
    Shard& s = h->sh[size_t(d)];
    QMFB_CUDA(cudaSetDevice(s.device));
    for (int e = 0; e < N; ++e) {
      if (e != d) QMFB_CUDA(cudaStreamWaitEvent(s.stream, h->sh[size_t(e)].ev_solve, 0));
    }
    // leftData.setFactors(0), WALSEngine.cpp:170-171 (this shard's rows; the others arrive as peer stores)
    QMFB_CUDA(cudaMemsetAsync(s.F[side] + s.row_begin[side] * h->kp, 0, size_t(s.nrows[side]) * h->kp * sizeof(double), s.stream));
    if (d == 0) QMFB_CUDA(cudaEventRecord(h->tev[0], s.stream));
    if (int rc = qmfb_gram_parts_dev(s.stream, s.F[other], h->kp, 0, h->n[other], h->k, s.part_begin[other], s.part_end[other], s.gram_ws)) return rc;
    if (s.part_end[other] > s.part_begin[other]) ++h->launches;
    QMFB_CUDA(cudaEventRecord(s.ev_gram, s.stream));
  }
  for (int d = 0; d < N; ++d) {
    Shard& s = h->sh[size_t(d)];
    QMFB_CUDA(cudaSetDevice(s.device));
    for (int e = 0; e < N; ++e) {
      if (e != d) QMFB_CUDA(cudaStreamWaitEvent(s.stream, h->sh[size_t(e)].ev_gram, 0));
    }
    if (int rc = qmfb_gram_reduce_parts_dev(s.stream, ws.data(), part_end.data(), N, h->k, s.gram_packed)) return rc;
    if (d == 0) QMFB_CUDA(cudaEventRecord(h->tev[1], s.stream));
    double* peers[16];
    int np = 0;
    for (int e = 0; e < N; ++e) {
      if (e != d) peers[np++] = h->sh[size_t(e)].F[side];
    }
    if (int rc = qmfb_wals_solve_peers_dev(s.stream, s.F[side], h->kp, s.row_begin[side], s.F[other], h->kp, h->k, s.row_ptr[side],
                                           s.col[side], s.val[side], s.order[side], s.nrows[side], s.nnz[side], s.gram_packed, alpha, lambda,
                                           s.row_loss, s.loss_sum, s.scratch, peers, np)) return rc;
    if (d == 0) QMFB_CUDA(cudaEventRecord(h->tev[2], s.stream));
    h->launches += 2 + ((s.nrows[side] > 0) ? (h->kp <= 128 ? 4 : 1) : 0);  // reduce, sum, [long plan, long partial, long reduce,] solve
    QMFB_CUDA(cudaMemcpyAsync(h->err_host + size_t(slot * N + d) * 2, s.scratch, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
    // loss terms of this shard -> device 0's full array (global row order)
    if (s.nrows[side] > 0) {
      QMFB_CUDA(cudaMemcpyPeerAsync(h->row_loss_full + s.row_begin[side], h->sh[0].device, s.row_loss, s.device,
                                    size_t(s.nrows[side]) * sizeof(double), s.stream));
    }
    QMFB_CUDA(cudaEventRecord(s.ev_solve, s.stream));
    QMFB_CUDA(cudaEventRecord(s.ev_aux, s.stream));
  }
  {  // device 0: deterministic sum of all row losses in global row order (same order as one GPU)
    Shard& s0 = h->sh[0];
    QMFB_CUDA(cudaSetDevice(s0.device));
    for (int e = 1; e < N; ++e) QMFB_CUDA(cudaStreamWaitEvent(s0.stream, h->sh[size_t(e)].ev_aux, 0));
    if (int rc = det_sum_launch(s0.stream, h->row_loss_full, h->n[side], h->loss_total)) return rc;
    ++h->launches;
  }
  return QMFB_OK;
}

int finish(qmfb_wals_sharded* h, int nslots, double* loss_sum) {
  double loss = 0.0;
  Shard& s0 = h->sh[0];
  QMFB_CUDA(cudaSetDevice(s0.device));
  QMFB_CUDA(cudaMemcpyAsync(&loss, h->loss_total, sizeof(double), cudaMemcpyDeviceToHost, s0.stream));
  for (auto& s : h->sh) {
    QMFB_CUDA(cudaSetDevice(s.device));
    QMFB_CUDA(cudaStreamSynchronize(s.stream));
    QMFB_CUDA(cudaStreamSynchronize(s.copy_stream));
  }
  QMFB_CUDA(cudaSetDevice(s0.device));
  QMFB_CUDA(cudaEventElapsedTime(&h->gram_ms, h->tev[0], h->tev[1]));
  QMFB_CUDA(cudaEventElapsedTime(&h->solve_ms, h->tev[1], h->tev[2]));
  for (int i = 0; i < nslots * h->ndev; ++i) {
    if (h->err_host[size_t(i) * 2 + 1] != 0) {
      return set_error(QMFB_ERR_NOT_SPD, "normal equations not positive definite (reference: dsysv failed) on device slot %d, half-step %d",
                       i % h->ndev, i / h->ndev);
    }
  }
  if (loss_sum) *loss_sum = loss;
  return QMFB_OK;
}

}  // namespace

extern "C" {

int qmfb_wals_sharded_create(int ndev, const int* dev_ids, int64_t nusers, int64_t nitems, int nfactors,
                             qmfb_wals_sharded_t** out) {
  if (!out || !dev_ids || ndev < 1 || ndev > 16 || nusers < 1 || nitems < 1) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_create: bad argument");
  const int kp = qmfb_padded_k(nfactors);
  if (kp < 0) return kp;
  int count = 0;
  QMFB_CUDA(cudaGetDeviceCount(&count));
  for (int d = 0; d < ndev; ++d) {
    if (dev_ids[d] < 0 || dev_ids[d] >= count) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_create: device %d does not exist (%d visible)", dev_ids[d], count);
  }
  // every device stores into every other device's replicas: peer access both ways
  for (int d = 0; d < ndev; ++d) {
    QMFB_CUDA(cudaSetDevice(dev_ids[d]));
    for (int e = 0; e < ndev; ++e) {
      if (dev_ids[e] == dev_ids[d]) continue;
      int can = 0;
      QMFB_CUDA(cudaDeviceCanAccessPeer(&can, dev_ids[d], dev_ids[e]));
      if (!can) return set_error(QMFB_ERR_UNSUPPORTED, "device %d cannot access device %d as a peer", dev_ids[d], dev_ids[e]);
      const cudaError_t pe = cudaDeviceEnablePeerAccess(dev_ids[e], 0);
      if (pe == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
      } else {
        QMFB_CUDA(pe);
      }
    }
  }
  auto* h = new qmfb_wals_sharded;
  h->ndev = ndev;
  h->n[0] = nusers;
  h->n[1] = nitems;
  h->k = nfactors;
  h->kp = kp;
  h->sh.resize(size_t(ndev));
  const int rc = [&]() -> int {
    const int parts[2] = {qmfb_gram_parts_count(nusers, nfactors), qmfb_gram_parts_count(nitems, nfactors)};
    for (int d = 0; d < ndev; ++d) {
      Shard& s = h->sh[size_t(d)];
      s.device = dev_ids[d];
      QMFB_CUDA(cudaSetDevice(s.device));
      QMFB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      QMFB_CUDA(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
      for (cudaEvent_t* e : {&s.ev_gram, &s.ev_solve, &s.ev_aux, &s.ev_copy}) QMFB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
      for (int side = 0; side < 2; ++side) {
        QMFB_CUDA(cudaMalloc(&s.F[side], size_t(h->n[side]) * kp * sizeof(double)));
        QMFB_CUDA(cudaMemsetAsync(s.F[side], 0, size_t(h->n[side]) * kp * sizeof(double), s.stream));
        s.part_begin[side] = int(int64_t(parts[side]) * d / ndev);
        s.part_end[side] = int(int64_t(parts[side]) * (d + 1) / ndev);
      }
      QMFB_CUDA(cudaMalloc(&s.gram_ws, size_t(qmfb_gram_workspace_len(nfactors)) * sizeof(double)));
      QMFB_CUDA(cudaMalloc(&s.gram_packed, size_t(qmfb_gram_packed_len(nfactors)) * sizeof(double)));
      QMFB_CUDA(cudaMalloc(&s.row_loss, size_t(std::max(nusers, nitems)) * sizeof(double)));
      QMFB_CUDA(cudaMalloc(&s.loss_sum, sizeof(double)));
      QMFB_CUDA(cudaMalloc(&s.scratch, 2 * sizeof(int32_t)));
      QMFB_CUDA(cudaEventRecord(s.ev_solve, s.stream));  // "previous solve" of the first half-step
      QMFB_CUDA(cudaEventRecord(s.ev_gram, s.stream));
      QMFB_CUDA(cudaEventRecord(s.ev_aux, s.stream));
    }
    QMFB_CUDA(cudaSetDevice(h->sh[0].device));
    QMFB_CUDA(cudaMallocHost(&h->err_host, size_t(ndev) * 4 * sizeof(int32_t)));
    for (int i = 0; i < ndev * 4; ++i) h->err_host[i] = 0;
    QMFB_CUDA(cudaMalloc(&h->row_loss_full, size_t(std::max(nusers, nitems)) * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->loss_total, sizeof(double)));
    for (auto& e : h->tev) QMFB_CUDA(cudaEventCreate(&e));
    for (auto& s : h->sh) {
      QMFB_CUDA(cudaSetDevice(s.device));
      QMFB_CUDA(cudaStreamSynchronize(s.stream));
    }
    return QMFB_OK;
  }();
  if (rc != QMFB_OK) {
    sharded_release(h);
    return rc;
  }
  *out = h;
  return QMFB_OK;
}

int qmfb_wals_sharded_destroy(qmfb_wals_sharded_t* h) { return h ? sharded_release(h) : QMFB_OK; }

int qmfb_wals_sharded_ndev(const qmfb_wals_sharded_t* h) { return h ? h->ndev : 0; }

int qmfb_wals_sharded_shard(const qmfb_wals_sharded_t* h, int slot, int side, int* device, int64_t* row_begin, int64_t* nrows,
                            int64_t* nnz) {
  if (!h || slot < 0 || slot >= h->ndev || side < 0 || side > 1) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_shard: bad argument");
  const Shard& s = h->sh[size_t(slot)];
  if (device) *device = s.device;
  if (row_begin) *row_begin = s.row_begin[side];
  if (nrows) *nrows = s.nrows[side];
  if (nnz) *nnz = s.nnz[side];
  return QMFB_OK;
}

int qmfb_wals_sharded_set_csr(qmfb_wals_sharded_t* h, int side, const int64_t* row_ptr, const int32_t* col_idx, const double* val) {
  if (!h || side < 0 || side > 1 || !row_ptr || row_ptr[0] != 0) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_csr: bad argument");
  const int64_t nrows = h->n[side], ncols = h->n[1 - side], nnz = row_ptr[nrows];
  if (nnz > 0 && (!col_idx || !val)) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_csr: null col/val");
  for (int64_t r = 0; r < nrows; ++r) {
    if (row_ptr[r + 1] < row_ptr[r]) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_csr: row_ptr not monotone at %lld", (long long)r);
  }
  for (int64_t p = 0; p < nnz; ++p) {
    if (col_idx[p] < 0 || col_idx[p] >= ncols) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_csr: col_idx[%lld]=%d out of range", (long long)p, col_idx[p]);
  }
  const std::vector<int64_t> cuts = balanced_cuts(row_ptr, nrows, h->ndev);
  for (int d = 0; d < h->ndev; ++d) {
    Shard& s = h->sh[size_t(d)];
    QMFB_CUDA(cudaSetDevice(s.device));
    QMFB_CUDA(cudaStreamSynchronize(s.stream));
    free_side(s, side);
    const int64_t b = cuts[size_t(d)], e = cuts[size_t(d) + 1], p0 = row_ptr[b], p1 = row_ptr[e];
    if (int rc = alloc_side(s, side, e - b, p1 - p0)) return rc;
    std::vector<int64_t> lrp(size_t(e - b) + 1);
    for (int64_t r = b; r <= e; ++r) lrp[size_t(r - b)] = row_ptr[r] - p0;
    std::vector<int32_t> order(size_t(e - b));
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return lrp[size_t(x) + 1] - lrp[size_t(x)] > lrp[size_t(y) + 1] - lrp[size_t(y)]; });
    QMFB_CUDA(cudaMemcpy(s.row_ptr[side], lrp.data(), lrp.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    if (p1 > p0) {
      QMFB_CUDA(cudaMemcpy(s.col[side], col_idx + p0, size_t(p1 - p0) * sizeof(int32_t), cudaMemcpyHostToDevice));
      QMFB_CUDA(cudaMemcpy(s.val[side], val + p0, size_t(p1 - p0) * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (e > b) QMFB_CUDA(cudaMemcpy(s.order[side], order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    s.row_begin[side] = b;
    s.nrows[side] = e - b;
    s.nnz[side] = p1 - p0;
  }
  h->has_csr[side] = true;
  return QMFB_OK;
}

int qmfb_wals_sharded_set_signals(qmfb_wals_sharded_t* h, const qmfb_signals_t* sig) {
  int64_t nu = 0, ni = 0, nnz = 0;
  if (!h || !sig || qmfb_signals_dims(sig, &nu, &ni, &nnz) != QMFB_OK) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_signals: bad argument");
  if (nu != h->n[0] || ni != h->n[1]) {
    return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_signals: engine is %lld x %lld, signals are %lld x %lld", (long long)h->n[0],
                     (long long)h->n[1], (long long)nu, (long long)ni);
  }
  const int src_dev = qmfb_signals_device_ordinal(sig);
  for (int side = 0; side < 2; ++side) {
    const int64_t nrows = h->n[side];
    const int64_t* rp = nullptr;
    const int32_t* col = nullptr;
    const double* val = nullptr;
    if (int rc = qmfb_signals_device(sig, side, &rp, &col, &val, nullptr)) return rc;
    std::vector<int64_t> host_rp(size_t(nrows) + 1);
    if (int rc = qmfb_signals_csr(sig, side, host_rp.data(), nullptr, nullptr, nullptr)) return rc;
    const std::vector<int64_t> cuts = balanced_cuts(host_rp.data(), nrows, h->ndev);
    for (int d = 0; d < h->ndev; ++d) {
      Shard& s = h->sh[size_t(d)];
      QMFB_CUDA(cudaSetDevice(s.device));
      QMFB_CUDA(cudaStreamSynchronize(s.stream));
      free_side(s, side);
      const int64_t b = cuts[size_t(d)], e = cuts[size_t(d) + 1], p0 = host_rp[size_t(b)], p1 = host_rp[size_t(e)];
      if (int rc = alloc_side(s, side, e - b, p1 - p0)) return rc;
      int64_t* tmp = nullptr;  // this shard's slice of the global row pointers
      QMFB_CUDA(cudaMalloc(&tmp, size_t(e - b + 1) * sizeof(int64_t)));
      cudaError_t ce = cudaMemcpyPeerAsync(tmp, s.device, rp + b, src_dev, size_t(e - b + 1) * sizeof(int64_t), s.stream);
      if (ce == cudaSuccess && p1 > p0) ce = cudaMemcpyPeerAsync(s.col[side], s.device, col + p0, src_dev, size_t(p1 - p0) * sizeof(int32_t), s.stream);
      if (ce == cudaSuccess && p1 > p0) ce = cudaMemcpyPeerAsync(s.val[side], s.device, val + p0, src_dev, size_t(p1 - p0) * sizeof(double), s.stream);
      int rc = QMFB_OK;
      if (ce == cudaSuccess) rc = rebase_row_ptr(s.stream, tmp, 0, e - b, s.row_ptr[side]);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(s.stream);
      cudaFree(tmp);
      if (rc) return rc;
      QMFB_CUDA(ce);
      if (int rc2 = longest_first_order(s.row_ptr[side], e - b, p1 - p0, s.order[side])) return rc2;
      s.row_begin[side] = b;
      s.nrows[side] = e - b;
      s.nnz[side] = p1 - p0;
    }
    h->has_csr[side] = true;
  }
  return QMFB_OK;
}

int qmfb_wals_sharded_set_factors(qmfb_wals_sharded_t* h, int side, const double* host) {
  if (!h || side < 0 || side > 1 || !host) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_set_factors: bad argument");
  for (auto& s : h->sh) {
    QMFB_CUDA(cudaSetDevice(s.device));
    QMFB_CUDA(cudaMemcpy2DAsync(s.F[side], size_t(h->kp) * 8, host, size_t(h->k) * 8, size_t(h->k) * 8, size_t(h->n[side]),
                                cudaMemcpyHostToDevice, s.stream));
  }
  for (auto& s : h->sh) {
    QMFB_CUDA(cudaSetDevice(s.device));
    QMFB_CUDA(cudaStreamSynchronize(s.stream));
  }
  return QMFB_OK;
}

int qmfb_wals_sharded_get_factors(qmfb_wals_sharded_t* h, int side, int slot, double* host) {
  if (!h || side < 0 || side > 1 || !host || slot < 0 || slot >= h->ndev) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_get_factors: bad argument");
  Shard& s = h->sh[size_t(slot)];
  QMFB_CUDA(cudaSetDevice(s.device));
  QMFB_CUDA(cudaMemcpy2DAsync(host, size_t(h->k) * 8, s.F[side], size_t(h->kp) * 8, size_t(h->k) * 8, size_t(h->n[side]),
                              cudaMemcpyDeviceToHost, s.stream));
  QMFB_CUDA(cudaStreamSynchronize(s.stream));
  return QMFB_OK;
}

int qmfb_wals_sharded_half_step(qmfb_wals_sharded_t* h, int update_side, double alpha, double lambda, double* loss_sum) {
  if (!h || update_side < 0 || update_side > 1) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_half_step: bad argument");
  if (int rc = half_step_async(h, update_side, alpha, lambda, 0)) return rc;
  return finish(h, 1, loss_sum);
}

int qmfb_wals_sharded_epoch_host(qmfb_wals_sharded_t* h, double alpha, double lambda, const double* item_factors_in,
                                 double* user_factors_out, double* item_factors_out, double* loss_out) {
  if (!h) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_epoch_host: null handle");
  const size_t kb = size_t(h->k) * 8, kpb = size_t(h->kp) * 8;
  if (item_factors_in) {  // every device pulls the item factors over its own PCIe link
    for (auto& s : h->sh) {
      QMFB_CUDA(cudaSetDevice(s.device));
      QMFB_CUDA(cudaMemcpy2DAsync(s.F[1], kpb, item_factors_in, kb, kb, size_t(h->n[1]), cudaMemcpyHostToDevice, s.stream));
      QMFB_CUDA(cudaEventRecord(s.ev_solve, s.stream));  // the first half-step waits for every device's upload
    }
  }
  if (int rc = half_step_async(h, QMFB_SIDE_USER, alpha, lambda, 0)) return rc;
  if (user_factors_out) {
    // each device returns the user rows IT solved, on a second stream underneath the item half-step
    for (auto& s : h->sh) {
      if (s.nrows[0] == 0) continue;
      QMFB_CUDA(cudaSetDevice(s.device));
      QMFB_CUDA(cudaStreamWaitEvent(s.copy_stream, s.ev_solve, 0));
      QMFB_CUDA(cudaMemcpy2DAsync(user_factors_out + s.row_begin[0] * h->k, kb, s.F[0] + s.row_begin[0] * h->kp, kpb, kb,
                                  size_t(s.nrows[0]), cudaMemcpyDeviceToHost, s.copy_stream));
      QMFB_CUDA(cudaEventRecord(s.ev_copy, s.copy_stream));
    }
  }
  if (int rc = half_step_async(h, QMFB_SIDE_ITEM, alpha, lambda, 1)) return rc;
  if (item_factors_out) {
    for (auto& s : h->sh) {
      if (s.nrows[1] == 0) continue;
      QMFB_CUDA(cudaSetDevice(s.device));
      QMFB_CUDA(cudaMemcpy2DAsync(item_factors_out + s.row_begin[1] * h->k, kb, s.F[1] + s.row_begin[1] * h->kp, kpb, kb,
                                  size_t(s.nrows[1]), cudaMemcpyDeviceToHost, s.stream));
    }
  }
  double loss = 0.0;
  if (int rc = finish(h, 2, &loss)) return rc;
  // loss / nusers / nitems, WALSEngine.cpp:215
  if (loss_out) *loss_out = loss / double(h->n[0]) / double(h->n[1]);
  return QMFB_OK;
}

int qmfb_wals_sharded_eval_rank(qmfb_wals_sharded_t* h, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                                const int32_t* label_items, int32_t* cnt, double* pos_scores) {
  if (!h) return set_error(QMFB_ERR_INVALID, "qmfb_wals_sharded_eval_rank: null handle");
  std::vector<int> devices;
  std::vector<const double*> U, V;
  for (auto& s : h->sh) {
    QMFB_CUDA(cudaSetDevice(s.device));
    QMFB_CUDA(cudaStreamSynchronize(s.stream));
    devices.push_back(s.device);
    U.push_back(s.F[0]);
    V.push_back(s.F[1]);
  }
  return eval_rank_sharded(h->ndev, devices.data(), U.data(), h->kp, h->n[0], V.data(), h->kp, h->n[1], h->k, test_users, nT, label_ptr,
                           label_items, cnt, pos_scores);
}

double* qmfb_wals_sharded_factors_device(qmfb_wals_sharded_t* h, int side, int slot) {
  return (h && side >= 0 && side <= 1 && slot >= 0 && slot < h->ndev) ? h->sh[size_t(slot)].F[side] : nullptr;
}
int64_t qmfb_wals_sharded_launch_count(qmfb_wals_sharded_t* h) { return h ? h->launches : 0; }
int qmfb_wals_sharded_last_timing(qmfb_wals_sharded_t* h, float* gram_ms, float* solve_ms) {
  if (!h) return set_error(QMFB_ERR_INVALID, "null handle");
  if (gram_ms) *gram_ms = h->gram_ms;
  if (solve_ms) *solve_ms = h->solve_ms;
  return QMFB_OK;
}

}  // extern "C"
