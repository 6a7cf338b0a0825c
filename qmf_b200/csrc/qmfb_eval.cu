// C-ABI (include/qmf_b200.h) for the ranking evaluation: launchers of eval_kernels.cuh, the host-buffer
// entry point and the resident-factor helper the engines use (no host round trip of the factors).
#include "qmfb_common.h"
#include "eval_kernels.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace qmfb {

namespace {

struct DevMem {  // frees on scope exit
  void* p = nullptr;
  ~DevMem() { cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

__global__ void pad_rows_kernel(const double* __restrict__ src, int64_t lds, int k, int64_t n, double* __restrict__ dst, int kp) {
  const int64_t total = n * kp;
  for (int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; q < total; q += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = q / kp;
    const int f = int(q % kp);
    dst[q] = f < k ? src[r * lds + f] : 0.0;
  }
}

}  // namespace

int eval_rank_resident(int device, const double* U, int64_t ldu, int64_t nusers, const double* V, int64_t ldv, int64_t nitems, int k,
                       const double* bias, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                       const int32_t* label_items, int32_t* cnt, double* pos_scores) {
  if (!U || !V || !test_users || !label_ptr || !cnt || nT < 0 || nusers < 1 || nitems < 1) return set_error(QMFB_ERR_INVALID, "eval_rank: bad argument");
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  const int64_t base = label_ptr[0], nl = label_ptr[nT] - base;
  int64_t max_pos = 0;
  for (int64_t t = 0; t < nT; ++t) {
    if (test_users[t] < 0 || test_users[t] >= nusers) return set_error(QMFB_ERR_INVALID, "eval_rank: test user %lld out of range", (long long)t);
    max_pos = std::max(max_pos, label_ptr[t + 1] - label_ptr[t]);
    for (int64_t q = label_ptr[t]; q < label_ptr[t + 1]; ++q) {
      if (label_items[q] < 0 || label_items[q] >= nitems || (q > label_ptr[t] && label_items[q] <= label_items[q - 1])) {
        return set_error(QMFB_ERR_INVALID, "eval_rank: label items of user %lld must be ascending, distinct and in range", (long long)t);
      }
    }
  }
  if (nT == 0) return QMFB_OK;
  QMFB_CUDA(cudaSetDevice(device));
  cudaStream_t st = nullptr;
  QMFB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  DevMem dUp, dVp, dS, dT, dL, dC, dP;
  const int rc = [&]() -> int {
    // the kernels need rows of KP doubles (zero padded), 16-byte aligned: repack if the caller's layout is tighter
    auto padded = [&](const double* src, int64_t ld, int64_t n, DevMem& buf, const double** out, int64_t* ld_out) -> int {
      if (ld >= kp && (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        *out = src;
        *ld_out = ld;
        return QMFB_OK;
      }
      QMFB_CUDA(cudaMalloc(&buf.p, size_t(n) * kp * 8));
      pad_rows_kernel<<<int(std::min<int64_t>(148 * 8, (n * kp + 255) / 256)), 256, 0, st>>>(src, ld, k, n, buf.as<double>(), kp);
      QMFB_CUDA(cudaGetLastError());
      *out = buf.as<double>();
      *ld_out = kp;
      return QMFB_OK;
    };
    const double *Up = nullptr, *Vp = nullptr;
    int64_t ldup = 0, ldvp = 0;
    if (int r = padded(U, ldu, nusers, dUp, &Up, &ldup)) return r;
    if (int r = padded(V, ldv, nitems, dVp, &Vp, &ldvp)) return r;
    QMFB_CUDA(cudaMalloc(&dS.p, size_t(std::max<int64_t>(nl, 1)) * 8));
    QMFB_CUDA(cudaMalloc(&dT.p, size_t(nT) * 4));
    QMFB_CUDA(cudaMalloc(&dL.p, size_t(std::max<int64_t>(nl, 1)) * 4));
    QMFB_CUDA(cudaMalloc(&dC.p, size_t(nl + nT) * 4));
    QMFB_CUDA(cudaMalloc(&dP.p, size_t(nT + 1) * 8));
    std::vector<int64_t> lp(static_cast<size_t>(nT) + 1);
    for (int64_t t = 0; t <= nT; ++t) lp[size_t(t)] = label_ptr[t] - base;  // a slice of a larger test set
    QMFB_CUDA(cudaMemcpyAsync(dT.p, test_users, size_t(nT) * 4, cudaMemcpyHostToDevice, st));
    if (nl > 0) QMFB_CUDA(cudaMemcpyAsync(dL.p, label_items + base, size_t(nl) * 4, cudaMemcpyHostToDevice, st));
    QMFB_CUDA(cudaMemcpyAsync(dP.p, lp.data(), size_t(nT + 1) * 8, cudaMemcpyHostToDevice, st));
    if (int r = qmfb_eval_rank_dev(st, Up, ldup, Vp, ldvp, nitems, k, bias, dT.as<int32_t>(), nT, dP.as<int64_t>(), dL.as<int32_t>(), nl,
                                   max_pos, dC.as<int32_t>(), dS.as<double>())) return r;
    QMFB_CUDA(cudaMemcpyAsync(cnt, dC.p, size_t(nl + nT) * 4, cudaMemcpyDeviceToHost, st));
    if (pos_scores && nl > 0) QMFB_CUDA(cudaMemcpyAsync(pos_scores, dS.p, size_t(nl) * 8, cudaMemcpyDeviceToHost, st));
    QMFB_CUDA(cudaStreamSynchronize(st));
    return QMFB_OK;
  }();
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return rc;
}

int eval_rank_sharded(int ndev, const int* devices, const double* const* U, int64_t ldu, int64_t nusers, const double* const* V, int64_t ldv,
                      int64_t nitems, int k, const int32_t* test_users, int64_t nT, const int64_t* label_ptr, const int32_t* label_items,
                      int32_t* cnt, double* pos_scores) {
  if (ndev < 1 || !label_ptr || nT < 0) return set_error(QMFB_ERR_INVALID, "eval_rank_sharded: bad argument");
  // test users are independent: contiguous slices of the test-user list, one per device, each
  // scored against that device's replicas by its own host thread; the integer counts do not depend on the split
  std::vector<int> rcs(static_cast<size_t>(ndev), QMFB_OK);
  std::vector<std::string> msgs(static_cast<size_t>(ndev));
  std::vector<std::thread> pool;
  for (int d = 0; d < ndev; ++d) {
    const int64_t tb = nT * d / ndev, te = nT * (d + 1) / ndev;
    if (te <= tb) continue;
    pool.emplace_back([=, &rcs, &msgs]() {
      rcs[size_t(d)] = eval_rank_resident(devices[d], U[d], ldu, nusers, V[d], ldv, nitems, k, nullptr, test_users + tb, te - tb,
                                          label_ptr + tb, label_items, cnt + label_ptr[tb] + tb,
                                          pos_scores ? pos_scores + label_ptr[tb] : nullptr);
      if (rcs[size_t(d)] != QMFB_OK) msgs[size_t(d)] = qmfb_last_error();
    });
  }
  for (auto& th : pool) th.join();
  for (int d = 0; d < ndev; ++d) {
    if (rcs[size_t(d)] != QMFB_OK) return set_error(rcs[size_t(d)], "device %d: %s", devices[d], msgs[size_t(d)].c_str());
  }
  return QMFB_OK;
}

}  // namespace qmfb

using namespace qmfb;

extern "C" {

int qmfb_eval_rank_dev(void* stream, const double* U, int64_t ldu, const double* V, int64_t ldv, int64_t nitems, int k,
                       const double* biases, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                       const int32_t* label_items, int64_t nlabels, int64_t max_positives, int32_t* cnt, double* pos_scores) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  if (!U || !V || !test_users || !label_ptr || !cnt || !pos_scores || nT < 0 || nitems < 1 || nitems > INT32_MAX || nT > INT32_MAX ||
      nlabels < 0 || max_positives < 0) {
    return set_error(QMFB_ERR_INVALID, "qmfb_eval_rank_dev: bad argument");
  }
  if (ldu < kp || ldv < kp || (ldu & 1) || (ldv & 1) || (reinterpret_cast<uintptr_t>(U) & 15) || (reinterpret_cast<uintptr_t>(V) & 15)) {
    return set_error(QMFB_ERR_INVALID, "qmfb_eval_rank_dev: factor rows must be 16-byte aligned with an even stride >= %d (zero padded)", kp);
  }
  auto st = static_cast<cudaStream_t>(stream);
  QMFB_CUDA(cudaMemsetAsync(cnt, 0, size_t(nlabels + nT) * sizeof(int32_t), st));
  if (nT == 0) return QMFB_OK;
  // k <= 128: two warp groups per CTA (512 threads) with a 3-stage ring each; above, the 64-user tile alone takes
  // 133 KB and one group with a 3-stage ring is what fits
  const int ng = kp <= 128 ? 2 : 1;
  const int stages = 3;
  const size_t smem = eval_smem_bytes(kp, stages, ng);
  auto kernel = ng == 2 ? eval_score_kernel<2> : eval_score_kernel<1>;
  // per device / context attribute: set on every call (cheap), a process may use several devices
  QMFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int dev = 0, sms = 148, occ = 1;
  QMFB_CUDA(cudaGetDevice(&dev));
  QMFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  QMFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kEvThreads * ng, smem));
  if (occ < 1) return set_error(QMFB_ERR_UNSUPPORTED, "eval_score_kernel does not fit on an SM (nfactors %d)", k);
  double* vnorm = nullptr;  // | nitems norms | unit counter | nlabels positive buckets |
  QMFB_CUDA(cudaMallocAsync(&vnorm, size_t(nitems) * 8 + 16 + size_t(std::max<int64_t>(nlabels, 1)) * 4, st));
  int* counter = reinterpret_cast<int*>(vnorm + nitems);
  int32_t* pos_bucket = counter + 4;
  QMFB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  const int grid = sms * occ;
  const int64_t ngroups = (nT + kEvU - 1) / kEvU;
  int64_t nsplit = (4 * int64_t(grid) + ngroups - 1) / ngroups;
  nsplit = std::max<int64_t>(1, std::min<int64_t>(nsplit, std::max<int64_t>(1, nitems / (4 * kEvI))));
  EvalParams p{U, ldu, V, ldv, biases, k, kp, int(nitems), test_users, int(nT), label_ptr, label_items, cnt, pos_scores, pos_bucket,
               vnorm, int(nsplit), counter, stages, max_positives <= kEvPosSmem ? 1 : 0};
  eval_item_norm_kernel<<<std::min<int64_t>(int64_t(sms) * 8, (nitems + 7) / 8), 256, 0, st>>>(V, ldv, k, int(nitems), vnorm);
  eval_pos_kernel<<<int(std::min<int64_t>(nT, int64_t(sms) * 8)), 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && !p.sort_in_kernel) {
    // a user with more positives than fit one CTA's shared memory: device-wide segmented sort of every
    // user's exact positive scores (CUB: plumbing, off the hot path)
    if (nlabels > INT32_MAX) {
      cudaFreeAsync(vnorm, st);
      return set_error(QMFB_ERR_UNSUPPORTED, "more than 2^31-1 test positives with a user above %d positives", kEvPosSmem);
    }
    double* sorted = nullptr;
    void* tmp = nullptr;
    size_t bytes = 0;
    e = cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, pos_scores, sorted, int(nlabels), int(nT), label_ptr, label_ptr + 1, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&sorted, size_t(nlabels) * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, std::max<size_t>(bytes, 16), st);
    if (e == cudaSuccess) e = cub::DeviceSegmentedSort::SortKeys(tmp, bytes, pos_scores, sorted, int(nlabels), int(nT), label_ptr, label_ptr + 1, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pos_scores, sorted, size_t(nlabels) * 8, cudaMemcpyDeviceToDevice, st);
    if (sorted) cudaFreeAsync(sorted, st);
    if (tmp) cudaFreeAsync(tmp, st);
  }
  if (e == cudaSuccess && nlabels > 0) {
    eval_pos_bucket_kernel<<<int(std::min<int64_t>(int64_t(sms) * 8, (nlabels + 255) / 256)), 256, 0, st>>>(p, nlabels);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    kernel<<<grid, kEvThreads * ng, smem, st>>>(p);
    e = cudaGetLastError();
  }
  cudaFreeAsync(vnorm, st);
  QMFB_CUDA(e);
  return QMFB_OK;
}

int qmfb_eval_rank(int device, const double* U, int64_t nusers, const double* V, int64_t nitems, int k,
                   const double* biases, const int32_t* test_users, int64_t nT, const int64_t* label_ptr,
                   const int32_t* label_items, int32_t* cnt, double* pos_scores) {
  if (!U || !V || nusers < 1 || nitems < 1 || k < 1) return set_error(QMFB_ERR_INVALID, "qmfb_eval_rank: bad argument");
  QMFB_CUDA(cudaSetDevice(device));
  DevMem dU, dV, dB;
  QMFB_CUDA(cudaMalloc(&dU.p, size_t(nusers) * k * 8));
  QMFB_CUDA(cudaMalloc(&dV.p, size_t(nitems) * k * 8));
  QMFB_CUDA(cudaMemcpy(dU.p, U, size_t(nusers) * k * 8, cudaMemcpyHostToDevice));
  QMFB_CUDA(cudaMemcpy(dV.p, V, size_t(nitems) * k * 8, cudaMemcpyHostToDevice));
  if (biases) {
    QMFB_CUDA(cudaMalloc(&dB.p, size_t(nitems) * 8));
    QMFB_CUDA(cudaMemcpy(dB.p, biases, size_t(nitems) * 8, cudaMemcpyHostToDevice));
  }
  return eval_rank_resident(device, dU.as<double>(), k, nusers, dV.as<double>(), k, nitems, k, dB.as<double>(), test_users, nT, label_ptr,
                            label_items, cnt, pos_scores);
}

}  // extern "C"
