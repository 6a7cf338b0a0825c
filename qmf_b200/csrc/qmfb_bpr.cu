// C-ABI (include/qmf_b200.h) for BPR Hogwild SGD (the ranking evaluation is qmfb_eval.cu).
#include "qmfb_common.h"
#include "bpr_kernels.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <numeric>
#include <vector>

namespace qmfb {
__global__ void sum_partials_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  // single thread, fixed order: deterministic
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    out[0] = s;
  }
}
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
}  // namespace qmfb

using namespace qmfb;

struct qmfb_bpr {
  int device = 0;
  int64_t n[2] = {0, 0};
  int k = 0;
  bool use_biases = false;
  cudaStream_t stream = nullptr;
  double* F[2] = {nullptr, nullptr};
  double* bias = nullptr;
  int32_t *du = nullptr, *di = nullptr, *pos_items = nullptr;
  int64_t* pos_ptr = nullptr;
  int64_t npairs = 0;
  int32_t* error = nullptr;
  double *partial = nullptr, *sum = nullptr;
  int32_t* trip = nullptr;
  int64_t trip_cap = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float epoch_ms = 0.f;
  int64_t launches = 0;
  int sms = 148;
  int64_t max_warps = 0;  // 0 = automatic (see bpr_concurrency)
  int64_t hogwild_blocks = 1;  // --num_hogwild_threads: only its tail-drop is mirrored (qmfb_bpr_set_hogwild_blocks)
  // per-epoch shuffle: sort (Philox key, index) pairs
  uint32_t *shuf_key[2] = {nullptr, nullptr};
  int32_t *shuf_val[2] = {nullptr, nullptr};
  void* shuf_tmp = nullptr;
  size_t shuf_tmp_bytes = 0;
};

// Number of pairs processed concurrently (one warp each).  Every in-flight triplet computes its
// step from values that do not yet contain the other in-flight steps, so the steps landing on
// one row behave like ONE step with the learning rate multiplied by their count: with the
// whole GPU in flight on a small catalogue (e.g. 17 000 warps on 300 items) SGD diverges.
// Keeping at most ~min(nusers, nitems) / 2 pairs in flight bounds the expected number of
// concurrent steps per row by ~1 for the uniformly sampled negative (2 item rows per triplet);
// large catalogues use the whole machine.
static int64_t bpr_concurrency(const qmfb_bpr* h) {
  const int64_t hw = int64_t(h->sms) * 64;  // 64 resident warps per SM
  if (h->max_warps > 0) return std::min<int64_t>(h->max_warps, hw);
  const int64_t rows = std::min(h->n[0], h->n[1]);
  return std::max<int64_t>(32, std::min<int64_t>(hw, rows / 2));
}

static constexpr int kLossBlocks = 1184;  // 8 CTAs per SM on 148 SMs

static void bpr_release(qmfb_bpr* h) {
  cudaSetDevice(h->device);
  cudaFree(h->F[0]);
  cudaFree(h->F[1]);
  cudaFree(h->bias);
  cudaFree(h->du);
  cudaFree(h->di);
  cudaFree(h->pos_items);
  cudaFree(h->pos_ptr);
  cudaFree(h->error);
  cudaFree(h->partial);
  cudaFree(h->sum);
  cudaFree(h->trip);
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->shuf_key[b]);
    cudaFree(h->shuf_val[b]);
  }
  cudaFree(h->shuf_tmp);
  for (auto& e : h->ev) {
    if (e) cudaEventDestroy(e);
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" {

int qmfb_bpr_create(int device, int64_t nusers, int64_t nitems, int nfactors, int use_biases, qmfb_bpr_t** out) {
  if (!out || nusers < 1 || nitems < 1 || nitems > INT32_MAX || nusers > INT32_MAX) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_create: bad argument");
  if (nfactors < 1 || nfactors > 32 * kBprMaxPerLane) return set_error(QMFB_ERR_UNSUPPORTED, "nfactors must be in [1, %d] (got %d)", 32 * kBprMaxPerLane, nfactors);
  QMFB_CUDA(cudaSetDevice(device));
  auto* h = new qmfb_bpr;
  h->device = device;
  h->n[0] = nusers;
  h->n[1] = nitems;
  h->k = nfactors;
  h->use_biases = use_biases != 0;
  int rc = [&]() -> int {
    QMFB_CUDA(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device));
    QMFB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int s = 0; s < 2; ++s) {
      QMFB_CUDA(cudaMalloc(&h->F[s], size_t(h->n[s]) * nfactors * sizeof(double)));
      QMFB_CUDA(cudaMemsetAsync(h->F[s], 0, size_t(h->n[s]) * nfactors * sizeof(double), h->stream));
    }
    if (h->use_biases) {
      QMFB_CUDA(cudaMalloc(&h->bias, size_t(nitems) * sizeof(double)));
      QMFB_CUDA(cudaMemsetAsync(h->bias, 0, size_t(nitems) * sizeof(double), h->stream));
    }
    QMFB_CUDA(cudaMalloc(&h->error, sizeof(int32_t)));
    QMFB_CUDA(cudaMalloc(&h->partial, kLossBlocks * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->sum, sizeof(double)));
    for (auto& e : h->ev) QMFB_CUDA(cudaEventCreate(&e));
    QMFB_CUDA(cudaStreamSynchronize(h->stream));
    return QMFB_OK;
  }();
  if (rc != QMFB_OK) {
    bpr_release(h);
    return rc;
  }
  *out = h;
  return QMFB_OK;
}

int qmfb_bpr_destroy(qmfb_bpr_t* h) {
  if (h) bpr_release(h);
  return QMFB_OK;
}

int qmfb_bpr_set_data(qmfb_bpr_t* h, const int32_t* user_idx, const int32_t* item_idx, int64_t npairs) {
  if (!h || npairs < 1 || !user_idx || !item_idx) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_set_data: bad argument");
  for (int64_t p = 0; p < npairs; ++p) {
    if (user_idx[p] < 0 || user_idx[p] >= h->n[0] || item_idx[p] < 0 || item_idx[p] >= h->n[1]) {
      return set_error(QMFB_ERR_INVALID, "qmfb_bpr_set_data: pair %lld out of range", (long long)p);
    }
  }
  // itemMap_: per-user set of positive items (BPREngine.cpp:79-82) as sorted, de-duplicated CSR
  std::vector<int64_t> ptr(size_t(h->n[0]) + 1, 0);
  for (int64_t p = 0; p < npairs; ++p) ++ptr[size_t(user_idx[p]) + 1];
  std::partial_sum(ptr.begin(), ptr.end(), ptr.begin());
  std::vector<int32_t> items(static_cast<size_t>(npairs));
  {
    std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t p = 0; p < npairs; ++p) items[size_t(fill[user_idx[p]]++)] = item_idx[p];
  }
  std::vector<int64_t> optr(size_t(h->n[0]) + 1, 0);
  int64_t w = 0;
  for (int64_t u = 0; u < h->n[0]; ++u) {
    std::sort(items.begin() + ptr[u], items.begin() + ptr[u + 1]);
    optr[u] = w;
    for (int64_t q = ptr[u]; q < ptr[u + 1]; ++q) {
      if (q == ptr[u] || items[q] != items[q - 1]) items[w++] = items[q];
    }
  }
  optr[h->n[0]] = w;
  QMFB_CUDA(cudaSetDevice(h->device));
  cudaFree(h->du); cudaFree(h->di); cudaFree(h->pos_items); cudaFree(h->pos_ptr);
  h->du = h->di = h->pos_items = nullptr;
  h->pos_ptr = nullptr;
  QMFB_CUDA(cudaMalloc(&h->du, size_t(npairs) * sizeof(int32_t)));
  QMFB_CUDA(cudaMalloc(&h->di, size_t(npairs) * sizeof(int32_t)));
  QMFB_CUDA(cudaMalloc(&h->pos_items, size_t(std::max<int64_t>(w, 1)) * sizeof(int32_t)));
  QMFB_CUDA(cudaMalloc(&h->pos_ptr, (size_t(h->n[0]) + 1) * sizeof(int64_t)));
  QMFB_CUDA(cudaMemcpyAsync(h->du, user_idx, size_t(npairs) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaMemcpyAsync(h->di, item_idx, size_t(npairs) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaMemcpyAsync(h->pos_items, items.data(), size_t(w) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaMemcpyAsync(h->pos_ptr, optr.data(), optr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  h->npairs = npairs;
  return QMFB_OK;
}

static int bpr_copy(qmfb_bpr* h, double* dev, double* host_out, const double* host_in, size_t count) {
  QMFB_CUDA(cudaSetDevice(h->device));
  if (host_in) QMFB_CUDA(cudaMemcpyAsync(dev, host_in, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (host_out) QMFB_CUDA(cudaMemcpyAsync(host_out, dev, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return QMFB_OK;
}

int qmfb_bpr_set_factors(qmfb_bpr_t* h, int side, const double* host) {
  if (!h || side < 0 || side > 1 || !host) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_set_factors: bad argument");
  return bpr_copy(h, h->F[side], nullptr, host, size_t(h->n[side]) * h->k);
}
int qmfb_bpr_get_factors(qmfb_bpr_t* h, int side, double* host) {
  if (!h || side < 0 || side > 1 || !host) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_get_factors: bad argument");
  return bpr_copy(h, h->F[side], host, nullptr, size_t(h->n[side]) * h->k);
}
int qmfb_bpr_set_biases(qmfb_bpr_t* h, const double* host) {
  if (!h || !host || !h->use_biases) return set_error(QMFB_ERR_INVALID, "can't access bias when withBiases = false");
  return bpr_copy(h, h->bias, nullptr, host, size_t(h->n[1]));
}
int qmfb_bpr_get_biases(qmfb_bpr_t* h, double* host) {
  if (!h || !host || !h->use_biases) return set_error(QMFB_ERR_INVALID, "can't access bias when withBiases = false");
  return bpr_copy(h, h->bias, host, nullptr, size_t(h->n[1]));
}

static BprParams bpr_params(qmfb_bpr* h, double lr, double ul, double il, double bl) {
  BprParams p{};
  p.P = h->F[0];
  p.Q = h->F[1];
  p.bias = h->bias;
  p.k = h->k;
  p.nitems = int(h->n[1]);
  p.du = h->du;
  p.di = h->di;
  p.npairs = h->npairs;
  p.pos_ptr = h->pos_ptr;
  p.pos_items = h->pos_items;
  p.lr = lr;
  p.user_lambda = ul;
  p.item_lambda = il;
  p.bias_lambda = bl;
  p.error = h->error;
  p.perm = nullptr;
  p.nvisit = h->npairs;
  return p;
}

static int bpr_check_error(qmfb_bpr* h) {
  int32_t err = 0;
  QMFB_CUDA(cudaMemcpyAsync(&err, h->error, sizeof(err), cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  if (err & 1) return set_error(QMFB_ERR_NOT_FINITE, "gradients too big, try decreasing the learning rate (--init_learning_rate)");
  if (err & 2) return set_error(QMFB_ERR_INVALID, "negative sampling gave up: a user is positive on (nearly) every item");
  return QMFB_OK;
}

int qmfb_bpr_epoch(qmfb_bpr_t* h, double lr, double user_lambda, double item_lambda, double bias_lambda, int num_neg,
                   uint64_t seed, uint64_t epoch, int shuffle, int64_t* n_updates) {
  if (!h || num_neg < 0) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_epoch: bad argument");
  if (!h->du) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_epoch: no training data set");
  QMFB_CUDA(cudaSetDevice(h->device));
  BprParams p = bpr_params(h, lr, user_lambda, item_lambda, bias_lambda);
  p.num_neg = num_neg;
  const uint64_t key = splitmix64(seed ^ splitmix64(epoch));
  p.seed_lo = uint32_t(key);
  p.seed_hi = uint32_t(key >> 32);
  if (uint64_t(h->npairs) >= (1ull << 31)) return set_error(QMFB_ERR_UNSUPPORTED, "more than 2^31 training pairs");
  if (shuffle && h->npairs > 1) {
    // a fresh uniformly random visiting order (what std::shuffle(data_) gives the reference's next epoch)
    const int64_t n = h->npairs;
    if (!h->shuf_key[0]) {
      for (int b = 0; b < 2; ++b) {
        QMFB_CUDA(cudaMalloc(&h->shuf_key[b], size_t(n) * 4));
        QMFB_CUDA(cudaMalloc(&h->shuf_val[b], size_t(n) * 4));
      }
      QMFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, h->shuf_tmp_bytes, h->shuf_key[0], h->shuf_key[1], h->shuf_val[0], h->shuf_val[1], n));
      QMFB_CUDA(cudaMalloc(&h->shuf_tmp, std::max<size_t>(h->shuf_tmp_bytes, 16)));
    }
    bpr_shuffle_keys_kernel<<<int(std::min<int64_t>(int64_t(h->sms) * 8, (n + 255) / 256)), 256, 0, h->stream>>>(p.seed_lo, p.seed_hi, n, h->shuf_key[0],
                                                                                                                 h->shuf_val[0]);
    QMFB_CUDA(cudaGetLastError());
    QMFB_CUDA(cub::DeviceRadixSort::SortPairs(h->shuf_tmp, h->shuf_tmp_bytes, h->shuf_key[0], h->shuf_key[1], h->shuf_val[0], h->shuf_val[1], n, 0, 32,
                                              h->stream));
    p.perm = h->shuf_val[1];
    h->launches += 1;
  }
  // Hogwild dispatch of the reference: numTasks blocks of floor(ndata / numTasks) pairs, the tail is skipped
  p.nvisit = (h->npairs / h->hogwild_blocks) * h->hogwild_blocks;
  QMFB_CUDA(cudaMemsetAsync(h->error, 0, sizeof(int32_t), h->stream));
  const int64_t warps = std::max<int64_t>(1, std::min<int64_t>(p.nvisit, bpr_concurrency(h)));
  const int blocks = int((warps + 7) / 8);
  QMFB_CUDA(cudaEventRecord(h->ev[0], h->stream));
  switch ((h->k + 31) / 32) {
    case 1: bpr_epoch_kernel<1><<<blocks, 256, 0, h->stream>>>(p); break;
    case 2: bpr_epoch_kernel<2><<<blocks, 256, 0, h->stream>>>(p); break;
    case 3: bpr_epoch_kernel<3><<<blocks, 256, 0, h->stream>>>(p); break;
    case 4: bpr_epoch_kernel<4><<<blocks, 256, 0, h->stream>>>(p); break;
    case 5: bpr_epoch_kernel<5><<<blocks, 256, 0, h->stream>>>(p); break;
    case 6: bpr_epoch_kernel<6><<<blocks, 256, 0, h->stream>>>(p); break;
    case 7: bpr_epoch_kernel<7><<<blocks, 256, 0, h->stream>>>(p); break;
    default: bpr_epoch_kernel<8><<<blocks, 256, 0, h->stream>>>(p); break;
  }
  QMFB_CUDA(cudaGetLastError());
  QMFB_CUDA(cudaEventRecord(h->ev[1], h->stream));
  h->launches += 1;
  int rc = bpr_check_error(h);
  QMFB_CUDA(cudaEventElapsedTime(&h->epoch_ms, h->ev[0], h->ev[1]));
  if (n_updates) *n_updates = p.nvisit * num_neg;
  return rc;
}

static int bpr_upload_triplets(qmfb_bpr* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t n) {
  for (int64_t t = 0; t < n; ++t) {
    if (u[t] < 0 || u[t] >= h->n[0] || i[t] < 0 || i[t] >= h->n[1] || j[t] < 0 || j[t] >= h->n[1]) {
      return set_error(QMFB_ERR_INVALID, "triplet %lld out of range", (long long)t);
    }
  }
  if (n > h->trip_cap) {
    cudaFree(h->trip);
    h->trip = nullptr;
    h->trip_cap = 0;
    QMFB_CUDA(cudaMalloc(&h->trip, size_t(3 * n) * sizeof(int32_t)));
    h->trip_cap = n;
  }
  QMFB_CUDA(cudaMemcpyAsync(h->trip, u, size_t(n) * 4, cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaMemcpyAsync(h->trip + n, i, size_t(n) * 4, cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaMemcpyAsync(h->trip + 2 * n, j, size_t(n) * 4, cudaMemcpyHostToDevice, h->stream));
  return QMFB_OK;
}

int qmfb_bpr_update_triplets(qmfb_bpr_t* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t n, double lr,
                             double user_lambda, double item_lambda, double bias_lambda) {
  if (!h || n < 0 || (n > 0 && (!u || !i || !j))) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_update_triplets: bad argument");
  if (n == 0) return QMFB_OK;
  QMFB_CUDA(cudaSetDevice(h->device));
  int rc = bpr_upload_triplets(h, u, i, j, n);
  if (rc) return rc;
  BprParams p = bpr_params(h, lr, user_lambda, item_lambda, bias_lambda);
  QMFB_CUDA(cudaMemsetAsync(h->error, 0, sizeof(int32_t), h->stream));
  switch ((h->k + 31) / 32) {
    case 1: bpr_replay_kernel<1><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    case 2: bpr_replay_kernel<2><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    case 3: bpr_replay_kernel<3><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    case 4: bpr_replay_kernel<4><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    case 5: bpr_replay_kernel<5><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    case 6: bpr_replay_kernel<6><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    case 7: bpr_replay_kernel<7><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
    default: bpr_replay_kernel<8><<<1, 32, 0, h->stream>>>(p, h->trip, h->trip + n, h->trip + 2 * n, n); break;
  }
  QMFB_CUDA(cudaGetLastError());
  h->launches += 1;
  return bpr_check_error(h);
}

int qmfb_bpr_eval_loss(qmfb_bpr_t* h, const int32_t* u, const int32_t* i, const int32_t* j, int64_t n, double* sum_out) {
  if (!h || n < 0 || !sum_out || (n > 0 && (!u || !i || !j))) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_eval_loss: bad argument");
  if (n == 0) {
    *sum_out = 0.0;
    return QMFB_OK;
  }
  QMFB_CUDA(cudaSetDevice(h->device));
  int rc = bpr_upload_triplets(h, u, i, j, n);
  if (rc) return rc;
  const int blocks = int(std::min<int64_t>((n + 7) / 8, kLossBlocks));
  switch ((h->k + 31) / 32) {
    case 1: bpr_eval_loss_kernel<1><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    case 2: bpr_eval_loss_kernel<2><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    case 3: bpr_eval_loss_kernel<3><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    case 4: bpr_eval_loss_kernel<4><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    case 5: bpr_eval_loss_kernel<5><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    case 6: bpr_eval_loss_kernel<6><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    case 7: bpr_eval_loss_kernel<7><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
    default: bpr_eval_loss_kernel<8><<<blocks, 256, 0, h->stream>>>(h->F[0], h->F[1], h->bias, h->k, h->trip, h->trip + n, h->trip + 2 * n, n, h->partial); break;
  }
  QMFB_CUDA(cudaGetLastError());
  sum_partials_kernel<<<1, 32, 0, h->stream>>>(h->partial, blocks, h->sum);
  QMFB_CUDA(cudaGetLastError());
  h->launches += 2;
  QMFB_CUDA(cudaMemcpyAsync(sum_out, h->sum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return QMFB_OK;
}

int qmfb_bpr_set_concurrency(qmfb_bpr_t* h, int64_t max_pairs_in_flight) {
  if (!h || max_pairs_in_flight < 0) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_set_concurrency: bad argument");
  h->max_warps = max_pairs_in_flight;
  return QMFB_OK;
}

int qmfb_bpr_set_hogwild_blocks(qmfb_bpr_t* h, int64_t num_hogwild_threads) {
  if (!h || num_hogwild_threads < 0) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_set_hogwild_blocks: bad argument");
  h->hogwild_blocks = std::max<int64_t>(1, num_hogwild_threads);
  return QMFB_OK;
}

int qmfb_bpr_last_epoch_ms(qmfb_bpr_t* h, float* ms) {
  if (!h || !ms) return set_error(QMFB_ERR_INVALID, "null argument");
  *ms = h->epoch_ms;
  return QMFB_OK;
}
int qmfb_bpr_eval_rank(qmfb_bpr_t* h, const int32_t* test_users, int64_t nT, const int64_t* label_ptr, const int32_t* label_items,
                       int32_t* cnt, double* pos_scores) {
  if (!h) return set_error(QMFB_ERR_INVALID, "qmfb_bpr_eval_rank: null handle");
  QMFB_CUDA(cudaSetDevice(h->device));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return eval_rank_resident(h->device, h->F[0], h->k, h->n[0], h->F[1], h->k, h->n[1], h->k, h->use_biases ? h->bias : nullptr,
                            test_users, nT, label_ptr, label_items, cnt, pos_scores);
}

double* qmfb_bpr_factors_device(qmfb_bpr_t* h, int side) { return (h && side >= 0 && side <= 1) ? h->F[side] : nullptr; }
double* qmfb_bpr_biases_device(qmfb_bpr_t* h) { return h ? h->bias : nullptr; }
int64_t qmfb_bpr_launch_count(qmfb_bpr_t* h) { return h ? h->launches : 0; }

}  // extern "C"
