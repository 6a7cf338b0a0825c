// Ranking evaluation on the GPU (sm_100a): all-item scores of a test user fused with the rank
// statistics that AUC / AP / P@k / R@k need, never materialising the nT x nitems score matrix.
// Replaces
//   Engine::computeTestScores                     qmf/Engine.cpp:73-96     (K7)
//   AUC / Precision / Recall / AveragePrecision   qmf/metrics/Metrics.cpp:65-164 (K8, the sort)
//
// Exactness: a score is bias_i + sum_f U[u,f] * V[i,f] accumulated in f order with separately
// rounded multiply and add (__dmul_rn / __dadd_rn, no FMA contraction) - bit-identical to the
// reference compiled for x86-64 - so the ranking (every integer below) is bit-exact.
//
// Per test user t with positives P (test items with label > 0) and negatives N (all other
// items, train positives included, Engine.cpp:58-69):
//   sorted positives' scores ascending  s_(0) <= ... <= s_(nP-1)
//   cnt[i], i = 0..nP  = #{ x in N : exactly i positives score strictly less than s_x }
// Under the reference order (score descending, positives first on ties, Metrics.cpp:85-86) a
// negative in bucket i is preceded by exactly nP - i positives, which is all that the metrics
// use; the host turns cnt into AUC / AP / P@k / R@k with the reference's own arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmfb {

constexpr int kEvalThreads = 256;
constexpr int kEvalMaxPos = 2048;   // positives per test user held in shared memory
constexpr int kEvalTileItems = 32;  // items scored per warp pass

struct EvalParams {
  const double* U;          // user factors, row stride ldu
  int64_t ldu;
  const double* V;          // item factors, row stride ldv
  int64_t ldv;
  const double* bias;       // item biases or nullptr
  int k;
  int nitems;
  const int32_t* test_users;   // nT user idx
  const int64_t* label_ptr;    // nT + 1: offsets into label_items and (shifted by t) into cnt
  const int32_t* label_items;  // per user: positive item idx, ascending
  int32_t* cnt;                // out: per user nP + 1 counters at offset label_ptr[t] + t (zeroed by the launcher)
  double* pos_scores;          // out: per user the positives' scores in ascending order (offset label_ptr[t])
  int* error;                  // bit 2: a user has more than kEvalMaxPos positives
};

constexpr int kEvalFC = 32;  // factors per staged chunk

__host__ __device__ inline size_t eval_smem_bytes(int k) {
  const size_t pu = size_t((k + 1) & ~1) * 8;
  const size_t spos = size_t(kEvalMaxPos) * 8;
  const size_t scnt = size_t(kEvalMaxPos + 2) * 4;
  const size_t tiles = size_t(kEvalThreads / 32) * kEvalTileItems * (kEvalFC + 1) * 8;
  return pu + spos + scnt + tiles;
}

__global__ void __launch_bounds__(kEvalThreads) eval_rank_kernel(const EvalParams prm, int nT) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: pu[k] | spos[kEvalMaxPos] | scnt[kEvalMaxPos + 2] (int) | tile[nwarps][32][kEvalFC + 1]
  double* pu = reinterpret_cast<double*>(smem_raw);
  double* spos = pu + ((prm.k + 1) & ~1);
  int* scnt = reinterpret_cast<int*>(spos + kEvalMaxPos);
  double* tiles = reinterpret_cast<double*>(scnt + kEvalMaxPos + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kEvalThreads / 32;
  constexpr int ldt = kEvalFC + 1;  // odd stride: lanes reading one row each hit distinct banks
  double* tile = tiles + size_t(warp) * kEvalTileItems * ldt;

  for (int t = blockIdx.x; t < nT; t += gridDim.x) {
    const int u = prm.test_users[t];
    const int64_t lp0 = prm.label_ptr[t];
    const int nP = int(prm.label_ptr[t + 1] - lp0);
    if (nP > kEvalMaxPos) {
      if (tid == 0) atomicOr(prm.error, 4);
      continue;
    }
    const int32_t* pos_items = prm.label_items + lp0;
    __syncthreads();  // previous user done with shared memory
    for (int f = tid; f < prm.k; f += kEvalThreads) pu[f] = prm.U[int64_t(u) * prm.ldu + f];
    __syncthreads();
    // ---- scores of the positives, then sort ascending (bitonic, padded with +inf) ------------------
    int n2 = 1;
    while (n2 < nP) n2 <<= 1;
    for (int i = tid; i < n2; i += kEvalThreads) {
      double s = __longlong_as_double(0x7ff0000000000000LL);
      if (i < nP) {
        const int item = pos_items[i];
        s = prm.bias != nullptr ? prm.bias[item] : 0.0;
        const double* v = prm.V + int64_t(item) * prm.ldv;
        for (int f = 0; f < prm.k; ++f) s = __dadd_rn(s, __dmul_rn(pu[f], v[f]));  // Engine.cpp:86-91
      }
      spos[i] = s;
    }
    for (int i = tid; i <= nP; i += kEvalThreads) scnt[i] = 0;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < n2; i += kEvalThreads) {
          const int j = i ^ stride;
          if (j > i) {
            const double a = spos[i], b = spos[j];
            const bool up = (i & size) == 0;
            if ((a > b) == up) {
              spos[i] = b;
              spos[j] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    // ---- all items: exact score, then bucket the negatives ------------------------------------------
    for (int x0 = warp * kEvalTileItems; x0 < prm.nitems; x0 += nwarps * kEvalTileItems) {
      const int x = x0 + lane;
      double s = (x < prm.nitems && prm.bias != nullptr) ? prm.bias[x] : 0.0;
      for (int fc = 0; fc < prm.k; fc += kEvalFC) {
        const int nf = min(kEvalFC, prm.k - fc);
        __syncwarp();
#pragma unroll 4
        for (int r = 0; r < kEvalTileItems; ++r) {  // one 8*nf-byte row segment per instruction
          const int xr = x0 + r;
          if (xr < prm.nitems && lane < nf) tile[r * ldt + lane] = prm.V[int64_t(xr) * prm.ldv + fc + lane];
        }
        __syncwarp();
        if (x < prm.nitems) {
          const double* row = tile + lane * ldt;
          for (int f = 0; f < nf; ++f) s = __dadd_rn(s, __dmul_rn(pu[fc + f], row[f]));
        }
      }
      if (x < prm.nitems) {
        int lo = 0, hi = nP;  // is x one of the positives?
        while (lo < hi) {
          const int m = (lo + hi) >> 1;
          if (pos_items[m] < x) lo = m + 1; else hi = m;
        }
        if (!(lo < nP && pos_items[lo] == x)) {
          int a = 0, b = nP;  // number of positives scoring strictly less than s
          while (a < b) {
            const int m = (a + b) >> 1;
            if (spos[m] < s) a = m + 1; else b = m;
          }
          atomicAdd(&scnt[a], 1);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i <= nP; i += kEvalThreads) prm.cnt[lp0 + t + i] = scnt[i];
    for (int i = tid; i < nP; i += kEvalThreads) prm.pos_scores[lp0 + i] = spos[i];
  }
}

}  // namespace qmfb
