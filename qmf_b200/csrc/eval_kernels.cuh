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

constexpr int kEvalThreads = 512;
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

constexpr int kEvalFC = 16;     // factors per staged chunk
constexpr int kEvalGroup = 4;   // test users scored together by a CTA: every staged item row is used kEvalGroup
                                // times (less L2 traffic) and every lane runs kEvalGroup independent
                                // multiply-add chains (the exact-order sum of one score is one dependent chain)

__host__ __device__ inline size_t eval_smem_bytes(int k) {
  const size_t pu = size_t(kEvalGroup) * size_t((k + 1) & ~1) * 8;
  const size_t spos = size_t(kEvalGroup) * kEvalMaxPos * 8;
  const size_t scnt = size_t(kEvalGroup) * (kEvalMaxPos + 2) * 4;
  const size_t tiles = size_t(kEvalThreads / 32) * kEvalTileItems * (kEvalFC + 1) * 8;
  return pu + spos + scnt + tiles;
}

__global__ void __launch_bounds__(kEvalThreads) eval_rank_kernel(const EvalParams prm, int nT) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: pu[G][kpad] | spos[G][kEvalMaxPos] | scnt[G][kEvalMaxPos + 2] (int) | tile[nwarps][32][kEvalFC + 1]
  constexpr int G = kEvalGroup;
  const int kpad = (prm.k + 1) & ~1;
  double* pu = reinterpret_cast<double*>(smem_raw);
  double* spos = pu + size_t(G) * kpad;
  int* scnt = reinterpret_cast<int*>(spos + size_t(G) * kEvalMaxPos);
  double* tiles = reinterpret_cast<double*>(scnt + size_t(G) * (kEvalMaxPos + 2));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kEvalThreads / 32;
  constexpr int ldt = kEvalFC + 1;  // odd stride: lanes reading one row each hit distinct banks
  double* tile = tiles + size_t(warp) * kEvalTileItems * ldt;
  const int ngroups = (nT + G - 1) / G;

  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int t0 = grp * G;
    int nPg[G];
    const int32_t* pos_items[G];
    int64_t lp0g[G];
    bool live[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int t = t0 + g;
      live[g] = t < nT;
      lp0g[g] = live[g] ? prm.label_ptr[t] : 0;
      nPg[g] = live[g] ? int(prm.label_ptr[t + 1] - lp0g[g]) : 0;
      pos_items[g] = prm.label_items + lp0g[g];
      if (nPg[g] > kEvalMaxPos) {  // reported, user skipped (its counters stay zero)
        if (tid == 0) atomicOr(prm.error, 4);
        live[g] = false;
        nPg[g] = 0;
      }
    }
    __syncthreads();  // previous group done with shared memory
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int u = live[g] ? prm.test_users[t0 + g] : 0;
      for (int f = tid; f < prm.k; f += kEvalThreads) pu[g * kpad + f] = live[g] ? prm.U[int64_t(u) * prm.ldu + f] : 0.0;
    }
    __syncthreads();
    // ---- scores of the positives, then sort ascending (bitonic, padded with +inf), user by user ----
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int nP = nPg[g];
      double* sp = spos + size_t(g) * kEvalMaxPos;
      int* sc = scnt + size_t(g) * (kEvalMaxPos + 2);
      int n2 = 1;
      while (n2 < nP) n2 <<= 1;
      for (int i = tid; i < n2; i += kEvalThreads) {
        double s = __longlong_as_double(0x7ff0000000000000LL);
        if (i < nP) {
          const int item = pos_items[g][i];
          s = prm.bias != nullptr ? prm.bias[item] : 0.0;
          const double* v = prm.V + int64_t(item) * prm.ldv;
          for (int f = 0; f < prm.k; ++f) s = __dadd_rn(s, __dmul_rn(pu[g * kpad + f], v[f]));  // Engine.cpp:86-91
        }
        sp[i] = s;
      }
      for (int i = tid; i <= nP; i += kEvalThreads) sc[i] = 0;
      __syncthreads();
      for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int i = tid; i < n2; i += kEvalThreads) {
            const int j = i ^ stride;
            if (j > i) {
              const double a = sp[i], b = sp[j];
              const bool up = (i & size) == 0;
              if ((a > b) == up) {
                sp[i] = b;
                sp[j] = a;
              }
            }
          }
          __syncthreads();
        }
      }
    }
    // ---- all items: exact scores for the G users, then bucket the negatives ---------------------------
    // the staged tile is software-pipelined: the global loads of the NEXT (item pass, factor chunk)
    // are in flight in registers while the current chunk is multiplied out of shared memory
    constexpr int kPre = kEvalTileItems / 2;  // two item rows per load instruction (a half-warp each)
    const int nchunks = (prm.k + kEvalFC - 1) / kEvalFC;
    double pre[kPre];
    auto fetch = [&](int x0, int fc) {
      const int nf = min(kEvalFC, prm.k - fc);
#pragma unroll
      for (int r = 0; r < kPre; ++r) {
        const int xr = x0 + 2 * r + (lane >> 4), ff = lane & 15;
        pre[r] = (xr < prm.nitems && ff < nf) ? prm.V[int64_t(xr) * prm.ldv + fc + ff] : 0.0;
      }
    };
    if (warp * kEvalTileItems < prm.nitems) fetch(warp * kEvalTileItems, 0);
    for (int x0 = warp * kEvalTileItems; x0 < prm.nitems; x0 += nwarps * kEvalTileItems) {
      const int x = x0 + lane;
      const double b0 = (x < prm.nitems && prm.bias != nullptr) ? prm.bias[x] : 0.0;
      double s[G];
#pragma unroll
      for (int g = 0; g < G; ++g) s[g] = b0;
      for (int c = 0; c < nchunks; ++c) {
        const int fc = c * kEvalFC, nf = min(kEvalFC, prm.k - fc);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < kPre; ++r) tile[(2 * r + (lane >> 4)) * ldt + (lane & 15)] = pre[r];
        __syncwarp();
        if (c + 1 < nchunks) {
          fetch(x0, fc + kEvalFC);
        } else if (x0 + nwarps * kEvalTileItems < prm.nitems) {
          fetch(x0 + nwarps * kEvalTileItems, 0);
        }
        if (x < prm.nitems) {
          const double* row = tile + lane * ldt;
          const double* pf = pu + fc;
          for (int f = 0; f < nf; ++f) {
            const double v = row[f];
#pragma unroll
            for (int g = 0; g < G; ++g) s[g] = __dadd_rn(s[g], __dmul_rn(pf[g * kpad + f], v));
          }
        }
      }
      // Bucket EVERY item as if it were a negative (the positives are taken out again after the loop:
      // no membership test in the hot loop); lanes with the same bucket share one shared-memory atomic.
      const unsigned act = __ballot_sync(0xffffffffu, x < prm.nitems);
      if (x < prm.nitems) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (!live[g]) continue;
          const double* sp = spos + size_t(g) * kEvalMaxPos;
          int a = 0, b = nPg[g];  // number of positives scoring strictly less than s
          while (a < b) {
            const int m = (a + b) >> 1;
            if (sp[m] < s[g]) a = m + 1; else b = m;
          }
          const unsigned same = __match_any_sync(act, a);
          if (lane == __ffs(same) - 1) atomicAdd(&scnt[size_t(g) * (kEvalMaxPos + 2) + a], __popc(same));
        }
      }
    }
    __syncthreads();
    // a positive item's own score is bit-identical to its entry in spos, so the bucket it was counted
    // in above is lower_bound(spos, that score): take the nP positives out again
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const double* sp = spos + size_t(g) * kEvalMaxPos;
      for (int i = tid; i < nPg[g]; i += kEvalThreads) {
        const double v = sp[i];
        int a = 0, b = nPg[g];
        while (a < b) {
          const int m = (a + b) >> 1;
          if (sp[m] < v) a = m + 1; else b = m;
        }
        atomicSub(&scnt[size_t(g) * (kEvalMaxPos + 2) + a], 1);
      }
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (!live[g]) continue;
      const int t = t0 + g;
      for (int i = tid; i <= nPg[g]; i += kEvalThreads) prm.cnt[lp0g[g] + t + i] = scnt[size_t(g) * (kEvalMaxPos + 2) + i];
      for (int i = tid; i < nPg[g]; i += kEvalThreads) prm.pos_scores[lp0g[g] + i] = spos[size_t(g) * kEvalMaxPos + i];
    }
  }
}

}  // namespace qmfb
