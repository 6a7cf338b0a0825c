// BPR Hogwild SGD on the GPU (sm_100a): lock-free, one warp per positive pair, Philox negative
// sampling on the device.  Replaces the reference's CPU loops
//   BPREngine::update / predictDifference / lossDerivative   qmf/bpr/BPREngine.cpp:178-244  (K5)
//   BPREngine::iterate / iterateBlock / sampleRandomNegative  qmf/bpr/BPREngine-inl.h:19-60
//   loss half of BPREngine::evaluate                          qmf/bpr/BPREngine.cpp:246-261  (K6)
//
// This path is HBM/L2-bandwidth bound (three gathered rows read and written per triplet, ~16k
// flops): no tensor cores, the work is coalescing and keeping enough warps in flight.  A warp
// owns one (user, positive item) pair for all of its num_neg negatives: lane l holds factors
// l, l+32, ... of p_u and q_i in registers across the negatives (read once, written once per
// pair) and streams q_j; row accesses are contiguous 8 k-byte segments.  No locks: every warp
// computes its step from whatever values it reads (stale under concurrency, as in Hogwild) and
// applies the step as fire-and-forget FP64 reductions (red.global.add.f64) so that the steps of
// the ~10^4 concurrent warps add up instead of overwriting each other - with plain load/store
// as on the CPU, two orders of magnitude more concurrent writers than the reference's <= 64
// threads lose most updates on small item sets (measured: no learning on a 600 x 400 problem).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmfb {

constexpr int kBprPrefetch = 3;    // negatives sampled / fetched together per positive pair
constexpr int kBprMaxPerLane = 8;  // factors per lane: supports nfactors <= 256

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based: no RNG state in memory ------------------
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) const {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

struct BprParams {
  double* P;              // user factors, nusers x k row-major
  double* Q;              // item factors, nitems x k
  double* bias;           // item biases or nullptr (BPRConfig::useBiases)
  int k;
  int nitems;
  const int32_t* du;      // positive pairs in data_ order (BPREngine.cpp:76)
  const int32_t* di;
  int64_t npairs;
  const int64_t* pos_ptr; // per-user sorted positive item sets (itemMap_, BPREngine.cpp:79-82)
  const int32_t* pos_items;
  double lr, user_lambda, item_lambda, bias_lambda;
  int num_neg;
  uint32_t seed_lo, seed_hi;  // Philox key: (seed, epoch)
  const int32_t* perm;        // this epoch's visiting order of the pairs (a uniformly random permutation, the
                              // distribution std::shuffle(data_) gives, BPREngine.cpp:276-278) or nullptr = file order
  int64_t nvisit;             // positions visited: npairs minus the tail the Hogwild block split drops (:156-160)
  int* error;             // bit 0: non-finite gradient (CHECK(std::isfinite(e)), BPREngine.cpp:184-185)
                          // bit 1: a user is positive on (almost) every item, negative sampling gave up
};

__device__ __forceinline__ bool bpr_contains(const int32_t* items, int64_t lo, int64_t hi, int32_t x) {
  const int64_t end = hi;
  while (lo < hi) {
    const int64_t m = (lo + hi) >> 1;
    if (__ldg(items + m) < x) {
      lo = m + 1;
    } else {
      hi = m;
    }
  }
  return lo < end && __ldg(items + lo) == x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One SGD step on the triplet (u, i, j) with p_u and q_i held in registers (BPREngine::update).
// Returns e.  q_j is read from and written back to global memory.
template <int NPL>
__device__ __forceinline__ void bpr_load_row(const BprParams& prm, int32_t j, double (&qj)[NPL], double& bj, int lane) {
  const double* qjp = prm.Q + int64_t(j) * prm.k;
#pragma unroll
  for (int m = 0; m < NPL; ++m) {
    const int f = lane + 32 * m;
    qj[m] = f < prm.k ? qjp[f] : 0.0;
  }
  bj = prm.bias != nullptr ? prm.bias[j] : 0.0;
}

template <int NPL, bool ATOMIC>
__device__ __forceinline__ double bpr_step(const BprParams& prm, double (&pu)[NPL], double (&qi)[NPL], double& bi,
                                           double (&dpu)[NPL], double (&dqi)[NPL], double& dbi, int32_t j,
                                           const double (&qj)[NPL], double bj, int lane) {
  double* qjp = prm.Q + int64_t(j) * prm.k;
  // x = b_i - b_j + p_u . (q_i - q_j), BPREngine.cpp:222-235
  double part = 0.0;
#pragma unroll
  for (int m = 0; m < NPL; ++m) part += pu[m] * (qi[m] - qj[m]);
  double x = warp_sum(part);
  if (prm.bias != nullptr) x += bi - bj;
  const double e = 1.0 / (1.0 + exp(x));  // lossDerivative, BPREngine.cpp:241-244
  if (prm.bias != nullptr) {  // BPREngine.cpp:189-196
    const double si = prm.lr * (e - prm.bias_lambda * bi);
    const double sj = prm.lr * (-e - prm.bias_lambda * bj);
    bi += si;
    dbi += si;
    if (lane == 0) {
      if (ATOMIC) atomicAdd(prm.bias + j, sj); else prm.bias[j] = bj + sj;
    }
  }
#pragma unroll
  for (int m = 0; m < NPL; ++m) {
    // p_u with the OLD q_i, q_j (:200-205); q_i, q_j with the NEW p_u (:208-219)
    const double su = prm.lr * (e * (qi[m] - qj[m]) - prm.user_lambda * pu[m]);
    pu[m] += su;
    dpu[m] += su;
    const double sq = prm.lr * (e * pu[m] - prm.item_lambda * qi[m]);
    qi[m] += sq;
    dqi[m] += sq;
    const double sn = prm.lr * (-e * pu[m] - prm.item_lambda * qj[m]);
    const int f = lane + 32 * m;
    if (f < prm.k) {
      if (ATOMIC) atomicAdd(qjp + f, sn); else qjp[f] = qj[m] + sn;
    }
  }
  return e;
}

template <int NPL>
__global__ void __launch_bounds__(256) bpr_epoch_kernel(const BprParams prm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  Philox rng{prm.seed_lo, prm.seed_hi};
  bool bad = false, gave_up = false;
  for (int64_t w = warp0; w < prm.nvisit; w += nwarps) {
    const int64_t p = prm.perm != nullptr ? int64_t(__ldg(prm.perm + w)) : w;
    const int32_t u = __ldg(prm.du + p), i = __ldg(prm.di + p);
    const int64_t lo = __ldg(prm.pos_ptr + u), hi = __ldg(prm.pos_ptr + u + 1);
    double pu[NPL], qi[NPL];
    double* pup = prm.P + int64_t(u) * prm.k;
    double* qip = prm.Q + int64_t(i) * prm.k;
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
      const int f = lane + 32 * m;
      pu[m] = f < prm.k ? pup[f] : 0.0;
      qi[m] = f < prm.k ? qip[f] : 0.0;
    }
    double bi = prm.bias != nullptr ? prm.bias[i] : 0.0;
    double dpu[NPL], dqi[NPL], dbi = 0.0;
#pragma unroll
    for (int m = 0; m < NPL; ++m) dpu[m] = dqi[m] = 0.0;
    // sampleRandomNegative (BPREngine-inl.h:48-60): uniform item, reject the user's train positives.
    // The candidates of up to kBprPrefetch negatives are drawn and checked TOGETHER (the binary
    // searches over the user's positives advance in lockstep: independent loads in flight instead of
    // one dependent chain per negative) and their rows are fetched together before the sequential
    // SGD steps - the kernel is latency bound (profiles/r01_bpr_large_ncu.csv: 33 % of the warps
    // resident, 6.8 warps stalled on memory per issue), not bandwidth bound.
    auto draw = [&](int n, int first) -> int32_t {  // candidates first, first+1, ... of negative n
      for (int c = first; c < 64; ++c) {
        uint32_t r[4];
        rng(uint32_t(p), uint32_t(uint64_t(p) >> 32), uint32_t(n), uint32_t(c >> 2), r);
        const int32_t cand = int32_t(__umulhi(r[c & 3], uint32_t(prm.nitems)));
        if (!bpr_contains(prm.pos_items, lo, hi, cand)) return cand;
      }
      return -1;
    };
    for (int n0 = 0; n0 < prm.num_neg; n0 += kBprPrefetch) {
      int32_t js[kBprPrefetch];
      int64_t l[kBprPrefetch];
#pragma unroll
      for (int a = 0; a < kBprPrefetch; ++a) {
        uint32_t r[4];
        rng(uint32_t(p), uint32_t(uint64_t(p) >> 32), uint32_t(n0 + a), 0u, r);
        js[a] = int32_t(__umulhi(r[0], uint32_t(prm.nitems)));
        l[a] = lo;
      }
      // Membership of the candidates in the user's sorted positives.  Up to 32 positives (the common case): every lane
      // holds one of them - ONE coalesced load instead of log2(n) dependent ones, then a vote per candidate (the
      // rejection search was 18 % of the warp samples, all of it load latency: profiles/r02_bpr_large_lines.txt).
      const bool few = hi - lo <= 32;
      if (few) {
        const int32_t mine = lane < hi - lo ? __ldg(prm.pos_items + lo + lane) : -1;
#pragma unroll
        for (int a = 0; a < kBprPrefetch; ++a) {
          // leave l[a] at a position whose item equals the candidate iff the candidate is a positive
          const unsigned hit = __ballot_sync(0xffffffffu, mine == js[a]);
          l[a] = hit != 0u ? lo + (__ffs(hit) - 1) : hi;
        }
      } else {  // lockstep lower_bound of every candidate in pos_items[lo, hi)
        int64_t h[kBprPrefetch];
#pragma unroll
        for (int a = 0; a < kBprPrefetch; ++a) h[a] = hi;
        for (int64_t span = hi - lo; span > 0; span >>= 1) {
#pragma unroll
          for (int a = 0; a < kBprPrefetch; ++a) {
            if (l[a] < h[a]) {
              const int64_t mid = (l[a] + h[a]) >> 1;
              if (__ldg(prm.pos_items + mid) < js[a]) l[a] = mid + 1; else h[a] = mid;
            }
          }
        }
      }
#pragma unroll
      for (int a = 0; a < kBprPrefetch; ++a) {
        if (n0 + a < prm.num_neg) {
          if (l[a] < hi && __ldg(prm.pos_items + l[a]) == js[a]) js[a] = draw(n0 + a, 1);  // rare: a train positive
          if (js[a] < 0) gave_up = true;
        } else {
          js[a] = -1;
        }
      }
      double qj[kBprPrefetch][NPL], bj[kBprPrefetch];
#pragma unroll
      for (int a = 0; a < kBprPrefetch; ++a) {
        if (js[a] >= 0) bpr_load_row<NPL>(prm, js[a], qj[a], bj[a], lane);
      }
#pragma unroll
      for (int a = 0; a < kBprPrefetch; ++a) {
        if (js[a] >= 0) {
          const double e = bpr_step<NPL, true>(prm, pu, qi, bi, dpu, dqi, dbi, js[a], qj[a], bj[a], lane);
          bad = bad || !isfinite(e);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
      const int f = lane + 32 * m;
      if (f < prm.k) {
        atomicAdd(pup + f, dpu[m]);
        atomicAdd(qip + f, dqi[m]);
      }
    }
    if (prm.bias != nullptr && lane == 0) atomicAdd(prm.bias + i, dbi);
  }
  if (lane == 0 && (bad || gave_up)) atomicOr(prm.error, (bad ? 1 : 0) | (gave_up ? 2 : 0));
}

// sort keys of the per-epoch shuffle: key[p] = Philox(p; seed, epoch), val[p] = p.  Sorting the pairs by key
// gives a uniformly random permutation (ties of the 32-bit keys keep index order: negligible bias).
__global__ void bpr_shuffle_keys_kernel(uint32_t seed_lo, uint32_t seed_hi, int64_t n, uint32_t* __restrict__ key,
                                        int32_t* __restrict__ val) {
  Philox rng{seed_lo ^ 0x5bd1e995u, seed_hi ^ 0x27d4eb2fu};
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    uint32_t r[4];
    rng(uint32_t(p), uint32_t(uint64_t(p) >> 32), 0x9e3779b9u, 0u, r);
    key[p] = r[0];
    val[p] = int32_t(p);
  }
}

// Deterministic replay: ONE warp applies the given triplets in order, writing every row back
// after every step exactly like BPREngine::update.  Used for single-step parity with the oracle
// and as the sequential (num_hogwild_threads <= 1) semantics on explicit triplets.
template <int NPL>
__global__ void bpr_replay_kernel(BprParams prm, const int32_t* __restrict__ tu, const int32_t* __restrict__ ti,
                                  const int32_t* __restrict__ tj, int64_t n) {
  const int lane = threadIdx.x & 31;
  bool bad = false;
  for (int64_t t = 0; t < n; ++t) {
    const int32_t u = tu[t], i = ti[t], j = tj[t];
    double pu[NPL], qi[NPL];
    double* pup = prm.P + int64_t(u) * prm.k;
    double* qip = prm.Q + int64_t(i) * prm.k;
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
      const int f = lane + 32 * m;
      pu[m] = f < prm.k ? pup[f] : 0.0;
      qi[m] = f < prm.k ? qip[f] : 0.0;
    }
    double bi = prm.bias != nullptr ? prm.bias[i] : 0.0;
    double dpu[NPL], dqi[NPL], dbi = 0.0;
#pragma unroll
    for (int m = 0; m < NPL; ++m) dpu[m] = dqi[m] = 0.0;
    double qj[NPL], bj;
    bpr_load_row<NPL>(prm, j, qj, bj, lane);
    const double e = bpr_step<NPL, false>(prm, pu, qi, bi, dpu, dqi, dbi, j, qj, bj, lane);
    bad = bad || !isfinite(e);
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
      const int f = lane + 32 * m;
      if (f < prm.k) {
        pup[f] = pu[m];
        qip[f] = qi[m];
      }
    }
    if (prm.bias != nullptr && lane == 0) prm.bias[i] = bi;
    __syncwarp();
    __threadfence_block();
  }
  if (lane == 0 && bad) atomicOr(prm.error, 1);
}

// sum over triplets of log(1 + exp(-x)) (BPREngine::loss, BPREngine.cpp:237-239): one warp per
// triplet, per-block partial sums, reduced in fixed order by sum_kernel (deterministic)
template <int NPL>
__global__ void __launch_bounds__(256) bpr_eval_loss_kernel(const double* __restrict__ P, const double* __restrict__ Q,
                                                            const double* __restrict__ bias, int k,
                                                            const int32_t* __restrict__ tu, const int32_t* __restrict__ ti,
                                                            const int32_t* __restrict__ tj, int64_t n,
                                                            double* __restrict__ partial) {
  __shared__ double sh[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  double acc = 0.0;
  for (int64_t t = warp0; t < n; t += nwarps) {
    const int32_t u = tu[t], i = ti[t], j = tj[t];
    double part = 0.0;
#pragma unroll
    for (int m = 0; m < NPL; ++m) {
      const int f = lane + 32 * m;
      if (f < k) part += P[int64_t(u) * k + f] * (Q[int64_t(i) * k + f] - Q[int64_t(j) * k + f]);
    }
    double x = warp_sum(part);
    if (bias != nullptr) x += bias[i] - bias[j];
    acc += log(1.0 + exp(-x));
  }
  if (lane == 0) sh[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) s += sh[w];
    partial[blockIdx.x] = s;
  }
}

}  // namespace qmfb
