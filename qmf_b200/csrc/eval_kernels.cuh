// Ranking evaluation on the GPU (sm_100a): a users x items score GEMM on the FP64 tensor cores fused
// with the per-user rank statistics that AUC / AP / P@k / R@k need; the nT x nitems score matrix is
// never materialised.  Replaces
//   Engine::computeTestScores                     qmf/Engine.cpp:73-96     (K7)
//   AUC / Precision / Recall / AveragePrecision   qmf/metrics/Metrics.cpp:65-164 (K8, the sort)
//
// What must be exact.  Per test user t with positives P (test items with label > 0) and negatives N
// (all other items, train positives included, Engine.cpp:58-69) the metrics depend only on
//   cnt[i], i = 0..nP  = #{ x in N : exactly i positives score strictly less than s_x }
// where s is the REFERENCE score: bias_i + sum_f U[u,f] * V[i,f] accumulated in f order with
// separately rounded multiply and add (Engine.cpp:85-91).  Under the reference order (score
// descending, positives first on ties, Metrics.cpp:85-86) a negative in bucket i is preceded by exactly
// nP - i positives; the host turns cnt into AUC / AP / P@k / R@k with the reference's own arithmetic.
//
// How it is made fast AND exact (north_star part 3):
//   1. eval_pos_kernel: the nP positives' scores of every test user in the reference's exact order
//      (__dmul_rn / __dadd_rn, bit-identical), sorted ascending.
//   2. eval_score_kernel: scores of ALL items by DMMA (mma.sync.m8n8k4.f64): a CTA keeps 64 users'
//      factors in shared memory and streams item tiles through a cp.async ring.  A DMMA score s^ differs
//      from the reference score s by at most
//           eps = 4 (k + 4) 2^-53 (|bias| + ||u||_2 ||v||_2)
//      (both are floating-point evaluations of the same k-term dot product: |s - S|, |s^ - S| <=
//      gamma_{k+1} (|b| + sum |u_f v_f|), Cauchy-Schwarz, and a factor 2 of slack for the rounded norms).
//      If no positive's exact score lies in [s^ - eps, s^ + eps] the bucket of the item is decided by
//      s^ exactly as it would be by s.  Otherwise (a tie, a positive item itself, or a 1e-13 near-miss)
//      the pair goes to a per-tile list and is re-scored in the reference's exact order.  Every integer
//      is therefore the one the reference's sort would produce.
//   3. the positives are taken out of their own buckets, the counts go to global memory.
// Work units are (64 test users) x (an item range), dealt dynamically, so that a few hundred test
// users still fill 148 SMs; counts of the item ranges of one user add up in global memory (integers:
// deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmfb {

// D(8x8) += A(8x4) * B(4x8), FP64 (SASS: DMMA.8x8x4).  Lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2*(T%4)+{0,1}].
__device__ __forceinline__ void ev_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
// 16-byte asynchronous copy global -> shared (SASS: LDGSTS), L1 bypass
__device__ __forceinline__ void ev_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src)
               : "memory");
}

constexpr int kEvThreads = 256;
constexpr int kEvU = 64;        // test users per work unit (8 DMMA row tiles)
constexpr int kEvI = 64;        // items per tile
constexpr int kEvKC = 32;       // factors per staged chunk
constexpr int kEvLDB = kEvKC + 4;  // stage row stride in doubles: == 4 mod 16 -> conflict-free fragment loads
constexpr int kEvSpCap = 2048;  // sorted positives' scores of one unit held in shared memory (else: global)
constexpr int kEvLinear = 32;   // up to this many positives are counted linearly in the epilogue (else: binary search)
constexpr int kEvPosSmem = 4096;  // eval_pos_kernel: positives of one user sorted in shared memory

struct EvalParams {
  const double* U;          // user factors, row stride ldu >= KP (KP = k rounded up to 32), pad columns zero
  int64_t ldu;
  const double* V;          // item factors, row stride ldv >= KP, pad columns zero
  int64_t ldv;
  const double* bias;       // item biases or nullptr
  int k, kp;
  int nitems;
  const int32_t* test_users;   // nT user idx
  int nT;
  const int64_t* label_ptr;    // nT + 1: offsets into label_items and (shifted by t) into cnt
  const int32_t* label_items;  // per user: positive item idx, ascending
  int32_t* cnt;                // out: per user nP + 1 counters at offset label_ptr[t] + t (zeroed by the launcher)
  double* pos_scores;          // out: per user the positives' scores in ascending order (offset label_ptr[t])
  int32_t* pos_bucket;         // per positive (label order): # positives of its user scoring strictly less
  const double* vnorm;         // ||v_x||_2 per item
  int nsplit;                  // item ranges per user group
  int* unit_counter;           // dynamic unit scheduler
  int stages;
  int sort_in_kernel;          // 0: eval_pos_kernel leaves the scores unsorted (a segmented sort follows)
};

__host__ __device__ inline size_t eval_smem_bytes(int kp, int stages, int ng) {
  size_t b = size_t(kEvU) * (kp + 4) * 8;              // U tile
  b += size_t(ng) * stages * kEvI * kEvLDB * 8;        // item chunk ring of every warp group
  b += size_t(kEvSpCap) * 8;                           // sorted positives
  b += size_t(kEvSpCap + kEvU) * 4;                    // bucket counters
  b += size_t(kEvU) * (8 + 8 + 4 + 4);                 // unorm, sp offset(int64), nP, user idx
  b += size_t(ng) * kEvU * kEvI * 2;                   // re-score list of every warp group (uint16)
  b += 64;                                             // unit id, flags, slot total, list counts
  return b;
}

// ||v_x||_2 of every item row (one warp per row)
__global__ void eval_item_norm_kernel(const double* __restrict__ V, int64_t ldv, int k, int nitems, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t x = w0; x < nitems; x += nw) {
    double s = 0.0;
    for (int f = lane; f < k; f += 32) {
      const double v = V[x * ldv + f];
      s = fma(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[x] = sqrt(s);
  }
}

// the reference's score of (user row u, item x): Engine.cpp:85-91, bit-identical
__device__ __forceinline__ double eval_exact_score(const double* __restrict__ u, const double* __restrict__ v, double b, int k) {
  double s = b;
  for (int f = 0; f < k; ++f) s = __dadd_rn(s, __dmul_rn(u[f], v[f]));
  return s;
}

__device__ __forceinline__ int eval_lower_bound(const double* sp, int n, double s) {  // #{m : sp[m] < s}
  int a = 0, b = n;
  while (a < b) {
    const int m = (a + b) >> 1;
    if (sp[m] < s) a = m + 1; else b = m;
  }
  return a;
}

// ---- 1. positives: exact scores, sorted ascending, one CTA per test user (grid-stride) -------------
__global__ void __launch_bounds__(256) eval_pos_kernel(const EvalParams prm) {
  __shared__ double sp[kEvPosSmem];
  const int tid = threadIdx.x;
  for (int t = blockIdx.x; t < prm.nT; t += gridDim.x) {
    const int64_t lp0 = prm.label_ptr[t];
    const int nP = int(prm.label_ptr[t + 1] - lp0);
    if (nP == 0) continue;
    const double* u = prm.U + int64_t(prm.test_users[t]) * prm.ldu;
    const bool in_smem = nP <= kEvPosSmem && prm.sort_in_kernel;
    int n2 = 1;
    while (n2 < nP) n2 <<= 1;
    __syncthreads();  // previous user's sort is done with sp
    for (int i = tid; i < (in_smem ? n2 : nP); i += blockDim.x) {
      double s = __longlong_as_double(0x7ff0000000000000LL);  // +inf padding
      if (i < nP) {
        const int item = prm.label_items[lp0 + i];
        s = eval_exact_score(u, prm.V + int64_t(item) * prm.ldv, prm.bias != nullptr ? prm.bias[item] : 0.0, prm.k);
      }
      if (in_smem) sp[i] = s; else prm.pos_scores[lp0 + i] = s;
    }
    if (!in_smem) continue;  // sorted afterwards by a device-wide segmented sort
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = tid; i < n2; i += blockDim.x) {
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
    for (int i = tid; i < nP; i += blockDim.x) prm.pos_scores[lp0 + i] = sp[i];
  }
}

// bucket of every positive ITEM (label order) among its user's sorted positives: the score kernel counts
// positives like negatives and takes them out again at exactly this bucket.  One thread per label.
__global__ void eval_pos_bucket_kernel(const EvalParams prm, int64_t nlabels) {
  for (int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; q < nlabels; q += int64_t(gridDim.x) * blockDim.x) {
    int lo = 0, hi = prm.nT;  // user t with label_ptr[t] <= q < label_ptr[t + 1]
    while (hi - lo > 1) {
      const int m = (lo + hi) >> 1;
      if (prm.label_ptr[m] <= q) lo = m; else hi = m;
    }
    const int64_t lp0 = prm.label_ptr[lo];
    const int nP = int(prm.label_ptr[lo + 1] - lp0);
    const int item = prm.label_items[q];
    const double s = eval_exact_score(prm.U + int64_t(prm.test_users[lo]) * prm.ldu, prm.V + int64_t(item) * prm.ldv,
                                      prm.bias != nullptr ? prm.bias[item] : 0.0, prm.k);
    prm.pos_bucket[q] = eval_lower_bound(prm.pos_scores + lp0, nP, s);
  }
}

// ---- 2. all items by DMMA + buckets -----------------------------------------------------------------
// NG warp groups of 8 warps share one CTA (one per SM) and one 64-user tile, and take the item tiles of the unit in
// turn (group g: tiles g, g + NG, ...) with their own cp.async ring, re-score list and NAMED barrier: nothing makes
// them run in lockstep, so one group's epilogue (search + shared atomics, no DMMA) overlaps the other's DMMA main
// loop - with a single group the tensor pipe idled through every epilogue (34.8 % DMMA-active in
// profiles/r02_eval_large_ncu.csv).  NG = 2 whenever the shared memory allows it (k <= 128).
template <int NG>
__global__ void __launch_bounds__(kEvThreads * NG) eval_score_kernel(const EvalParams prm) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NTH = kEvThreads * NG;
  const int KP = prm.kp, LDU = KP + 4, NS = prm.stages;
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = warp >> 3, gw = warp & 7, gtid = tid & (kEvThreads - 1);
  double* utile = reinterpret_cast<double*>(smem);
  double* ring = utile + size_t(kEvU) * LDU + size_t(g) * NS * kEvI * kEvLDB;
  double* sps = utile + size_t(kEvU) * LDU + size_t(NG) * NS * kEvI * kEvLDB;
  int* cnts = reinterpret_cast<int*>(sps + kEvSpCap);
  double* unorm = reinterpret_cast<double*>(cnts + kEvSpCap + kEvU);
  int64_t* lp0s = reinterpret_cast<int64_t*>(unorm + kEvU);
  int* nPs = reinterpret_cast<int*>(lp0s + kEvU);
  int* soff = nPs + kEvU;                      // offset of the user's positives / counters in sps / cnts
  uint16_t* list = reinterpret_cast<uint16_t*>(soff + kEvU) + size_t(g) * kEvU * kEvI;
  int* misc = reinterpret_cast<int*>(reinterpret_cast<uint16_t*>(soff + kEvU) + size_t(NG) * kEvU * kEvI);  // [1] unit, [2] positives-in-smem flag, [3] slots
  int* lcount = misc + 8 + g;                  // this group's re-score list length
  auto group_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kEvThreads) : "memory"); };
  const int uh = gw & 1, iq = gw >> 1;         // users 32 uh .. +31 (4 row tiles), items 16 iq .. +15 (2 column tiles)
  const int ngroups = (prm.nT + kEvU - 1) / kEvU;
  const int nunits = ngroups * prm.nsplit;
  const int nchunks = KP / kEvKC;
  const double ceps = 4.0 * double(prm.k + 4) * 1.1102230246251565e-16;  // 4 (k + 4) 2^-53

  for (;;) {
    __syncthreads();  // everyone is done with the previous unit's shared memory
    if (tid == 0) misc[1] = atomicAdd(prm.unit_counter, 1);
    __syncthreads();
    const int unit = misc[1];
    if (unit >= nunits) break;
    const int grp = unit / prm.nsplit, split = unit % prm.nsplit;
    const int t0 = grp * kEvU, nU = min(kEvU, prm.nT - t0);
    const int xb = int(int64_t(prm.nitems) * split / prm.nsplit), xe = int(int64_t(prm.nitems) * (split + 1) / prm.nsplit);

    // ---- unit prologue: user rows, norms, positives, counters ------------------------------------------
    if (tid < kEvU) {
      const int64_t lp0 = tid < nU ? prm.label_ptr[t0 + tid] : 0;
      lp0s[tid] = lp0;
      nPs[tid] = tid < nU ? int(prm.label_ptr[t0 + tid + 1] - lp0) : -1;
    }
    __syncthreads();
    if (warp == 0) {  // soff = exclusive prefix sum of (nP + 1) over the 64 users (two per lane)
      const int a = max(nPs[2 * lane], 0) + 1, b = max(nPs[2 * lane + 1], 0) + 1;
      int incl = a + b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      soff[2 * lane] = incl - a - b;
      soff[2 * lane + 1] = incl - b;
      if (lane == 31) {
        misc[2] = (incl <= kEvSpCap) ? 1 : 0;
        misc[3] = incl;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) misc[8 + gg] = 0;
      }
    }
    {  // U tile: 16-byte cp.async, rows of users past the end are rows of the last valid user (never bucketed)
      const int ppr = KP / 2;
      for (int q = tid; q < kEvU * ppr; q += NTH) {
        const int r = q / ppr, piece = q % ppr;
        const int t = t0 + min(r, nU - 1);
        ev_cp_async16(utile + size_t(r) * LDU + piece * 2, prm.U + int64_t(prm.test_users[t]) * prm.ldu + piece * 2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const bool in_smem = misc[2] != 0;
    const int ntot = misc[3];  // sum over the unit's users of (nP + 1)
    auto owner = [&](int qq) {  // user u with soff[u] <= qq < soff[u + 1]
      int lo = 0, hi = kEvU;
      while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (soff[m] <= qq) lo = m; else hi = m;
      }
      return lo;
    };
    {  // ||u||_2 (4 threads per user), counters, sorted positives
      if (tid < 4 * kEvU) {  // whole warps
        const int r = tid >> 2, q = tid & 3;
        double s = 0.0;
        for (int f = q; f < prm.k; f += 4) {
          const double v = utile[size_t(r) * LDU + f];
          s = fma(v, v, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) unorm[r] = sqrt(s);
      }
      if (in_smem) {  // flattened over (user, slot): no serial loop over the 64 users
        for (int qq = tid; qq < ntot; qq += NTH) {
          cnts[qq] = 0;
          const int u = owner(qq), i = qq - soff[u];
          if (i < nPs[u]) sps[qq] = prm.pos_scores[lp0s[u] + i];
        }
      }
    }
    __syncthreads();

    // ---- item tiles ------------------------------------------------------------------------------------
    const int ntiles = (xe - xb + kEvI - 1) / kEvI;
    const int ntiles_g = ntiles > g ? (ntiles - g + NG - 1) / NG : 0;  // this group's tiles: g, g + NG, ...
    const int total = ntiles_g * nchunks;
    auto issue = [&](int idx) {  // chunk idx = (local tile) * nchunks + c
      if (idx < total) {
        const int tile = g + NG * (idx / nchunks), c = idx % nchunks;
        double* st = ring + size_t(idx % NS) * kEvI * kEvLDB;
#pragma unroll
        for (int m = 0; m < (kEvI * kEvKC / 2) / kEvThreads; ++m) {
          const int q = gtid + kEvThreads * m, r = q / (kEvKC / 2), piece = q % (kEvKC / 2);
          const int x = min(xb + tile * kEvI + r, prm.nitems - 1);
          ev_cp_async16(st + r * kEvLDB + piece * 2, prm.V + int64_t(x) * prm.ldv + c * kEvKC + piece * 2);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");  // always: uniform group accounting
    };
    for (int i = 0; i < NS - 1; ++i) issue(i);

    int c0[4] = {0, 0, 0, 0};  // bucket-0 hits of this lane's four user rows over the whole unit
    for (int lt = 0; lt < ntiles_g; ++lt) {
      const int tile = g + NG * lt;
      const int x0 = xb + tile * kEvI;
      double acc[4][2][2];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
      // this lane's 4 item columns: norms and biases for the epilogue (loaded early, used late)
      double vn[2][2], vb[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int x = x0 + 16 * iq + 8 * nt + 2 * (lane & 3) + e;
          const bool ok = x < xe;
          vn[nt][e] = ok ? __ldg(prm.vnorm + x) : 0.0;
          vb[nt][e] = (ok && prm.bias != nullptr) ? __ldg(prm.bias + x) : 0.0;
        }
      for (int c = 0; c < nchunks; ++c) {
        const int idx = lt * nchunks + c;
        if (NS == 4) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        group_bar();  // chunk idx landed for the whole group; the stage consumed in the previous iteration is free
        issue(idx + NS - 1);
        const double* st = ring + size_t(idx % NS) * kEvI * kEvLDB + size_t(16 * iq + (lane >> 2)) * kEvLDB + (lane & 3);
        const double* ua = utile + size_t(32 * uh + (lane >> 2)) * LDU + c * kEvKC + (lane & 3);
#pragma unroll
        for (int s = 0; s < kEvKC / 4; ++s) {
          double a[4], b[2];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) a[mt] = ua[size_t(8 * mt) * LDU + 4 * s];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) b[nt] = st[8 * nt * kEvLDB + 4 * s];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) ev_dmma(acc[mt][nt], a[mt], b[nt]);
        }
      }
      // ---- epilogue: bucket the 16 scores of this lane --------------------------------------------------
      // `sp` / `cn` are passed from two call sites so that the common case (the unit's positives and counters in
      // shared memory) compiles to LDS / ATOMS instead of generic loads: the profile of the first version had 55 %
      // of the warp samples in this block, most of them on the dependent generic loads of a binary search.
      auto bucket_row = [&](int mt, int ul, int nP, const double* sp, int* cn) {
        const double un = unorm[ul];
        double lo[4], hi[4];
        int a[4], b[4];
        bool live[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int nt = q >> 1, e = q & 1;
          const double s = acc[mt][nt][e] + vb[nt][e];
          const double eps = ceps * (fabs(vb[nt][e]) + un * vn[nt][e]);
          lo[q] = s - eps;
          hi[q] = s + eps;
          a[q] = 0;
          b[q] = nP;
          live[q] = (x0 + 16 * iq + 8 * nt + 2 * (lane & 3) + e) < xe;
        }
        if (nP <= kEvLinear) {  // few positives: count them - independent loads and compares, no dependent search steps
#pragma unroll 4
          for (int m = 0; m < nP; ++m) {
            const double v = sp[m];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] += v < lo[q] ? 1 : 0;
          }
        } else {  // four lower_bound(sp, lo) searches in lockstep (independent loads in flight)
          for (int span = nP; span > 0; span >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (a[q] < b[q]) {
                const int m = (a[q] + b[q]) >> 1;
                if (sp[m] < lo[q]) a[q] = m + 1; else b[q] = m;
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!live[q]) continue;
          if (a[q] < nP && sp[a[q]] <= hi[q]) {
            // a positive's score within the error bound: re-score this pair in the reference's exact order
            const int xl = 16 * iq + 8 * (q >> 1) + 2 * (lane & 3) + (q & 1);
            list[atomicAdd(lcount, 1)] = uint16_t((ul << 8) | xl);
          } else if (a[q] == 0) {
            ++c0[mt];  // below every positive - by far the most common bucket of a trained model: counted in a register
          } else {
            atomicAdd(cn + a[q], 1);
          }
        }
      };
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int ul = 32 * uh + 8 * mt + (lane >> 2);
        const int nP = nPs[ul];
        if (nP < 0) continue;
        if (in_smem) {
          bucket_row(mt, ul, nP, sps + soff[ul], cnts + soff[ul]);
        } else {
          bucket_row(mt, ul, nP, prm.pos_scores + lp0s[ul], prm.cnt + lp0s[ul] + (t0 + ul));
        }
      }
      group_bar();
      const int nlist = *lcount;
      if (nlist > 0) {
        for (int e = gtid; e < nlist; e += kEvThreads) {
          const int ul = list[e] >> 8, x = x0 + (list[e] & 255);
          const double s = eval_exact_score(utile + size_t(ul) * LDU, prm.V + int64_t(x) * prm.ldv,
                                            prm.bias != nullptr ? prm.bias[x] : 0.0, prm.k);
          const double* sp = in_smem ? sps + soff[ul] : prm.pos_scores + lp0s[ul];
          int* cn = in_smem ? cnts + soff[ul] : prm.cnt + lp0s[ul] + (t0 + ul);
          atomicAdd(cn + eval_lower_bound(sp, nPs[ul], s), 1);
        }
        group_bar();
        if (gtid == 0) *lcount = 0;
      }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {  // the register-held bucket-0 counts join the unit's counters
      const int ul = 32 * uh + 8 * mt + (lane >> 2);
      if (nPs[ul] >= 0 && c0[mt] != 0) atomicAdd(in_smem ? cnts + soff[ul] : prm.cnt + lp0s[ul] + (t0 + ul), c0[mt]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- the positive ITEMS of this item range were bucketed like negatives: take them out again, at the
    //      bucket eval_pos_bucket_kernel computed from their exact score; then the unit's counters go to
    //      global memory (the item ranges of one user add up).  Flattened over (user, slot).
    for (int qq = tid; qq < ntot; qq += NTH) {
      const int u = owner(qq), i = qq - soff[u];
      if (i < nPs[u]) {
        const int item = prm.label_items[lp0s[u] + i];
        if (item >= xb && item < xe) {
          int* cn = in_smem ? cnts + soff[u] : prm.cnt + lp0s[u] + (t0 + u);
          atomicSub(cn + prm.pos_bucket[lp0s[u] + i], 1);
        }
      }
    }
    if (in_smem) {
      __syncthreads();
      for (int qq = tid; qq < ntot; qq += NTH) {
        const int u = owner(qq), i = qq - soff[u];
        if (u < nU && i <= nPs[u]) {
          const int v = cnts[qq];
          if (v != 0) atomicAdd(prm.cnt + lp0s[u] + (t0 + u) + i, v);
        }
      }
    }
  }
}

}  // namespace qmfb
