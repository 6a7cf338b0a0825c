// WALS half-step kernels for sm_100a (B200): Gram matrix + fused per-row normal-equation build,
// blocked Cholesky solve and loss.  Replaces the reference's CPU loops
//   WALSEngine::computeXtX           qmf/wals/WALSEngine.cpp:246-264   (K1)
//   WALSEngine::updateFactorsForOne  qmf/wals/WALSEngine.cpp:266-310   (K2 build, K3 solve, K4 loss)
//   linearSymmetricSolve / dsysv_    qmf/Matrix.cpp:81-96              (K3)
//
// Design:
//  * FP64 tensor cores: B200 has no tcgen05 FP64 kind; the native FP64 MMA is DMMA.8x8x4
//    (mma.sync.m8n8k4.f64).  Measured peak 37.1 TFLOP/s (profiles/r01_fp64_peak.txt); the
//    shared-memory-fed loop below reaches ~35 TFLOP/s in isolation.
//  * A row's k x k system is held as the upper triangle of 8x8 tiles.  Warp w of a CTA owns
//    tile-rows w and NT-1-w (NT+1 tiles -> perfectly balanced), accumulators live in registers.
//  * Gathered factor rows are staged into padded shared memory (row stride KP+8 doubles ->
//    conflict-free DMMA fragment loads) by the TMA engine: one cp.async.bulk per gathered row,
//    completion on an mbarrier, NSTAGE-deep ring, no register staging.
//  * After the build the tiles are written to shared memory (aliasing the staging ring), the
//    Gram matrix and lambda are added, and a blocked right-looking Cholesky (8-wide panels,
//    DMMA trailing updates, b carried as an extra tile column so the forward solve is free)
//    followed by a single-warp blocked back substitution produces x.  A never leaves the SM.
//  * Rows are scheduled longest-first through an atomic counter (persistent CTAs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qmfb {

constexpr int kChunk = 16;   // gathered rows per pipeline stage
constexpr int kStages = 4;   // ring depth

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D(8x8) += A(8x4) * B(4x8), FP64.  Lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2*(T%4)+{0,1}].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// c[m][n] += sum_kk ta[kk][m] * tb[kk][n] for two 8x8 row-major tiles in shared memory
// (conflict-free: lanes read 32 consecutive doubles per fragment).
__device__ __forceinline__ void tile_mma_tn(double (&c)[2], const double* ta, const double* tb, int lane, double sa) {
  const int o = (lane & 3) * 8 + (lane >> 2);
  dmma(c, sa * ta[o], tb[o]);
  dmma(c, sa * ta[o + 32], tb[o + 32]);
}

// ------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------
template <int NT>
struct WalsSmem {
  static constexpr int KP = NT * 8;            // padded factor dimension
  static constexpr int LD = KP + 8;            // staging row stride (doubles); LD % 16 == 8
  static constexpr int NWARPS = NT / 2;
  static constexpr int NTHREADS = NWARPS * 32;  // == 2 * KP
  static constexpr int NTILE_A = NT * (NT + 1) / 2;
  static constexpr int NTILE = NTILE_A + NT;    // + one tile column for b
  static constexpr size_t kStageBytes = size_t(kStages) * kChunk * LD * 8;
  static constexpr size_t kTileBytes = size_t(NTILE) * 64 * 8;
  static constexpr size_t kMainBytes = kStageBytes > kTileBytes ? kStageBytes : kTileBytes;
  static constexpr size_t kOffWts = kMainBytes;                              // kStages*2*kChunk doubles
  static constexpr size_t kOffW = kOffWts + size_t(kStages) * 2 * kChunk * 8;  // NT inverse diagonal tiles
  static constexpr size_t kOffB = kOffW + size_t(NT) * 64 * 8;               // b copy (KP)
  static constexpr size_t kOffX = kOffB + size_t(KP) * 8;                    // x (KP)
  static constexpr size_t kOffR = kOffX + size_t(KP) * 8;                    // back-substitution rhs (KP)
  static constexpr size_t kOffBh = kOffR + size_t(KP) * 8;                   // b half sums (2*KP)
  static constexpr size_t kOffBar = kOffBh + size_t(KP) * 16;                // full[kStages], empty[kStages]
  static constexpr size_t kOffMisc = kOffBar + size_t(kStages) * 16;
  static constexpr size_t kBytes = kOffMisc + 64;

  // tile (I,J), I <= J <= NT (J == NT is the b column), row-major upper storage
  __host__ __device__ static constexpr int tidx(int I, int J) { return I * (NT + 1) - I * (I - 1) / 2 + (J - I); }
  // packed Gram tile (I,J), I <= J < NT
  __host__ __device__ static constexpr int gidx(int I, int J) { return I * NT - I * (I - 1) / 2 + (J - I); }
};

struct RowSource {
  const double* Y;       // right-side factors, row stride ldy (>= KP, zero padded)
  int64_t ldy;
  const int32_t* col;    // gather indices (nullptr => contiguous rows p0..p1 of Y, Gram mode)
  const double* val;
  double alpha;
};

// ------------------------------------------------------------------------------------------
// build: accumulate sum_s wa[s] * y_s y_s^T (upper tiles) and, in gather mode, b = sum_s wb[s] y_s
// ------------------------------------------------------------------------------------------
template <int NT, bool GATHER>
struct Producer {
  using SM = WalsSmem<NT>;
  const double* src;  // next chunk's gathered row for this lane
  double wa, wb;
  double csum;

  __device__ __forceinline__ void load(const RowSource& rs, int64_t p0, int64_t p1, int n, int lane) {
    const int64_t p = p0 + int64_t(n) * kChunk + lane;
    const bool valid = lane < kChunk && p < p1;
    if (GATHER) {
      const int32_t c = valid ? __ldg(rs.col + p) : 0;
      const double v = valid ? __ldg(rs.val + p) : 0.0;
      src = rs.Y + int64_t(c) * rs.ldy;
      wa = valid ? rs.alpha * v : 0.0;          // WALSEngine.cpp:282  alpha * r
      wb = valid ? 1.0 + rs.alpha * v : 0.0;    // WALSEngine.cpp:280  1 + alpha * r
    } else {
      src = rs.Y + (valid ? p : 0) * rs.ldy;
      wa = valid ? 1.0 : 0.0;
      wb = 0.0;
    }
  }

  // issue chunk n of the current row (global chunk number gn) and prefetch chunk n+1's indices
  __device__ __forceinline__ void issue(unsigned char* smem, const RowSource& rs, int64_t p0, int64_t p1, int n,
                                        uint32_t gn, int lane) {
    double* stagebuf = reinterpret_cast<double*>(smem);
    double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
    uint64_t* empty = full + kStages;
    const uint32_t st = gn % kStages;
    if (gn >= kStages) mbar_wait(&empty[st], ((gn / kStages) & 1u) ^ 1u);
    if (lane < kChunk) {
      wts[st * 2 * kChunk + lane] = wa;
      wts[st * 2 * kChunk + kChunk + lane] = wb;
    }
    csum += wb;
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(&full[st], uint32_t(kChunk) * SM::KP * 8);
    __syncwarp();
    if (lane < kChunk) bulk_g2s(stagebuf + (size_t(st) * kChunk + lane) * SM::LD, src, SM::KP * 8, &full[st]);
    load(rs, p0, p1, n + 1, lane);
  }
};

template <int NT, int W, bool GATHER, typename Epilogue>
__device__ __forceinline__ void build_warp(unsigned char* smem, const RowSource& rs, int64_t p0, int64_t p1,
                                           uint32_t chunk_base, double& csum_out, Epilogue&& epilogue) {
  using SM = WalsSmem<NT>;
  constexpr int I0 = W, I1 = NT - 1 - W, N0 = NT - I0, N1 = NT - I1, D = I1 - I0;
  const int lane = threadIdx.x & 31;
  const double* stagebuf = reinterpret_cast<const double*>(smem);
  const double* wts = reinterpret_cast<const double*>(smem + SM::kOffWts);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + kStages;

  double acc0[N0][2], acc1[N1][2];
#pragma unroll
  for (int j = 0; j < N0; ++j) acc0[j][0] = acc0[j][1] = 0.0;
#pragma unroll
  for (int j = 0; j < N1; ++j) acc1[j][0] = acc1[j][1] = 0.0;
  double bacc = 0.0;
  const int bi = threadIdx.x % SM::KP, bh = threadIdx.x / SM::KP;

  const int nch = int((p1 - p0 + kChunk - 1) / kChunk);
  Producer<NT, GATHER> prod;
  prod.csum = 0.0;
  if (W == 0) {
    prod.load(rs, p0, p1, 0, lane);
    const int pre = nch < kStages - 1 ? nch : kStages - 1;
    for (int n = 0; n < pre; ++n) prod.issue(smem, rs, p0, p1, n, chunk_base + n, lane);
  }
  for (int c = 0; c < nch; ++c) {
    const uint32_t gc = chunk_base + c;
    const uint32_t st = gc % kStages;
    if (W == 0) {
      const int n = c + kStages - 1;
      if (n < nch) prod.issue(smem, rs, p0, p1, n, chunk_base + n, lane);
    }
    mbar_wait(&full[st], (gc / kStages) & 1u);
    const double* sb = stagebuf + size_t(st) * kChunk * SM::LD;
    const double* wt = wts + st * 2 * kChunk;
#pragma unroll
    for (int s0 = 0; s0 < kChunk; s0 += 4) {
      const double* p = sb + (s0 + (lane & 3)) * SM::LD + (lane >> 2) + 8 * I0;
      const double wa = wt[s0 + (lane & 3)];
      double bf[N0];
#pragma unroll
      for (int j = 0; j < N0; ++j) bf[j] = p[8 * j];
      const double a0 = bf[0] * wa;
      const double a1 = bf[D] * wa;
#pragma unroll
      for (int j = 0; j < N0; ++j) dmma(acc0[j], a0, bf[j]);
#pragma unroll
      for (int j = 0; j < N1; ++j) dmma(acc1[j], a1, bf[D + j]);
    }
    if (GATHER) {
#pragma unroll
      for (int s = 0; s < kChunk / 2; ++s) bacc += wt[kChunk + 2 * s + bh] * sb[(2 * s + bh) * SM::LD + bi];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  if (W == 0) csum_out = prod.csum;
  epilogue(acc0, acc1, bacc);
}

// ------------------------------------------------------------------------------------------
// Gram kernel: partial upper-tile Gram of rows [r0, r1) per CTA, then a deterministic reduce
// ------------------------------------------------------------------------------------------
template <int NT, int W>
struct GramEpilogue {
  double* out;  // this CTA's packed partial (NTILE_A * 64 doubles)
  template <int N0, int N1>
  __device__ __forceinline__ void operator()(double (&acc0)[N0][2], double (&acc1)[N1][2], double) const {
    using SM = WalsSmem<NT>;
    const int lane = threadIdx.x & 31;
    constexpr int I0 = W, I1 = NT - 1 - W;
#pragma unroll
    for (int j = 0; j < N0; ++j) {
      *reinterpret_cast<double2*>(out + size_t(SM::gidx(I0, I0 + j)) * 64 + lane * 2) = make_double2(acc0[j][0], acc0[j][1]);
    }
#pragma unroll
    for (int j = 0; j < N1; ++j) {
      *reinterpret_cast<double2*>(out + size_t(SM::gidx(I1, I1 + j)) * 64 + lane * 2) = make_double2(acc1[j][0], acc1[j][1]);
    }
  }
};

template <int NT, int W>
__device__ __forceinline__ void gram_dispatch(int warp, unsigned char* smem, const RowSource& rs, int64_t r0,
                                              int64_t r1, double* out) {
  if constexpr (W < NT / 2) {
    if (warp == W) {
      double csum;
      build_warp<NT, W, false>(smem, rs, r0, r1, 0u, csum, GramEpilogue<NT, W>{out});
    } else {
      gram_dispatch<NT, W + 1>(warp, smem, rs, r0, r1, out);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(WalsSmem<NT>::NTHREADS) gram_partial_kernel(const double* __restrict__ Y, int64_t ldy,
                                                                              int64_t row_begin, int64_t row_end,
                                                                              double* __restrict__ partial) {
  using SM = WalsSmem<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + kStages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SM::NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t n = row_end - row_begin;
  const int64_t r0 = row_begin + n * blockIdx.x / gridDim.x;
  const int64_t r1 = row_begin + n * (blockIdx.x + 1) / gridDim.x;
  RowSource rs{Y, ldy, nullptr, nullptr, 0.0};
  gram_dispatch<NT, 0>(threadIdx.x >> 5, smem, rs, r0, r1, partial + size_t(blockIdx.x) * SM::NTILE_A * 64);
}

// out[t] = sum_b partial[b][t] in fixed order (deterministic), t over NTILE_A*64 packed entries
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nparts, int nelem, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nelem) return;
  double s = 0.0;
  for (int b = 0; b < nparts; ++b) s += partial[size_t(b) * nelem + t];
  out[t] = s;
}

// packed upper tiles -> full symmetric k x k row-major (for the host-visible Gram / tests)
template <int NT>
__global__ void gram_unpack_kernel(const double* __restrict__ packed, int k, double* __restrict__ out) {
  using SM = WalsSmem<NT>;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k * k) return;
  const int i = idx / k, j = idx % k;
  const int a = i < j ? i : j, b = i < j ? j : i;
  out[idx] = packed[size_t(SM::gidx(a >> 3, b >> 3)) * 64 + (a & 7) * 8 + (b & 7)];
}

// ------------------------------------------------------------------------------------------
// fused per-row kernel
// ------------------------------------------------------------------------------------------
struct SolveParams {
  double* X;              // left factors being solved (row stride ldx), global row = row_offset + local row
  int64_t ldx;
  int64_t row_offset;
  const double* Y;        // fixed right factors (row stride ldy)
  int64_t ldy;
  int k;                  // true number of factors (<= KP)
  const int64_t* row_ptr; // local CSR offsets (nrows + 1)
  const int32_t* col;
  const double* val;
  const int32_t* order;   // local rows, longest first
  int nrows;
  const double* gram;     // packed upper tiles of Y^T Y (all rows of Y)
  double alpha, lambda;
  double* row_loss;       // per local row loss term (WALSEngine.cpp:295-304)
  int* counter;           // dynamic scheduler
  int* error;             // set to 1 if a pivot is not positive (reference: CHECK_EQ(result, 0), Matrix.cpp:94)
};

template <int NT, int W>
struct SolveEpilogue {
  unsigned char* smem;
  const double* gram;
  double lambda;
  int k;
  template <int N0, int N1>
  __device__ __forceinline__ void operator()(double (&acc0)[N0][2], double (&acc1)[N1][2], double bacc) const {
    using SM = WalsSmem<NT>;
    const int lane = threadIdx.x & 31;
    constexpr int I0 = W, I1 = NT - 1 - W;
    // all warps must be done reading the staging ring before tiles (aliased) are written
    reinterpret_cast<double*>(smem + SM::kOffBh)[threadIdx.x] = bacc;
    __syncthreads();
    double* tiles = reinterpret_cast<double*>(smem);
    const int r = lane >> 2, c0 = 2 * (lane & 3);
    auto put = [&](int I, int J, double (&a)[2]) {
      const double2 g = *reinterpret_cast<const double2*>(gram + size_t(SM::gidx(I, J)) * 64 + lane * 2);
      double v0 = a[0] + g.x, v1 = a[1] + g.y;
      if (I == J) {  // A(i,i) += lambda (WALSEngine.cpp:290-292); padded dimensions get a unit pivot
        const int gi = 8 * I + r;
        if (c0 == r) v0 = gi < k ? v0 + lambda : 1.0;
        if (c0 + 1 == r) v1 = gi < k ? v1 + lambda : 1.0;
      }
      *reinterpret_cast<double2*>(tiles + size_t(SM::tidx(I, J)) * 64 + lane * 2) = make_double2(v0, v1);
    };
#pragma unroll
    for (int j = 0; j < N0; ++j) put(I0, I0 + j, acc0[j]);
#pragma unroll
    for (int j = 0; j < N1; ++j) put(I1, I1 + j, acc1[j]);
  }
};

template <int NT, int W>
__device__ __forceinline__ void solve_build_dispatch(int warp, unsigned char* smem, const RowSource& rs, int64_t p0,
                                                     int64_t p1, uint32_t chunk_base, double& csum,
                                                     const SolveParams& prm) {
  if constexpr (W < NT / 2) {
    if (warp == W) {
      build_warp<NT, W, true>(smem, rs, p0, p1, chunk_base, csum, SolveEpilogue<NT, W>{smem, prm.gram, prm.lambda, prm.k});
    } else {
      solve_build_dispatch<NT, W + 1>(warp, smem, rs, p0, p1, chunk_base, csum, prm);
    }
  }
}

// One warp: factor the 8x8 diagonal tile (upper Cholesky A = U^T U) in C-fragment layout with
// shuffles and write W = inv(U) (row-major, upper) to wtile.  Returns false on a bad pivot.
__device__ __forceinline__ bool factor_diag_tile(const double* tile, double* wtile, int lane) {
  const int r = lane >> 2, q = lane & 3;
  double2 a = *reinterpret_cast<const double2*>(tile + lane * 2);
  double a0 = a.x, a1 = a.y;
  double e0 = (2 * q == r) ? 1.0 : 0.0, e1 = (2 * q + 1 == r) ? 1.0 : 0.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double piv = __shfl_sync(0xffffffffu, (j & 1) ? a1 : a0, 4 * j + (j >> 1));
    ok = ok && (piv > 0.0);
    const double d = rsqrt(piv);
    if (r == j) {
      a0 *= d; a1 *= d; e0 *= d; e1 *= d;
    }
    const double uc0 = __shfl_sync(0xffffffffu, a0, 4 * j + q);
    const double uc1 = __shfl_sync(0xffffffffu, a1, 4 * j + q);
    const double ec0 = __shfl_sync(0xffffffffu, e0, 4 * j + q);
    const double ec1 = __shfl_sync(0xffffffffu, e1, 4 * j + q);
    const double t0 = __shfl_sync(0xffffffffu, a0, 4 * j + (r >> 1));
    const double t1 = __shfl_sync(0xffffffffu, a1, 4 * j + (r >> 1));
    const double ur = (r & 1) ? t1 : t0;
    if (r > j) {
      a0 -= ur * uc0; a1 -= ur * uc1; e0 -= ur * ec0; e1 -= ur * ec1;
    }
  }
  // e = inv(U)^T (lower triangular); store W = e^T
  wtile[(2 * q) * 8 + r] = e0;
  wtile[(2 * q + 1) * 8 + r] = e1;
  return ok;
}

template <int NT>
__global__ void __launch_bounds__(WalsSmem<NT>::NTHREADS, (NT >= 12 ? 2 : (NT >= 8 ? 4 : 8)))
  wals_solve_kernel(const SolveParams prm) {
  using SM = WalsSmem<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  double* tiles = reinterpret_cast<double*>(smem);
  double* wt = reinterpret_cast<double*>(smem + SM::kOffW);
  double* bcopy = reinterpret_cast<double*>(smem + SM::kOffB);
  double* xvec = reinterpret_cast<double*>(smem + SM::kOffX);
  double* rvec = reinterpret_cast<double*>(smem + SM::kOffR);
  double* bhalf = reinterpret_cast<double*>(smem + SM::kOffBh);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + kStages;
  int* misc = reinterpret_cast<int*>(smem + SM::kOffMisc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SM::NWARPS);
    }
    mbar_fence_init();
  }
  uint32_t chunk_base = 0;
  RowSource rs{prm.Y, prm.ldy, prm.col, prm.val, prm.alpha};

  for (;;) {
    __syncthreads();  // previous row fully retired (tiles, xvec, misc free)
    if (tid == 0) misc[0] = atomicAdd(prm.counter, 1);
    fence_proxy_async();  // order this row's generic writes to the aliased ring before the next bulk copies
    __syncthreads();
    const int slot = misc[0];
    if (slot >= prm.nrows) break;
    const int row = prm.order[slot];
    const int64_t p0 = prm.row_ptr[row], p1 = prm.row_ptr[row + 1];
    const int nch = int((p1 - p0 + kChunk - 1) / kChunk);

    // ---- build: accumulators -> (+ Gram, + lambda) -> tiles in shared memory ------------------
    double csum = 0.0;
    solve_build_dispatch<NT, 0>(warp, smem, rs, p0, p1, chunk_base, csum, prm);
    chunk_base += uint32_t(nch);
    // b = sum of the two half sums; b column tiles (column 0 = b, other columns 0); keep a copy
    if (tid < SM::KP) {
      const double b = bhalf[tid] + bhalf[tid + SM::KP];
      bcopy[tid] = b;
      double* bt = tiles + size_t(SM::tidx(tid >> 3, NT)) * 64 + (tid & 7) * 8;
      bt[0] = b;
#pragma unroll
      for (int c = 1; c < 8; ++c) bt[c] = 0.0;
    }
    __syncthreads();

    // ---- blocked Cholesky, panel width 8; forward substitution rides along in column NT ------
    bool ok = true;
    if (warp == 0) ok = factor_diag_tile(tiles + size_t(SM::tidx(0, 0)) * 64, wt, lane);
    for (int I = 0; I < NT; ++I) {
      __syncthreads();  // W_I ready, row I of tiles final up to panel I-1
      // (b) panel: U[I][J] = inv(U_II)^T * A[I][J]  for J = I+1 .. NT
      for (int J = I + 1 + warp; J <= NT; J += SM::NWARPS) {
        double* t = tiles + size_t(SM::tidx(I, J)) * 64;
        double c[2] = {0.0, 0.0};
        tile_mma_tn(c, wt + I * 64, t, lane, 1.0);
        __syncwarp();
        *reinterpret_cast<double2*>(t + lane * 2) = make_double2(c[0], c[1]);
      }
      if (I == NT - 1) break;
      __syncthreads();
      // (c) trailing update: A[J1][J2] -= U[I][J1]^T U[I][J2], I < J1 <= J2 <= NT, J1 < NT.
      //     The warp that owns the next diagonal tile updates it first and factors it right away
      //     (look-ahead) while the other warps sweep the rest.
      const int T = NT - 1 - I;               // rows J1 = I+1 .. NT-1
      const int ntr = T * (T + 1) / 2 + T;    // tiles incl. b column
      const int dwarp = (I + 1) % SM::NWARPS;
      const int nw = SM::NWARPS > 1 ? SM::NWARPS - 1 : 1;
      const int wslot = SM::NWARPS > 1 ? (warp > dwarp ? warp - 1 : warp) : 0;
      if (SM::NWARPS == 1 || warp == dwarp) {
        double* t = tiles + size_t(SM::tidx(I + 1, I + 1)) * 64;
        const double* u = tiles + size_t(SM::tidx(I, I + 1)) * 64;
        double2 cv = *reinterpret_cast<double2*>(t + lane * 2);
        double c[2] = {cv.x, cv.y};
        tile_mma_tn(c, u, u, lane, -1.0);
        *reinterpret_cast<double2*>(t + lane * 2) = make_double2(c[0], c[1]);
        __syncwarp();
        ok = factor_diag_tile(t, wt + (I + 1) * 64, lane) && ok;
      }
      if (SM::NWARPS == 1 || warp != dwarp) {
        // linear tile index e in [1, ntr): row J1 has (NT - J1 + 1) tiles; e == 0 is the diagonal tile above
        int J1 = I + 1, base = 0;
        for (int e = 1 + wslot; e < ntr; e += nw) {
          while (e - base >= NT - J1 + 1) {
            base += NT - J1 + 1;
            ++J1;
          }
          const int J2 = J1 + (e - base);
          double* t = tiles + size_t(SM::tidx(J1, J2)) * 64;
          double2 cv = *reinterpret_cast<double2*>(t + lane * 2);
          double c[2] = {cv.x, cv.y};
          tile_mma_tn(c, tiles + size_t(SM::tidx(I, J1)) * 64, tiles + size_t(SM::tidx(I, J2)) * 64, lane, -1.0);
          *reinterpret_cast<double2*>(t + lane * 2) = make_double2(c[0], c[1]);
        }
      }
    }
    if (!ok && lane == 0) *prm.error = 1;
    __syncthreads();

    // ---- back substitution U x = z (warp 0), loss, store ----------------------------------------
    if (warp == 0) {
      for (int i = lane; i < SM::KP; i += 32) rvec[i] = tiles[size_t(SM::tidx(i >> 3, NT)) * 64 + (i & 7) * 8];
      __syncwarp();
      for (int I = NT - 1; I >= 0; --I) {
        if (lane < 8) {  // x_I = W_I * r_I (W upper triangular)
          const double* w = wt + I * 64 + lane * 8;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            s0 += w[c] * rvec[8 * I + c];
            s1 += w[c + 1] * rvec[8 * I + c + 1];
          }
          xvec[8 * I + lane] = s0 + s1;
        }
        __syncwarp();
        // r_J -= U[J][I] x_I for J < I: one tile per iteration, lane reads its C-fragment pair
        const double x0 = xvec[8 * I + 2 * (lane & 3)], x1 = xvec[8 * I + 2 * (lane & 3) + 1];
        for (int J = 0; J < I; ++J) {
          const double2 u = *reinterpret_cast<const double2*>(tiles + size_t(SM::tidx(J, I)) * 64 + lane * 2);
          double pr = u.x * x0 + u.y * x1;
          pr += __shfl_xor_sync(0xffffffffu, pr, 1);
          pr += __shfl_xor_sync(0xffffffffu, pr, 2);
          if ((lane & 3) == 0) rvec[8 * J + (lane >> 2)] -= pr;
        }
        __syncwarp();
      }
      // loss term: c + x^T B x - 2 x^T b with x^T B x = z^T z - lambda x^T x (WALSEngine.cpp:295-304)
      double part = 0.0;
      for (int i = lane; i < prm.k; i += 32) {
        const double z = tiles[size_t(SM::tidx(i >> 3, NT)) * 64 + (i & 7) * 8];
        const double x = xvec[i];
        part += z * z - prm.lambda * x * x - 2.0 * x * bcopy[i];
      }
      part += csum;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) prm.row_loss[row] = part;
      double* xr = prm.X + (prm.row_offset + row) * prm.ldx;
      for (int i = lane; i < SM::KP; i += 32) xr[i] = i < prm.k ? xvec[i] : 0.0;
    }
  }
}

// deterministic sum of n doubles (fixed strided order + fixed tree), single block
__global__ void sum_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace qmfb
