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
//  * Factor rows are staged into padded shared memory (row stride KP+4 doubles -> conflict-free
//    DMMA fragment loads) through an mbarrier-tracked ring, no register staging.  Gathered rows
//    (solve kernel) use 16-byte cp.async (LDGSTS), one warp per chunk in rotation: the TMA engine's
//    outstanding-request window caps a DRAM-resident 1 KB-row gather at ~5 B/clk/SM
//    (profiles/r01_solve_v1_*: 49 % of warp samples waiting on the full barrier), LDGSTS has no
//    such cap.  The Gram kernel streams contiguous rows with TMA bulk copies (cp.async.bulk).
//  * Accumulators start from the Gram tiles, b is accumulated as two more DMMA tiles per warp; after
//    the build lambda is added and the tiles go to shared memory (XOR-swizzled 8x8 tiles: every
//    operand-fragment load is conflict-free), where a blocked right-looking Cholesky (8-wide panels,
//    DMMA trailing updates, look-ahead on the diagonal tile, b carried as an extra tile column so the
//    forward solve is free) and a blocked back substitution produce x.  A never leaves the SM.
//    (k > 128: wals_big.cuh keeps the tiles in an L2-resident workspace and shares solve_row.)
//  * Multi-GPU: the solved row is stored to the local replica AND to the peers' replicas (NVLink
//    peer memory) by the same warps - the all-gather of the half-step is part of this kernel.
//  * Two persistent CTAs per SM at k=128 (tile storage aliases the staging ring): while one CTA
//    is in its latency-bound solve phase the other keeps the DMMA pipe busy building.  The solve
//    phase is written for a short FP64 dependency chain (fraction-free 8x8 pivot block, see
//    factor_diag_tile) because its scalar FP64 ops queue behind the other CTA's DMMAs.
//  * Rows are sorted longest-first and dealt to the CTAs in serpentine order (static schedule,
//    so the next row is known early enough to prefetch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

namespace qmfb {

constexpr int kChunk = 16;   // gathered rows per pipeline stage
// trailing-update tiles in flight per warp.  ONE on purpose: the sweep has the whole duration of the
// diagonal-tile factor to finish, and bursts of DMMAs / shared-memory loads from it delay that factor's
// critical path on the shared FP64 pipe (4 in flight: 27.3 ms on the user-shaped half of C4, 1: 25.3 ms;
// tools/exp_user.py)
constexpr int kTU = 1;
#ifndef QMFB_B_DFMA
// 1: b = sum_s (1 + alpha r_s) y_s accumulated with DFMA, every warp 16 columns (8 LDS + 8 DFMA per lane and
//    chunk); 0: as two extra DMMA tiles per warp (7/8 of their columns are zeros: +12 % tensor-pipe time).
//    Measured on the item-shaped half of C4: 14.17 ms vs 15.03 ms (tools/exp_user.py); user half unchanged.
#define QMFB_B_DFMA 1
#endif
#ifndef QMFB_KSTAGES
#define QMFB_KSTAGES 5
#endif
constexpr int kStages = QMFB_KSTAGES;   // ring depth (k > 64)
// k <= 64 (NT <= 8): the rows are short and cheap, what counts is how many are in flight per SM - a shallower ring and a
// tighter register budget buy more resident CTAs (measured on C3 / C1)
// (gpurun r02_occ_variants.log, C3 epoch: ring 5 / 4 CTAs 10.85 ms, ring 4 / 5 CTAs 10.57, ring 3 / 5 CTAs 10.48, ring 3 / 6 CTAs 10.26)
#ifndef QMFB_RING_SMALL
#define QMFB_RING_SMALL 3
#endif
#ifndef QMFB_OCC8
#define QMFB_OCC8 6    // resident CTAs per SM the NT = 8 kernel is compiled for (80 registers, 32.5 KB)
#endif
#ifndef QMFB_OCC4
#define QMFB_OCC4 12   // ... the NT = 4 kernel (79 registers, 17.3 KB)
#endif

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
// 16-byte asynchronous copy global -> shared (SASS: LDGSTS), L1 bypass
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D(8x8) += A(8x4) * B(4x8), FP64.  Lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2*(T%4)+{0,1}].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// shared-memory layout
// ------------------------------------------------------------------------------------------
template <int NT>
struct WalsSmem {
  static constexpr int kNT = NT;
  static constexpr int KP = NT * 8;            // padded factor dimension
  // staging row stride (doubles).  A 64-bit shared load is served per half-warp (16 lanes x 8 B =
  // one 128-byte wavefront); the DMMA fragment address is (lane%4)*LD + lane/4, so LD % 16 == 4
  // gives 16 distinct bank pairs per half-warp (LD % 16 == 8 costs two wavefronts per half-warp:
  // measured 41 % conflict wavefronts, profiles/r01_build_v4).
  static constexpr int LD = KP + 4;
  static constexpr int NWARPS = NT / 2;
  static constexpr int NTHREADS = NWARPS * 32;  // == 2 * KP
  static constexpr int NTILE_A = NT * (NT + 1) / 2;
  static constexpr int NTILE = NTILE_A + NT;    // + one tile column for b
  static constexpr int kRing = NT <= 8 ? QMFB_RING_SMALL : kStages;  // ring depth of the gather pipeline
  static constexpr int kAhead = kRing - 2;      // chunks in flight beyond the one being consumed: the stage refilled
                                                // was consumed TWO chunks ago (nobody waits for the slowest warp)
  static constexpr size_t kStageBytes = size_t(kRing) * kChunk * LD * 8;
  static constexpr size_t kTileBytes = size_t(NTILE) * 64 * 8;
  static constexpr size_t kMainBytes = kStageBytes > kTileBytes ? kStageBytes : kTileBytes;
  static constexpr size_t kOffStage = 0;
  static constexpr size_t kOffTiles = 0;                                     // aliases the staging ring
  static constexpr size_t kOffWts = kMainBytes;                              // kRing*2*kChunk doubles
  static constexpr size_t kOffW = kOffWts + size_t(kRing) * 2 * kChunk * 8;  // NT inverse diagonal tiles
  static constexpr size_t kOffB = kOffW + size_t(NT) * 64 * 8;               // b copy (KP)
  static constexpr size_t kOffX = kOffB + size_t(KP) * 8;                    // x (KP)
  static constexpr size_t kOffR = kOffX + size_t(KP) * 8;                    // back-substitution rhs (8)
  static constexpr size_t kOffFs = kOffR + 64;                               // solve barriers F/S/G (8 doubles) + pivot-row scratch (16 doubles)
  static constexpr size_t kOffBh = kOffFs + 256;                             // per-warp partial sum of (1 + alpha r) (NWARPS, padded to 8)
  static constexpr size_t kOffBar = kOffBh + 64;                             // full[kRing], empty[kRing]
  static constexpr size_t kOffRow = kOffBar + size_t(kRing) * 16;            // 2 row slots x 32 bytes
  static constexpr size_t kBytes = kOffRow + 64;

  // tile (I,J), I <= J <= NT (J == NT is the b column), row-major upper storage
  __host__ __device__ static constexpr int tidx(int I, int J) { return I * (NT + 1) - I * (I - 1) / 2 + (J - I); }
  // packed Gram tile (I,J), I <= J < NT
  __host__ __device__ static constexpr int gidx(int I, int J) { return I * NT - I * (I - 1) / 2 + (J - I); }
};

// Layout of the WARP-SPECIALISED solve kernel (wals_solve_ws_kernel): one CTA per SM made of
//   NT/2 builder warps  - gather + DMMA build of row r into registers, then into tile buffer r % 2
//   2 groups of NT/4 solver warps - group g runs the Cholesky / back substitution / loss / store of the
//                         rows r = g (mod 2) out of tile buffer g while the builders are already on the
//                         next rows
// Nothing aliases: gather ring | 2 tile buffers | per-group solve scratch.  Three rows are in flight per
// SM (one building, two solving) instead of two, and the DMMA pipe is never idle because a CTA is in its
// latency-bound solve phase.
#ifndef QMFB_WS_RING
#define QMFB_WS_RING 3
#endif
template <int NT>
struct WalsSmemWs {
  static constexpr int kNT = NT;
  static constexpr int KP = NT * 8;
  static constexpr int LD = KP + 4;
  static constexpr int NWARPS = NT / 2;                 // builder warps (same warp roles as WalsSmem)
  static constexpr int NSOLVE = NT / 4;                 // warps per solver group: 32 * NSOLVE == KP threads
  static constexpr int NTHREADS = 32 * (NWARPS + 2 * NSOLVE);
  static constexpr int NTILE_A = NT * (NT + 1) / 2;
  static constexpr int NTILE = NTILE_A + NT;
  static constexpr int kRing = QMFB_WS_RING;
  static constexpr int kAhead = kRing - 1;      // one builder group per SM: deeper prefetch, the refilled stage was
                                                // consumed in the previous chunk (the gathering warp may wait briefly)
  static constexpr size_t kStageBytes = size_t(kRing) * kChunk * LD * 8;
  static constexpr size_t kTileBytes = size_t(NTILE) * 64 * 8;
  static constexpr size_t kOffStage = 0;
  static constexpr size_t kOffTiles = kStageBytes;                               // 2 buffers
  static constexpr size_t kOffWts = kOffTiles + 2 * kTileBytes;                  // kRing * 2 * kChunk doubles
  static constexpr size_t kOffW = kOffWts + size_t(kRing) * 2 * kChunk * 8;      // 2 groups x NT inverse diagonal tiles
  static constexpr size_t kOffB = kOffW + 2 * size_t(NT) * 64 * 8;               // 2 x b copy (KP)
  static constexpr size_t kOffX = kOffB + 2 * size_t(KP) * 8;                    // 2 x x (KP)
  static constexpr size_t kOffR = kOffX + 2 * size_t(KP) * 8;                    // 2 x back-substitution rhs (8)
  static constexpr size_t kOffFs = kOffR + 2 * 64;                               // 2 x (solve barriers + pivot-row scratch), 32 doubles each
  static constexpr size_t kOffBh = kOffFs + 2 * 256;                             // 2 x partial sums of (1 + alpha r) (8)
  static constexpr size_t kOffBar = kOffBh + 2 * 64;                             // full[kRing], empty[kRing], tfull[2], tempty[2]
  static constexpr size_t kOffRow = kOffBar + ((size_t(2 * kRing + 4) * 8 + 15) / 16) * 16;  // 2 row slots x 32 bytes
  static constexpr size_t kBytes = kOffRow + 64;
  static_assert(NT % 4 == 0, "NT must be a multiple of 4");
  __host__ __device__ static constexpr int tidx(int I, int J) { return I * (NT + 1) - I * (I - 1) / 2 + (J - I); }
  __host__ __device__ static constexpr int gidx(int I, int J) { return I * NT - I * (I - 1) / 2 + (J - I); }
};

// ------------------------------------------------------------------------------------------
// DMMA inner loop of one staged chunk for warp role W (each warp owns NT+1 of the upper tiles).
// This is the only code that differs between warps; everything else is a single copy.
// ------------------------------------------------------------------------------------------
// Row-pair tiling: warp W owns tile rows I0 = W (NT-W tiles) and I1 = NT-1-W (W+1 tiles);
// acc[0..N0) are row I0, acc[N0..NT+1) row I1.
// With WITH_B the right-hand side rides along on the tensor pipe: acc[NT+1] / acc[NT+2] are the b
// column tiles of tile rows I0 / I1,  b(8I+m) = sum_s y_s(8I+m) * wb_s, i.e. the operand fragment
// already in registers times a B fragment that holds wb_s in column 0 and zeros elsewhere - exactly
// the (column 0 = b) tile the blocked Cholesky carries as column NT.
template <int NT, int W, bool WITH_B, class SM = WalsSmem<NT>>
__device__ __forceinline__ void chunk_mma_rows(double (&acc)[NT + 3][2], const double* sb, const double* wt, int lane) {
  constexpr int I0 = W, I1 = NT - 1 - W, N0 = NT - I0, N1 = NT - I1, D = I1 - I0;
#pragma unroll
  for (int s0 = 0; s0 < kChunk; s0 += 4) {
    const double* p = sb + (s0 + (lane & 3)) * SM::LD + (lane >> 2) + 8 * I0;
    const double wa = wt[s0 + (lane & 3)];
    double wbf = 0.0;
    if constexpr (WITH_B) {
      const double wb = wt[kChunk + s0 + (lane & 3)];
      wbf = (lane >> 2) == 0 ? wb : 0.0;
    }
    double bf[N0];
#pragma unroll
    for (int j = 0; j < N0; ++j) bf[j] = p[8 * j];
    const double a0 = bf[0] * wa;
    const double a1 = bf[D] * wa;
#pragma unroll
    for (int j = 0; j < N0; ++j) dmma(acc[j], a0, bf[j]);
    if constexpr (WITH_B) dmma(acc[NT + 1], bf[0], wbf);
#pragma unroll
    for (int j = 0; j < N1; ++j) dmma(acc[N0 + j], a1, bf[D + j]);
    if constexpr (WITH_B) dmma(acc[NT + 2], bf[D], wbf);
  }
}

template <int NT, int W, bool WITH_B, class SM = WalsSmem<NT>>
__device__ __forceinline__ void chunk_mma_dispatch(int warp, double (&acc)[NT + 3][2], const double* sb,
                                                   const double* wt, int lane) {
  if constexpr (W < NT / 2) {
    if (warp == W) {
      chunk_mma_rows<NT, W, WITH_B, SM>(acc, sb, wt, lane);
    } else {
      chunk_mma_dispatch<NT, W + 1, WITH_B, SM>(warp, acc, sb, wt, lane);
    }
  }
}

// tile coordinates of this warp's t-th accumulator
template <int NT>
__device__ __forceinline__ void acc_tile(int warp, int t, int& I, int& J) {
  const int n0 = NT - warp;
  if (t < n0) {
    I = warp;
    J = warp + t;
  } else {
    I = NT - 1 - warp;
    J = I + (t - n0);
  }
}

// 8x8 tiles in shared memory are stored row-major with the column index XOR-swizzled by bit 1 of the
// row:  (a, b) -> a*8 + (b ^ tile_sw(a)).  A half-warp then reads the DMMA operand fragment
// (a = lane%4 [+4], b = lane/4) from 16 distinct 8-byte bank pairs; plain row-major puts rows a and
// a+2 on the same banks (2-way conflict on every operand load of the panel / trailing update:
// 4.7e9 excess wavefronts per C4 user half-step in profiles/r01_solve_final_ncu.csv).
__device__ __forceinline__ int tile_sw(int a) { return ((a >> 1) & 1) << 2; }
// operand-fragment offset: element (lane%4, lane/4); the k+4 half is at +32
__device__ __forceinline__ int tile_frag_off(int lane) { return (lane & 3) * 8 + ((lane >> 2) ^ tile_sw(lane & 3)); }
// accumulator-fragment offset (double2): element (lane/4, 2*(lane%4))
__device__ __forceinline__ int tile_acc_off(int lane) { return (lane >> 2) * 8 + ((2 * (lane & 3)) ^ tile_sw(lane >> 2)); }

// ------------------------------------------------------------------------------------------
// Gram kernel: partial upper-tile Gram of rows [r0, r1) per CTA (contiguous rows streamed by TMA
// bulk copies, one per row), then a deterministic reduce
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(WalsSmem<NT>::NTHREADS) gram_partial_kernel(const double* __restrict__ Y, int64_t ldy,
                                                                              int64_t row_begin, int64_t row_end,
                                                                              double* __restrict__ partial, int part0,
                                                                              int nparts) {
  // CTA b sums part (part0 + b) of the `nparts` fixed row parts of [row_begin, row_end) into
  // partial[part0 + b]: the parts and the order inside a part depend only on the row range, so the
  // reduced Gram is bit-identical however the parts are dealt to devices (qmfb_wals_sharded_*)
  using SM = WalsSmem<NT>;
  constexpr int kStages = SM::kRing;  // (shadows the global ring depth: this layout's)
  extern __shared__ __align__(128) unsigned char smem[];
  double* stagebuf = reinterpret_cast<double*>(smem + SM::kOffStage);
  double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + kStages;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SM::NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t n = row_end - row_begin;
  const int part = part0 + int(blockIdx.x);
  const int64_t r0 = row_begin + n * part / nparts;
  const int64_t r1 = row_begin + n * (part + 1) / nparts;
  const int nch = int((r1 - r0 + kChunk - 1) / kChunk);

  auto issue = [&](int c) {  // warp 0 only
    const uint32_t st = c % kStages;
    if (c >= kStages) mbar_wait(&empty[st], ((c / kStages) & 1u) ^ 1u);
    const int64_t p = r0 + int64_t(c) * kChunk + lane;
    const bool valid = lane < kChunk && p < r1;
    if (lane < kChunk) wts[st * 2 * kChunk + lane] = valid ? 1.0 : 0.0;
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(&full[st], uint32_t(kChunk) * SM::KP * 8);
    __syncwarp();
    if (lane < kChunk) bulk_g2s(stagebuf + (size_t(st) * kChunk + lane) * SM::LD, Y + (valid ? p : r0) * ldy, SM::KP * 8, &full[st]);
  };

  double acc[NT + 3][2];
#pragma unroll
  for (int t = 0; t < NT + 3; ++t) acc[t][0] = acc[t][1] = 0.0;
  if (warp == 0) {
    for (int c = 0; c < nch && c < kStages - 1; ++c) issue(c);
  }
  for (int c = 0; c < nch; ++c) {
    const uint32_t st = c % kStages;
    if (warp == 0 && c + kStages - 1 < nch) issue(c + kStages - 1);
    mbar_wait(&full[st], (c / kStages) & 1u);
    chunk_mma_dispatch<NT, 0, false>(warp, acc, stagebuf + size_t(st) * kChunk * SM::LD, wts + st * 2 * kChunk, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  double* out = partial + size_t(part) * SM::NTILE_A * 64;
#pragma unroll
  for (int t = 0; t <= NT; ++t) {
    int I, J;
    acc_tile<NT>(warp, t, I, J);
    *reinterpret_cast<double2*>(out + size_t(SM::gidx(I, J)) * 64 + lane * 2) = make_double2(acc[t][0], acc[t][1]);
  }
}

// out[t] = sum_p partial[p][t] over parts p = 0 .. nparts-1 in that order (deterministic), t over the
// NTILE_A*64 packed entries.  Part p lives in the workspace of the device that computed it:
// src[d] for part_end[d-1] <= p < part_end[d] (local or NVLink peer memory; one device: nsrc == 1).
constexpr int kMaxGramSrc = 16;
struct GramReduceParams {
  const double* src[kMaxGramSrc];
  int part_end[kMaxGramSrc];
  int nsrc;
  int nelem;
};
__global__ void gram_reduce_kernel(const GramReduceParams prm, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= prm.nelem) return;
  double s = 0.0;
  int p = 0;
  for (int d = 0; d < prm.nsrc; ++d) {
    const double* src = prm.src[d];
    for (; p < prm.part_end[d]; ++p) s += src[size_t(p) * prm.nelem + t];
  }
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
// Extremely long rows (blockbuster items of a power-law catalogue): a row is normally built by ONE
// CTA, so a row of several hundred thousand entries would be the tail of its half-step (and of every
// rank's, multi-GPU).  Every row at the head of `order` (longest first) with >= kLongRow entries is
// therefore cut into segments - a work list of (row, segment) units, any number of rows up to the
// buffer bounds below - that the persistent CTAs of long_row_partial_kernel build ahead of the solve
// kernel (the Gram kernel's TMA scheme on gathered, weighted rows) and that long_row_reduce_kernel
// sums in a fixed order; the solve kernel then starts such a row from Gram + that sum and skips its
// own gather loop.  The segment length of a row is a function of ITS length only, so its sum (and with
// it the factors and the loss) does not depend on how rows are sharded over GPUs.  The plan is made
// on the device (no host synchronisation); without long rows the cost is three near-empty launches.
// ------------------------------------------------------------------------------------------
constexpr int kLongMaxRows = 512;      // candidate rows: positions [0, kLongMaxRows) of `order`
constexpr int kLongMaxSegs = 2048;     // work units the partial buffer holds
constexpr int kLongMaxParts = 256;     // segments per row
constexpr int64_t kLongRow = 32768;    // entries from which a row counts as extremely long
constexpr int64_t kLongSegMin = 4096;  // entries per segment, at least

__host__ __device__ inline int64_t long_seg_len(int64_t len) {
  int64_t s = (len + kLongMaxParts - 1) / kLongMaxParts;
  s = (s + kChunk - 1) / kChunk * kChunk;
  return s < kLongSegMin ? kLongSegMin : s;
}

struct LongPlan {  // device memory, written by long_row_plan_kernel
  int nlong;       // rows order[0 .. nlong) are prebuilt
  int nseg;        // work units
  int pad[2];
  int seg_begin[kLongMaxRows + 1];  // first work unit of row position r
};

template <int NT>
struct LongRow {
  static constexpr int kLen = WalsSmem<NT>::NTILE_A * 64 + WalsSmem<NT>::KP + 8;  // tiles | b | sum of (1 + alpha r), padded
};

struct LongRowParams {
  const double* Y;
  int64_t ldy;
  const int64_t* row_ptr;
  const int32_t* col;
  const double* val;
  const int32_t* order;
  int nrows;
  double alpha;
  double* partial;  // [kLongMaxSegs][kLen]
  double* sum;      // [kLongMaxRows][kLen]
  LongPlan* plan;
};

// one CTA of kLongMaxRows threads: the longest prefix of `order` whose rows are all long and whose
// segments fit the buffer
__global__ void __launch_bounds__(kLongMaxRows) long_row_plan_kernel(const LongRowParams prm) {
  __shared__ int cum[kLongMaxRows];
  const int t = threadIdx.x;
  int parts = kLongMaxSegs + 1;  // "not long": ends the prefix
  if (t < prm.nrows) {
    const int row = prm.order[t];
    const int64_t len = prm.row_ptr[row + 1] - prm.row_ptr[row];
    if (len >= kLongRow) parts = int((len + long_seg_len(len) - 1) / long_seg_len(len));
  }
  cum[t] = parts;
  __syncthreads();
  for (int o = 1; o < kLongMaxRows; o <<= 1) {  // inclusive scan (values stay < 2^31: 512 * 2049)
    const int v = t >= o ? cum[t - o] : 0;
    __syncthreads();
    cum[t] += v;
    __syncthreads();
  }
  const bool ok = cum[t] <= kLongMaxSegs;
  const int nlong = __syncthreads_count(ok);  // cum is increasing: the ok positions are a prefix
  if (ok) prm.plan->seg_begin[t] = cum[t] - parts;
  if (t == nlong - 1) prm.plan->seg_begin[nlong] = cum[t];
  if (t == 0) {
    prm.plan->nlong = nlong;
    prm.plan->nseg = nlong > 0 ? cum[nlong - 1] : 0;
    if (nlong == 0) prm.plan->seg_begin[0] = 0;
  }
}

template <int NT>
__global__ void __launch_bounds__(WalsSmem<NT>::NTHREADS) long_row_partial_kernel(const LongRowParams prm) {
  using SM = WalsSmem<NT>;
  const int nseg = prm.plan->nseg;
  if (int(blockIdx.x) >= nseg) return;
  const int nlong = prm.plan->nlong;
  constexpr int kStages = SM::kRing;  // (shadows the global ring depth: this layout's)
  extern __shared__ __align__(128) unsigned char smem[];
  double* stagebuf = reinterpret_cast<double*>(smem + SM::kOffStage);
  double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + kStages;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SM::NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t cbase = 0;  // chunks this CTA has streamed so far: the ring's stage / phase run on across work units
  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    // work unit -> (row position r, segment j): last r with seg_begin[r] <= seg
    int lo = 0, hi = nlong - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (prm.plan->seg_begin[mid] <= seg) lo = mid; else hi = mid - 1;
    }
    const int r = lo, j = seg - prm.plan->seg_begin[r];
    const int row = prm.order[r];
    const int64_t p0 = prm.row_ptr[row], p1 = prm.row_ptr[row + 1];
    const int64_t sl = long_seg_len(p1 - p0);
    const int64_t s0 = p0 + int64_t(j) * sl, s1 = s0 + sl < p1 ? s0 + sl : p1;
    const int nch = int((s1 - s0 + kChunk - 1) / kChunk);
    double csum = 0.0;

    auto issue = [&](int c) {  // warp 0 only: one TMA bulk copy per gathered row
      const uint32_t gc = cbase + uint32_t(c);
      const uint32_t st = gc % kStages;
      if (gc >= kStages) mbar_wait(&empty[st], ((gc / kStages) & 1u) ^ 1u);
      const int64_t p = s0 + int64_t(c) * kChunk + lane;
      const bool valid = lane < kChunk && p < s1;
      const double v = valid ? prm.val[p] : 0.0;
      const int32_t src = valid ? prm.col[p] : 0;
      if (lane < kChunk) {
        const double wb = valid ? 1.0 + prm.alpha * v : 0.0;        // WALSEngine.cpp:280
        wts[st * 2 * kChunk + lane] = valid ? prm.alpha * v : 0.0;  // WALSEngine.cpp:282
        wts[st * 2 * kChunk + kChunk + lane] = wb;
        csum += wb;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&full[st], uint32_t(kChunk) * SM::KP * 8);
      __syncwarp();
      if (lane < kChunk) bulk_g2s(stagebuf + (size_t(st) * kChunk + lane) * SM::LD, prm.Y + int64_t(src) * prm.ldy, SM::KP * 8, &full[st]);
    };

    double acc[NT + 3][2];
#pragma unroll
    for (int t = 0; t < NT + 3; ++t) acc[t][0] = acc[t][1] = 0.0;
    double bacc = 0.0;
    if (warp == 0) {
      for (int c = 0; c < nch && c < kStages - 1; ++c) issue(c);
    }
    for (int c = 0; c < nch; ++c) {
      const uint32_t gc = cbase + uint32_t(c);
      const uint32_t st = gc % kStages;
      if (warp == 0 && c + kStages - 1 < nch) issue(c + kStages - 1);
      mbar_wait(&full[st], (gc / kStages) & 1u);
      const double* sb = stagebuf + size_t(st) * kChunk * SM::LD;
      const double* w8 = wts + st * 2 * kChunk;
      {
        const double* pb = sb + (lane >> 4) * SM::LD + warp * 16 + (lane & 15);
        const double* pw = w8 + kChunk + (lane >> 4);
#pragma unroll
        for (int jj = 0; jj < kChunk / 2; ++jj) bacc = fma(pw[2 * jj], pb[2 * jj * SM::LD], bacc);
      }
      chunk_mma_dispatch<NT, 0, false>(warp, acc, sb, w8, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    cbase += uint32_t(nch);
    double* out = prm.partial + size_t(seg) * LongRow<NT>::kLen;
#pragma unroll
    for (int t = 0; t <= NT; ++t) {
      int I, J;
      acc_tile<NT>(warp, t, I, J);
      *reinterpret_cast<double2*>(out + size_t(SM::gidx(I, J)) * 64 + lane * 2) = make_double2(acc[t][0], acc[t][1]);
    }
    bacc += __shfl_xor_sync(0xffffffffu, bacc, 16);
    if (lane < 16) out[SM::NTILE_A * 64 + warp * 16 + lane] = bacc;
    if (warp == 0) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
      if (lane == 0) out[SM::NTILE_A * 64 + SM::KP] = csum;
    }
  }
}

// sum[r][t] = sum over the segments of long row r, fixed order (deterministic)
template <int NT>
__global__ void long_row_reduce_kernel(const LongRowParams prm) {
  const int nlong = prm.plan->nlong;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int kLen = LongRow<NT>::kLen;
  if (t >= kLen) return;
  for (int r = blockIdx.y; r < nlong; r += gridDim.y) {
    const int b0 = prm.plan->seg_begin[r], b1 = prm.plan->seg_begin[r + 1];
    const double* p = prm.partial + t;
    double s = 0.0;
    for (int b = b0; b < b1; ++b) s += p[size_t(b) * kLen];
    prm.sum[size_t(r) * kLen + t] = s;
  }
}

// ------------------------------------------------------------------------------------------
// fused per-row kernel
// ------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 15;  // other ranks of one box

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
  int* error;             // set to 1 if a pivot is not positive (reference: CHECK_EQ(result, 0), Matrix.cpp:94)
  // Fused all-gather of the solved shard (multi-GPU): every solved row is ALSO stored into the same
  // row of the other ranks' replicas of X, straight from the solve kernel over NVLink peer memory
  // (same ldx / row_offset), instead of a separate collective after the kernel.
  int npeers;
  double* peerX[kMaxPeers];
  const double* long_sum;  // [kLongMaxRows][LongRow<NT>::kLen] prebuilt sums of the extremely long rows, or nullptr
  const LongPlan* long_plan;  // which rows those are (device memory; read when long_sum != nullptr)
  // 1: gather the factor rows with TMA bulk copies (one 8 KP-byte cp.async.bulk per row, issued by 16 lanes)
  // instead of 16-byte cp.async (32 per lane and chunk).  Cheap to issue, but the TMA engine's outstanding-request
  // window starves the loop when the gathered matrix is DRAM resident (profiles/r01_solve_v1_*): the launcher
  // sets it only when Y fits the L2 comfortably (user half-step of C4: 18 MB of item factors).
  int tma_gather;
};

// Store the solved row (KP doubles in shared memory) to X and to every peer replica.  Target t is
// written by warp (first + t) mod nwarps; posted stores, nothing waits for NVLink.
__device__ __forceinline__ void store_solved_row(const SolveParams& prm, const double* xvec, int64_t grow, int KP, int warp,
                                                 int lane, int first, int nwarps) {
  for (int t = 0; t <= prm.npeers; ++t) {
    if ((first + t) % nwarps != warp) continue;
    double* xr = (t == 0 ? prm.X : prm.peerX[t - 1]) + grow * prm.ldx;
    for (int i = lane; i < KP; i += 32) xr[i] = i < prm.k ? xvec[i] : 0.0;
  }
}



// One warp: factor the 8x8 diagonal tile A = U^T U and write W^T = inv(U)^T (swizzled tile) to
// wtile.  C-fragment layout: lane holds row lane/4, columns 2*(lane%4)+{0,1}; an identity is
// eliminated alongside.  The elimination is FRACTION-FREE so that the pivot-to-pivot dependency
// chain is one shuffle + three FP64 ops instead of a reciprocal/rsqrt sequence:
//     a_rc <- a_rc * pn - (a_jr * a_jc) * 2^-e,   p = a_jj = pn * 2^e, pn in [1, 2)
// which keeps every remaining row scaled by S = prod(pn) (<= 2^8; the power-of-two factors are
// exact).  Row r of the true factor is recovered at the end with a single
// g_r = rsqrt(p_r * S_r):  U[r][c] = a_rc * g_r,  inv(U)[c][r] = e_rc * g_r.
// Returns false on a non-positive pivot (reference: dsysv info != 0, qmf/Matrix.cpp:94).
__device__ __noinline__ bool factor_diag_tile(double a0, double a1, double* wtile, double* scratch, int lane) {
  const int r = lane >> 2, q = lane & 3;
  double e0 = (2 * q == r) ? 1.0 : 0.0, e1 = (2 * q + 1 == r) ? 1.0 : 0.0;
  double S = 1.0, prS = 1.0;
  bool ok = true;
  // deliberately NOT unrolled (instruction-cache footprint).  The pivot row is broadcast through 16
  // doubles of shared memory: 2 stores + 4 loads per step instead of 14 shuffles - the warp shares
  // the SM's memory-instruction queue with the trailing updates of the other warps
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    if (r == j) {
      *reinterpret_cast<double2*>(scratch + 2 * q) = make_double2(a0, a1);
      *reinterpret_cast<double2*>(scratch + 8 + 2 * q) = make_double2(e0, e1);
    }
    __syncwarp();
    const double p = scratch[j];
    const double2 uc = *reinterpret_cast<const double2*>(scratch + 2 * q);
    const double2 ec = *reinterpret_cast<const double2*>(scratch + 8 + 2 * q);
    const double ur = scratch[r];
    __syncwarp();
    ok = ok && (p > 0.0);
    const int hi = __double2hiint(p), lo = __double2loint(p);
    const double pn = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, lo);
    const double sc = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
    if (r == j) prS = p * S;
    S *= pn;
    const bool upd = r > j;
    // a*pn - (ur*sc)*u  ==  (a*p - ur*u) * sc bit for bit (sc is a power of two): the integer
    // normalisation of p is then off the dependent FP64 chain (product -> fma -> exact scale)
    const double n0 = fma(a0, p, -(ur * uc.x)) * sc, n1 = fma(a1, p, -(ur * uc.y)) * sc;
    const double m0 = fma(e0, p, -(ur * ec.x)) * sc, m1 = fma(e1, p, -(ur * ec.y)) * sc;
    a0 = upd ? n0 : a0;
    a1 = upd ? n1 : a1;
    e0 = upd ? m0 : e0;
    e1 = upd ? m1 : e1;
  }
  {  // last pivot a_77 (lane 31, second slot): nothing left to eliminate, only row 7's scale needs it
    const double p7 = __shfl_sync(0xffffffffu, a1, 31);
    ok = ok && (p7 > 0.0);
    if (r == 7) prS = p7 * S;
  }
  const double g = rsqrt(prS);
  // stored TRANSPOSED (wtile(r, c) = inv(U)[c][r]) so that this is one conflict-free 16-byte store
  *reinterpret_cast<double2*>(wtile + tile_acc_off(lane)) = make_double2(e0 * g, e1 * g);
  return ok;
}

// Same contract as factor_diag_tile (A = U^T U, W = inv(U)^T to wtile), but the pivot-to-pivot chain never
// leaves the register file: the TENSOR CORE does the data movement.  For an accumulator fragment X (lane
// 4n+q holds X[n][2q], X[n][2q+1]) slot h used as the A operand of an m8n8k4 DMMA is A[i][k] = X[i][2k+h] and
// used as the B operand it is B[k][m] = X[m][2k+h] - so for a SYMMETRIC trailing block the lanes with
// q == j/2 already hold column j of the tile (slot j%2) exactly where the DMMA wants both operands of the
// rank-1 update  C <- C pn - (u sc) u^T  (u = C[., j], rows > j; pn, sc as in factor_diag_tile).  No shared
// memory, no __syncwarp, and ONE shuffle per pivot (the pivot itself, needed by every lane for the exact
// power-of-two normalisation): the dependent chain per pivot is shuffle -> DMUL -> DMMA instead of
// store -> barrier -> 4 loads -> barrier -> DMUL -> DFMA -> DMUL.  The inverse rides along TRANSPOSED
// (T = E^T, T[n][r] <- T[n][r] s_r - T[n][j] (u_r sc)): its A operand is column j of T, its B operand the same
// u - also local.  Only the part of C below the diagonal (column j, rows > j) is ever read.
// Returns false on a non-positive pivot (reference: dsysv info != 0, qmf/Matrix.cpp:94).
__device__ __noinline__ bool factor_diag_tile_mma(double c0, double c1, double* wtile, int lane) {
  const int n = lane >> 2, q = lane & 3;
  double t0 = (2 * q == n) ? 1.0 : 0.0, t1 = (2 * q + 1 == n) ? 1.0 : 0.0;  // T = E^T = I
  double S = 1.0, prS0 = 1.0, prS1 = 1.0;  // prS_h = p_r S_r for r = 2q + h (columns of T this lane holds)
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const double cj = (j & 1) ? c1 : c0;                         // lanes q == j/2: C[n][j]
    const double p = __shfl_sync(0xffffffffu, cj, 4 * j + (j >> 1));  // C[j][j]
    ok = ok && (p > 0.0);
    const int hi = __double2hiint(p), lo = __double2loint(p);
    const double pn = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, lo);
    const double sc = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
    if ((j & 1) == 0) {
      if (2 * q == j) prS0 = p * S;
    } else {
      if (2 * q + 1 == j) prS1 = p * S;
    }
    S *= pn;
    const bool colj = q == (j >> 1);
    const double u = (colj && n > j) ? cj : 0.0;                 // u_n, rows below the pivot
    const double tj = colj ? ((j & 1) ? t1 : t0) : 0.0;          // T[n][j]
    const double us = u * sc;
    const double sr = n > j ? pn : 1.0;                          // live rows of C
    double c[2] = {c0 * sr, c1 * sr};
    double t[2] = {t0 * (2 * q > j ? pn : 1.0), t1 * (2 * q + 1 > j ? pn : 1.0)};  // live columns of T
    dmma(c, -us, u);
    dmma(t, -tj, us);
    c0 = c[0]; c1 = c[1];
    t0 = t[0]; t1 = t[1];
  }
  {  // last pivot C[7][7] (lane 31, second slot): nothing left to eliminate
    const double p7 = __shfl_sync(0xffffffffu, c1, 31);
    ok = ok && (p7 > 0.0);
    if (q == 3) prS1 = p7 * S;
  }
  // W[r][n] = T[n][r] rsqrt(p_r S_r): this lane holds rows r = 2q, 2q + 1 of W at column n
  wtile[(2 * q) * 8 + (n ^ tile_sw(2 * q))] = t0 * rsqrt(prS0);
  wtile[(2 * q + 1) * 8 + (n ^ tile_sw(2 * q + 1))] = t1 * rsqrt(prS1);
  return ok;
}

#ifndef QMFB_FACTOR_MMA
#define QMFB_FACTOR_MMA 1   // 1: factor_diag_tile_mma (tensor-core broadcast), 0: factor_diag_tile (shared-memory broadcast)
#endif
__device__ __forceinline__ bool factor_tile(double a0, double a1, double* wtile, double* scratch, int lane) {
#if QMFB_FACTOR_MMA
  return factor_diag_tile_mma(a0, a1, wtile, lane);
#else
  return factor_diag_tile(a0, a1, wtile, scratch, lane);
#endif
}

#ifdef QMFB_PROFILE_PHASES
// debug build only: thread 0 of every CTA accumulates clock64() deltas per phase into
// g_phase_cycles[phase] (build, tile store, factor (warp 0), panel, trailing/wait, back
// substitution, tail) and g_phase_cycles[15] counts rows
__device__ unsigned long long g_phase_cycles[24];
__device__ int g_debug_flags;  // bit 0: skip the trailing updates of the non-diagonal warps (timing experiments only)
#define QMFB_T(var) const long long var = clock64()
#define QMFB_ACC(idx, a, b) do { if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[idx], (unsigned long long)((b) - (a))); } while (0)
#define QMFB_ACC_IF(cond, idx, a, b) do { if (cond) atomicAdd(&g_phase_cycles[idx], (unsigned long long)((b) - (a))); } while (0)
#else
#define QMFB_T(var)
#define QMFB_ACC(idx, a, b)
#define QMFB_ACC_IF(cond, idx, a, b)
#endif

// Build phase of one row.  The kernel keeps its per-row schedule state in shared memory, so only
// the accumulators and a little loader state are live here.
// Gathers rows col[p0..p1) of Y through the cp.async ring, accumulates
//   A = G + sum_s (alpha r_s) y_s y_s^T   (upper tiles)      b = sum_s (1 + alpha r_s) y_s
// (WALSEngine.cpp:277-287), adds lambda to the diagonal (:290-292) and leaves the tiles, the
// per-warp partial sums of b and of sum_s (1 + alpha r_s) in shared memory.
//
// All non-DMMA work of a chunk is done by ONE warp, in rotation: warp (n mod NWARPS) gathers
// chunk n (32 x 16-byte cp.async per lane, weights, index prefetch) and accumulates b for it,
// the other warps do nothing but wait -> 4 k4-steps of DMMA -> release.  With every warp doing
// 1/NWARPS of the bookkeeping of every chunk (the first version) all warps left the DMMA pipe at
// the same time each chunk: measured 885 + 782 cycles of bookkeeping per chunk and warp around
// 1 088 cycles' worth of DMMA issue, tensor pipe 69 % busy (tools/exp_phases.py).
// WS (warp-specialised kernel): the tiles go to `tiles` (a buffer the solver group of row r - 2 may still be
// reading: wait for `tempty` first when `tempty_wait`), nothing aliases the ring and only the builder warps
// call this.  !WS: the tiles alias the ring, the whole CTA calls this.
template <class SM, bool WS>
__device__ __forceinline__ void build_row(unsigned char* smem, const double* __restrict__ Y, int64_t ldy,
                                          const int32_t* __restrict__ col, const double* __restrict__ val,
                                          const double* __restrict__ gram, double alpha, double lambda, int k,
                                          int64_t p0, int64_t p1, uint32_t base, const double* __restrict__ lsrc,
                                          double* tiles, double* bpart, uint64_t* tempty, uint32_t tempty_parity,
                                          bool tempty_wait, bool tma_gather) {
  // lsrc != nullptr: an extremely long row whose sum over its entries was built ahead by
  // long_row_partial_kernel (then p1 == p0 here: no gather loop, start from Gram + that sum)
  constexpr int NT = SM::kNT;
  constexpr int kStages = SM::kRing;  // (shadows the global ring depth: this layout's)
  double* stagebuf = reinterpret_cast<double*>(smem + SM::kOffStage);
  double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + kStages;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  constexpr int PPR = SM::KP / 2;             // 16-byte pieces per gathered row
  constexpr int NCOPY = kChunk * PPR / 32;    // cp.async per lane per chunk
  constexpr int kAhead = SM::kAhead;          // chunks in flight beyond the one being consumed
  QMFB_T(tq0);
  const int nch = int((p1 - p0 + kChunk - 1) / kChunk);
  const uint32_t lim = base + uint32_t(nch);  // the ring is reused as tile storage: no cross-row prefetch

  // loader state of THIS warp: the next chunk (absolute index, == warp mod NWARPS) it gathers
  uint32_t mine = base + (uint32_t(warp) + SM::NWARPS - base % SM::NWARPS) % SM::NWARPS;
  int32_t pcol = 0;   // lanes < kChunk: column of row `lane` of chunk `mine`
  double pval = 0.0;
  bool pvalid = false;
  double csum = 0.0;
  auto prefetch_idx = [&]() {
    const int64_t p = p0 + int64_t(mine - base) * kChunk + lane;
    pvalid = lane < kChunk && mine < lim && p < p1;
    pcol = pvalid ? __ldg(col + p) : 0;
    pval = pvalid ? __ldg(val + p) : 0.0;
  };
  auto issue_mine = [&]() {
    const uint32_t st = mine % kStages;
    if (mine >= kStages) mbar_wait(&empty[st], ((mine / kStages) & 1u) ^ 1u);
    if (lane < kChunk) {
      const double wb = pvalid ? 1.0 + alpha * pval : 0.0;  // WALSEngine.cpp:280  1 + alpha * r
      wts[st * 2 * kChunk + lane] = pvalid ? alpha * pval : 0.0;  // WALSEngine.cpp:282  alpha * r
      wts[st * 2 * kChunk + kChunk + lane] = wb;
      csum += wb;
      mbar_arrive(&full[st]);
    }
    if (tma_gather) {
      // same 32 + kChunk arrivals as the cp.async form; lane 0's carries the byte count of the chunk
      if (lane == 0) {
        mbar_arrive_expect_tx(&full[st], uint32_t(kChunk) * SM::KP * 8);
      } else {
        mbar_arrive(&full[st]);
      }
      __syncwarp();
      if (lane < kChunk) {
        bulk_g2s(stagebuf + (size_t(st) * kChunk + lane) * SM::LD, Y + int64_t(pcol) * ldy, SM::KP * 8, &full[st]);
      }
    } else {
#pragma unroll 8
      for (int m = 0; m < NCOPY; ++m) {
        const int q = lane + 32 * m, row = q / PPR, piece = q % PPR;
        const int32_t c = __shfl_sync(0xffffffffu, pcol, row);
        cp_async16(stagebuf + (size_t(st) * kChunk + row) * SM::LD + piece * 2, Y + int64_t(c) * ldy + piece * 2);
      }
      cp_async_arrive(&full[st]);
    }
    mine += SM::NWARPS;
    prefetch_idx();
  };
  prefetch_idx();
  while (mine < lim && mine < base + kAhead) issue_mine();  // prologue: the first kAhead chunks of the row (a warp may own several when kAhead > NWARPS)

  QMFB_T(tq1);
  double acc[NT + 3][2];
  acc[NT + 1][0] = acc[NT + 1][1] = acc[NT + 2][0] = acc[NT + 2][1] = 0.0;  // b column tiles of this warp's two tile rows
#pragma unroll
  for (int t = 0; t <= NT; ++t) {  // accumulators start from the Gram tiles
    int I, J;
    acc_tile<NT>(warp, t, I, J);
    const double2 g = *reinterpret_cast<const double2*>(gram + size_t(SM::gidx(I, J)) * 64 + lane * 2);
    acc[t][0] = g.x;
    acc[t][1] = g.y;
    if (lsrc != nullptr) {
      const double2 l = *reinterpret_cast<const double2*>(lsrc + size_t(SM::gidx(I, J)) * 64 + lane * 2);
      acc[t][0] += l.x;
      acc[t][1] += l.y;
    }
  }
  if (lsrc != nullptr && warp == 0 && lane == 0) csum += lsrc[SM::NTILE_A * 64 + SM::KP];
  QMFB_T(tq2);
#if QMFB_B_DFMA
  double bacc = (lsrc != nullptr && lane < 16) ? lsrc[SM::NTILE_A * 64 + warp * 16 + lane] : 0.0;
#endif
  for (int c = 0; c < nch; ++c) {
    const uint32_t gc = base + c, st = gc % kStages;
    // (classic layout: the stage refilled here was consumed TWO chunks ago, its empty barrier completed long ago)
    QMFB_T(tb0);
    if (mine == gc + kAhead && mine < lim) issue_mine();
    QMFB_T(tb1);
    QMFB_ACC(11, tb0, tb1);
    mbar_wait(&full[st], (gc / kStages) & 1u);
    QMFB_T(tb2);
    const double* sb = stagebuf + size_t(st) * kChunk * SM::LD;
    const double* w8 = wts + st * 2 * kChunk;
#if QMFB_B_DFMA
    {  // b: this warp's 16 columns, half of the chunk's rows per half-warp (conflict-free 128-byte segments)
      const double* pb = sb + (lane >> 4) * SM::LD + warp * 16 + (lane & 15);
      const double* pw = w8 + kChunk + (lane >> 4);
#pragma unroll
      for (int j = 0; j < kChunk / 2; ++j) bacc = fma(pw[2 * j], pb[2 * j * SM::LD], bacc);
    }
    chunk_mma_dispatch<NT, 0, false, SM>(warp, acc, sb, w8, lane);
#else
    chunk_mma_dispatch<NT, 0, true, SM>(warp, acc, sb, w8, lane);
#endif
    QMFB_T(tb3);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    QMFB_T(tb4);
    QMFB_ACC(12, tb1, tb2);
    QMFB_ACC(13, tb2, tb3);
    QMFB_ACC(14, tb3, tb4);
  }
  QMFB_T(tq3);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
  QMFB_T(tq4);
  if constexpr (WS) {
    if (tempty_wait) mbar_wait(tempty, tempty_parity);  // the solver group is done with this buffer (row r - 2)
  } else {
    __syncthreads();  // every warp is done reading the ring before the tiles overwrite it
  }
  QMFB_T(tq5);
  // slot = (chunk index INSIDE the row) mod NWARPS of the chunks this warp gathered, not the warp id: which
  // warp gathers which chunk depends on `base` (the CTA's history), the grouping of this sum must not -
  // the loss is then bit-identical whichever CTA / GPU / shard solves the row
  if (lane == 0) bpart[(uint32_t(warp) + SM::NWARPS - base % SM::NWARPS) % SM::NWARPS] = csum;
  // tiles to shared memory: A(i,i) += lambda (WALSEngine.cpp:290-292), unit pivot on padding
  const int r = lane >> 2, c0 = 2 * (lane & 3);
#pragma unroll
  for (int t = 0; t <= NT; ++t) {
    int I, J;
    acc_tile<NT>(warp, t, I, J);
    double v0 = acc[t][0], v1 = acc[t][1];
    if (I == J) {
      const int gi = 8 * I + r;
      if (c0 == r) v0 = gi < k ? v0 + lambda : 1.0;
      if (c0 + 1 == r) v1 = gi < k ? v1 + lambda : 1.0;
    }
    *reinterpret_cast<double2*>(tiles + size_t(SM::tidx(I, J)) * 64 + tile_acc_off(lane)) = make_double2(v0, v1);
  }
#if QMFB_B_DFMA
  bacc += __shfl_xor_sync(0xffffffffu, bacc, 16);  // lane l (and l + 16) now holds b(16 * warp + l % 16)
#pragma unroll
  for (int tt = 0; tt < 2; ++tt) {  // b column tiles 2 * warp + tt: column 0 = b, other columns 0
    const double v = __shfl_sync(0xffffffffu, bacc, 8 * tt + (lane >> 2));
    *reinterpret_cast<double2*>(tiles + size_t(SM::tidx(2 * warp + tt, NT)) * 64 + tile_acc_off(lane)) =
      make_double2((lane & 3) == 0 ? v : 0.0, 0.0);
  }
#else
  *reinterpret_cast<double2*>(tiles + size_t(SM::tidx(warp, NT)) * 64 + tile_acc_off(lane)) = make_double2(acc[NT + 1][0], acc[NT + 1][1]);
  *reinterpret_cast<double2*>(tiles + size_t(SM::tidx(NT - 1 - warp, NT)) * 64 + tile_acc_off(lane)) = make_double2(acc[NT + 2][0], acc[NT + 2][1]);
#endif
  QMFB_T(tq6);
  QMFB_ACC(16, tq0, tq1);
  QMFB_ACC(17, tq1, tq2);
  QMFB_ACC(18, tq2, tq3);
  QMFB_ACC(19, tq3, tq4);
  QMFB_ACC(20, tq4, tq5);
  QMFB_ACC(21, tq5, tq6);
}

// Factor + solve phase of one row, entirely in shared memory: blocked right-looking Cholesky of the
// upper tiles (panel width 8, b as tile column NT so that the forward substitution rides along),
// then the back substitution; leaves x in xvec, z in the b tiles, b in bcopy.  Deliberately its own
// (non-inlined) function: its register allocation is then independent of the build phase, whose
// 34+ accumulator registers otherwise push the operand fragments of the trailing update into local
// memory (3 STL.64 + 3 LDL.64 per four tiles in profiles/r01_solve_final_ncu.csv).
// Returns false on a non-positive pivot.

// Barrier of the warps that solve one row together: the whole CTA (NAMED == false, bar 0) or a group of
// NW warps inside a warp-specialised CTA (NAMED: named barrier `bar_id`, NW * 32 threads).
template <bool NAMED>
__device__ __forceinline__ void group_sync(int bar_id, int nthreads) {
  if constexpr (NAMED) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
  } else {
    __syncthreads();
  }
}

template <int NT>
__host__ __device__ constexpr int tile_index(int I, int J) { return I * (NT + 1) - I * (I - 1) / 2 + (J - I); }

template <int NT, int NW>
struct SolveDims {  // the names the body of solve_row_impl was written against
  static constexpr int KP = NT * 8;
  static constexpr int NWARPS = NW;
  static constexpr int NTILE = NT * (NT + 1) / 2 + NT;
  __host__ __device__ static constexpr int tidx(int I, int J) { return tile_index<NT>(I, J); }
};

// NW warps (warp = 0 .. NW-1, tid = 0 .. 32 NW - 1 inside the group; 32 NW >= 8 NT) solve the row whose
// tiles are at `tiles`; wt / bcopy / xvec / rvec / fscratch are the group's scratch areas.
//
// The Cholesky is DECOUPLED: warp 0 (the "chain" warp) owns the critical path - the 8 NT dependent pivots - and
// never waits at a block barrier.  Per step I it publishes W_I = inv(U_II)^T (mbarrier F), computes the one panel
// tile it needs itself, U(I,I+1) = W_I A(I,I+1), updates the next diagonal tile with it in registers and goes
// straight into its factorisation.  The other warps ("sweepers") follow one step behind: wait for W_I, panel of
// their columns, a barrier among themselves (mbarrier S, on which the chain warp only ARRIVES after storing its
// panel tile), then the trailing update - FIRST the two tiles the chain warp needs next, (I+1,I+2) and (I+2,I+2),
// signalled on mbarrier G, then the rest.  The chain warp therefore waits only for data (G of the previous sweep,
// complete long before it is needed); with block barriers it spent ~1.0k of every ~3.3k-cycle step waiting for the
// panel and the barriers (tools/exp_phases.py).  Every barrier completes exactly NT (even) times per row, so the
// phase parities are those of the step index.  mbarrier arrive/wait have release/acquire semantics at CTA scope:
// they order the shared- (or, k > 128, global-) memory tile traffic between the warps.
constexpr int kSolveBars = 3;  // F, S, G: 8 bytes each at the head of the group's `fscratch` (16 doubles)
template <int NW>
__device__ __forceinline__ void solve_bars_init(double* fscratch) {
  // one thread, before the group's first row (followed by a barrier of the group / the CTA)
  uint64_t* bar = reinterpret_cast<uint64_t*>(fscratch);
  constexpr int npri = NW - 1 >= 2 ? 2 : 1;
  mbar_init(&bar[0], 32);            // F: the chain warp's 32 lanes
  mbar_init(&bar[1], NW * 32);       // S: everyone arrives, the sweepers wait
  mbar_init(&bar[2], 32 * npri);     // G: the warps that update the two priority tiles
}

template <int NT, int NW, int TU, bool NAMED>
__device__ __noinline__ bool solve_row_impl_decoupled(double* tiles, double* wt, double* bcopy, double* xvec, double* rvec, double* fscratch,
                                            int warp, int lane, int tid, int bar_id) {
  using SM = SolveDims<NT, NW>;
  static_assert(NW * 32 >= NT * 8, "one thread per unknown in the back substitution");
  static_assert(NW >= 2 && NT % 2 == 0, "a chain warp plus at least one sweeper; even number of steps");
  const int fo = tile_frag_off(lane);                        // operand-fragment offset inside a tile
  const int fw = (lane >> 2) * 8 + ((lane & 3) ^ tile_sw(lane >> 2));  // same for the transposed W tiles; k+4 half at fw ^ 4
  const int co = tile_acc_off(lane);                         // accumulator-fragment offset
  uint64_t* barF = reinterpret_cast<uint64_t*>(fscratch);
  uint64_t* barS = barF + 1;
  uint64_t* barG = barF + 2;
  constexpr int nw = NW - 1;                                 // sweepers
  constexpr int npri = nw >= 2 ? 2 : 1;
  QMFB_T(tp1);
  // keep a copy of b (column 0 of the b tiles) for the loss before the factorisation overwrites it with z
  if (tid < SM::KP) bcopy[tid] = tiles[size_t(SM::tidx(tid >> 3, NT)) * 64 + (tid & 7) * 8 + tile_sw(tid & 7)];
  // every thread has its copy before any b tile is overwritten: the first panel tile written is (0, 1) by the chain
  // warp, b tiles (J == NT) only after the sweepers have seen W_0; NT >= 4, so one barrier of the group suffices
  group_sync<NAMED>(bar_id, NW * 32);
  bool ok = true;
  QMFB_T(tp2);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 1, tp1, tp2);
  if (warp == 0) {
    // ================= chain warp =================
    {
      const double2 a = *reinterpret_cast<const double2*>(tiles + size_t(SM::tidx(0, 0)) * 64 + co);
      ok = factor_tile(a.x, a.y, wt, fscratch + 8, lane);
    }
    for (int I = 0; I < NT; ++I) {
      QMFB_T(ts0);
      mbar_arrive(barF);                                     // F(I): this lane's part of W_I is published
      __syncwarp();                                          // ... and visible to the other lanes of this warp
      if (I > 0) mbar_wait(barG, uint32_t(I - 1) & 1u);      // (I, I+1) and (I+1, I+1) final through sweep I-1
      QMFB_T(ts1);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 4, ts0, ts1);
      double* t = tiles + size_t(SM::tidx(I, I + 1)) * 64;
      {
        const double w0 = wt[I * 64 + fw], w1 = wt[I * 64 + (fw ^ 4)];
        double c[2] = {0.0, 0.0};
        dmma(c, w0, t[fo]);
        dmma(c, w1, t[fo + 32]);
        __syncwarp();
        *reinterpret_cast<double2*>(t + co) = make_double2(c[0], c[1]);
      }
      mbar_arrive(barS);                                     // S(I): the panel tile (I, I+1) is published
      QMFB_T(ts2);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 3, ts1, ts2);
      if (I == NT - 1) break;
      __syncwarp();
      const double2 cv = *reinterpret_cast<const double2*>(tiles + size_t(SM::tidx(I + 1, I + 1)) * 64 + co);
      double c[2] = {cv.x, cv.y};
      const double u0 = t[fo], u1 = t[fo + 32];
      dmma(c, -u0, u0);
      dmma(c, -u1, u1);
      QMFB_T(tf0);  // the updated tile goes to the factor in registers (same fragment layout); U_II itself is never read again
      ok = factor_tile(c[0], c[1], wt + (I + 1) * 64, fscratch + 8, lane) && ok;
      QMFB_T(tf1);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 2, tf0, tf1);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 6, ts2, tf0);
      // every sweeper has observed F(I) (it arrived on S(I) afterwards): F may advance without a phase wrapping
      mbar_wait(barS, uint32_t(I) & 1u);
      QMFB_T(tf2);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 5, tf1, tf2);
    }
  } else {
    // ================= sweepers =================
    const int wslot = warp - 1;
    for (int I = 0; I < NT; ++I) {
      mbar_wait(barF, uint32_t(I) & 1u);                     // W_I
      // (b) panel: U[I][J] = inv(U_II)^T * A[I][J]  for J = I+2 .. NT  ((I, I+1) is the chain warp's)
      {
        const double w0 = wt[I * 64 + fw], w1 = wt[I * 64 + (fw ^ 4)];
        for (int J = I + 2 + wslot; J <= NT; J += nw) {
          double* t = tiles + size_t(SM::tidx(I, J)) * 64;
          double c[2] = {0.0, 0.0};
          dmma(c, w0, t[fo]);
          dmma(c, w1, t[fo + 32]);
          __syncwarp();
          *reinterpret_cast<double2*>(t + co) = make_double2(c[0], c[1]);
        }
      }
      mbar_arrive(barS);
      if (I == NT - 1) {
        if (wslot < npri) mbar_arrive(barG);                 // NT-th completion of G (nobody waits for it): parity stays even
        break;
      }
      mbar_wait(barS, uint32_t(I) & 1u);                     // the whole panel row I is in place
      // (c) trailing update: A[J1][J2] -= U[I][J1]^T U[I][J2], I < J1 <= J2 <= NT, J1 < NT, except (I+1, I+1) (chain warp)
      const int tstart = SM::tidx(I + 1, I + 1);
      const double* urow = tiles + size_t(SM::tidx(I, I)) * 64;  // tile (I, J) = urow + (J - I) * 64
      const int p1 = tstart + 1;                               // (I+1, I+2): exists (I <= NT - 2)
      const int p2 = tstart + (NT - I);                        // (I+2, I+2): exists iff I <= NT - 3
      const bool has_p2 = I <= NT - 3;
      auto update_one = [&](int ti, int j1, int j2) {
        const double* ta = urow + (j1 - I) * 64;
        const double* tb = urow + (j2 - I) * 64;
        double* tc = tiles + size_t(ti) * 64 + co;
        const double2 cv = *reinterpret_cast<const double2*>(tc);
        double c[2] = {cv.x, cv.y};
        const double a0 = -ta[fo], a1 = -ta[fo + 32], b0 = tb[fo], b1 = tb[fo + 32];
        dmma(c, a0, b0);
        dmma(c, a1, b1);
        *reinterpret_cast<double2*>(tc) = make_double2(c[0], c[1]);
      };
      if (wslot == 0) update_one(p1, I + 1, I + 2);
      if (wslot == npri - 1 && has_p2) update_one(p2, I + 2, I + 2);
      if (wslot < npri) {
        __syncwarp();
        mbar_arrive(barG);                                     // G(I)
      }
#ifdef QMFB_PROFILE_PHASES
      if ((g_debug_flags & 1) == 0)
#endif
      {
        // the remaining tiles: flat storage positions tstart + 2 .. NTILE - 1 except p2, dealt round-robin;
        // (J1, J2) decoded incrementally.  Straight-line body: out-of-range slots of the last pass recompute a
        // valid tile and skip the store.
        const int nrest = SM::NTILE - (tstart + 2) - (has_p2 ? 1 : 0);
        int J1 = I + 1, rowstart = tstart, rowlen = NT - I;    // row J1 holds tiles (J1, J1) .. (J1, NT)
        for (int d0 = wslot; d0 < nrest; d0 += TU * nw) {
          double c[TU][2], ua[TU][2], ub[TU][2];
          int tis[TU];
#pragma unroll
          for (int q = 0; q < TU; ++q) {
            const int d = d0 + q * nw;
            const bool v = d < nrest;
            int e = tstart + 2 + d;
            if (has_p2 && e >= p2) ++e;
            if (v) {
              while (e >= rowstart + rowlen) {
                rowstart += rowlen;
                --rowlen;
                ++J1;
              }
            }
            tis[q] = v ? e : -1;
            const double* ta = urow + (v ? J1 - I : 1) * 64;
            const double* tb = urow + (v ? J1 + (e - rowstart) - I : 1) * 64;
            const double2 cv = *reinterpret_cast<const double2*>(tiles + size_t(v ? e : tstart) * 64 + co);
            c[q][0] = cv.x; c[q][1] = cv.y;
            ua[q][0] = -ta[fo]; ua[q][1] = -ta[fo + 32];
            ub[q][0] = tb[fo]; ub[q][1] = tb[fo + 32];
          }
#pragma unroll
          for (int q = 0; q < TU; ++q) {
            dmma(c[q], ua[q][0], ub[q][0]);
            dmma(c[q], ua[q][1], ub[q][1]);
          }
#pragma unroll
          for (int q = 0; q < TU; ++q) {
            if (tis[q] >= 0) *reinterpret_cast<double2*>(tiles + size_t(tis[q]) * 64 + co) = make_double2(c[q][0], c[q][1]);
          }
        }
      }
    }
  }
  QMFB_T(tp4);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 7, tp2, tp4);

  // ---- back substitution U x = z: thread t < KP keeps r_t in a register; per block step one
  //      8x8 mat-vec by inv(U_JJ) and one rank-8 update of the rows above -----------------------
  //      ONE block barrier per step: the eight threads of block J live in one warp, so they exchange
  //      their finished r_J through shared memory under a __syncwarp and go straight on to x_J.
  group_sync<NAMED>(bar_id, NW * 32);
  double r = 0.0;
  if (tid < SM::KP) r = tiles[size_t(SM::tidx(tid >> 3, NT)) * 64 + (tid & 7) * 8 + tile_sw(tid & 7)];
  for (int J = NT - 1; J >= 0; --J) {
    const bool mine = (tid >> 3) == J;
    if (mine) rvec[tid & 7] = r;
    __syncwarp();
    if (mine) {  // x_J = W_J * r_J
      const double* w = wt + J * 64;  // transposed: inv(U_JJ)[row][c] = w(c, row)
      const int row = tid & 7;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        s0 += w[c * 8 + (row ^ tile_sw(c))] * rvec[c];
        s1 += w[(c + 1) * 8 + (row ^ tile_sw(c + 1))] * rvec[c + 1];
      }
      xvec[tid] = s0 + s1;
    }
    if (J == 0) break;
    group_sync<NAMED>(bar_id, NW * 32);  // x_J visible; also orders this step's rvec reads before the next step's writes
    if (tid < 8 * J) {  // r_t -= U[t][8J .. 8J+7] . x_J
      const double* u = tiles + size_t(SM::tidx(tid >> 3, J)) * 64 + (tid & 7) * 8;
      const double* x = xvec + 8 * J;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const int cr = (c + 2 * ((tid >> 1) & 3)) & 7;  // rotate the start pair by row/2: 8 rows -> 8 bank groups
        const double2 uv = *reinterpret_cast<const double2*>(u + (cr ^ tile_sw(tid & 7)));
        s0 += uv.x * x[cr];
        s1 += uv.y * x[cr + 1];
      }
      r -= s0 + s1;
    }
  }
  QMFB_T(tp5);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 8, tp4, tp5);
  return ok;
}


// SM is the shared-memory layout (WalsSmem<NT>: tiles in shared memory; WalsSmemBig<NT>, k > 128:
// GT = true and the tiles live in the CTA's L2-resident global workspace `gtiles`): the whole CTA solves.
// The same solve with BLOCK barriers: every warp of the group meets twice per panel step (W_I ready / panel row
// complete); warp 0 takes the next diagonal tile from its DMMA accumulators straight into the factor (look-ahead)
// while the others sweep the trailing tiles.  The measured round-1/2 default.
template <int NT, int NW, int TU, bool NAMED>
__device__ __noinline__ bool solve_row_impl_blocked(double* tiles, double* wt, double* bcopy, double* xvec, double* rvec, double* fscratch,
                                            int warp, int lane, int tid, int bar_id) {
  using SM = SolveDims<NT, NW>;
  static_assert(NW * 32 >= NT * 8, "one thread per unknown in the back substitution");
  const int fo = tile_frag_off(lane);                        // operand-fragment offset inside a tile
  const int fw = (lane >> 2) * 8 + ((lane & 3) ^ tile_sw(lane >> 2));  // same for the transposed W tiles; k+4 half at fw ^ 4
  const int co = tile_acc_off(lane);                         // accumulator-fragment offset
  QMFB_T(tp1);
  // keep a copy of b (column 0 of the b tiles) for the loss before the factorisation overwrites it with z
  if (tid < SM::KP) bcopy[tid] = tiles[size_t(SM::tidx(tid >> 3, NT)) * 64 + (tid & 7) * 8 + tile_sw(tid & 7)];
  // ---- blocked Cholesky, panel width 8; forward substitution rides along in column NT ------
  bool ok = true;
  QMFB_T(tp2);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 1, tp1, tp2);
  if (warp == 0) {
    const double2 a = *reinterpret_cast<const double2*>(tiles + size_t(SM::tidx(0, 0)) * 64 + co);
    ok = factor_tile(a.x, a.y, wt, fscratch + 8, lane);
  }
  QMFB_T(tp3);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 2, tp2, tp3);
  for (int I = 0; I < NT; ++I) {
    QMFB_T(ts0);
    group_sync<NAMED>(bar_id, NW * 32);  // W_I ready, row I of tiles final up to panel I-1
    QMFB_T(ts1);
    QMFB_ACC_IF(tid == 0 && bar_id <= 1, 4, ts0, ts1);
    // (b) panel: U[I][J] = inv(U_II)^T * A[I][J]  for J = I+1 .. NT
    {
      const double w0 = wt[I * 64 + fw], w1 = wt[I * 64 + (fw ^ 4)];
      for (int J = I + 1 + warp; J <= NT; J += SM::NWARPS) {
        double* t = tiles + size_t(SM::tidx(I, J)) * 64;
        double c[2] = {0.0, 0.0};
        dmma(c, w0, t[fo]);
        dmma(c, w1, t[fo + 32]);
        __syncwarp();
        *reinterpret_cast<double2*>(t + co) = make_double2(c[0], c[1]);
      }
    }
    QMFB_T(ts2);
    QMFB_ACC_IF(tid == 0 && bar_id <= 1, 3, ts1, ts2);
    if (I == NT - 1) break;
    group_sync<NAMED>(bar_id, NW * 32);
    QMFB_T(ts3);
    QMFB_ACC_IF(tid == 0 && bar_id <= 1, 5, ts2, ts3);
    // (c) trailing update: A[J1][J2] -= U[I][J1]^T U[I][J2], I < J1 <= J2 <= NT, J1 < NT.
    //     One warp updates the next diagonal tile first and factors it right away (look-ahead)
    //     while the other warps sweep the rest (kTU tiles in flight each).
    const int tstart = SM::tidx(I + 1, I + 1);
    const double* urow = tiles + size_t(SM::tidx(I, I)) * 64;  // tile (I, J) = urow + (J - I) * 64
    constexpr int dwarp = 0;
    const int nw = SM::NWARPS > 1 ? SM::NWARPS - 1 : 1;
    const int wslot = SM::NWARPS > 1 ? warp - 1 : 0;
    if (SM::NWARPS == 1 || warp == dwarp) {
      double* t = tiles + size_t(tstart) * 64;
      const double* u = urow + 64;
      double2 cv = *reinterpret_cast<double2*>(t + co);
      double c[2] = {cv.x, cv.y};
      const double u0 = u[fo], u1 = u[fo + 32];
      dmma(c, -u0, u0);
      dmma(c, -u1, u1);
      QMFB_T(tf0);  // the updated tile goes to the factor in registers (same fragment layout); U_II itself is never read again
      ok = factor_tile(c[0], c[1], wt + (I + 1) * 64, fscratch + 8, lane) && ok;
      QMFB_T(tf1);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 2, tf0, tf1);
      QMFB_ACC_IF(tid == 0 && bar_id <= 1, 6, ts3, tf0);
    }
#ifdef QMFB_PROFILE_PHASES
    if ((g_debug_flags & 1) == 0)
#endif
    if (SM::NWARPS == 1 || warp != dwarp) {
      // flat enumeration of the trailing tiles (contiguous in storage); (J1, J2) decoded incrementally
      int J1 = I + 1, off = 1 + wslot;  // position `off` inside row J1 (row J1 has NT - J1 + 1 tiles)
      for (int e = tstart + 1 + wslot; e < SM::NTILE; e += TU * nw) {
        // straight-line body: out-of-range slots of the last sweep recompute a valid tile and skip the store
        double c[TU][2], ua[TU][2], ub[TU][2];
#pragma unroll
        for (int q = 0; q < TU; ++q) {
          const int ti = e + q * nw;
          const bool v = ti < SM::NTILE;
          if (v) {
            while (off >= NT - J1 + 1) {
              off -= NT - J1 + 1;
              ++J1;
            }
          }
          const double* ta = urow + (v ? J1 - I : 1) * 64;
          const double* tb = urow + (v ? J1 + off - I : 1) * 64;
          const double2 cv = *reinterpret_cast<const double2*>(tiles + size_t(v ? ti : tstart) * 64 + co);
          c[q][0] = cv.x; c[q][1] = cv.y;
          ua[q][0] = -ta[fo]; ua[q][1] = -ta[fo + 32];
          ub[q][0] = tb[fo]; ub[q][1] = tb[fo + 32];
          off += nw;
        }
#pragma unroll
        for (int q = 0; q < TU; ++q) {
          dmma(c[q], ua[q][0], ub[q][0]);
          dmma(c[q], ua[q][1], ub[q][1]);
        }
#pragma unroll
        for (int q = 0; q < TU; ++q) {
          const int ti = e + q * nw;
          if (ti < SM::NTILE) *reinterpret_cast<double2*>(tiles + size_t(ti) * 64 + co) = make_double2(c[q][0], c[q][1]);
        }
      }
    }
  }
  QMFB_T(tp4);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 7, tp3, tp4);

  // ---- back substitution U x = z: thread t < KP keeps r_t in a register; per block step one
  //      8x8 mat-vec by inv(U_JJ) and one rank-8 update of the rows above -----------------------
  //      ONE block barrier per step: the eight threads of block J live in one warp, so they exchange
  //      their finished r_J through shared memory under a __syncwarp and go straight on to x_J.
  group_sync<NAMED>(bar_id, NW * 32);
  double r = 0.0;
  if (tid < SM::KP) r = tiles[size_t(SM::tidx(tid >> 3, NT)) * 64 + (tid & 7) * 8 + tile_sw(tid & 7)];
  for (int J = NT - 1; J >= 0; --J) {
    const bool mine = (tid >> 3) == J;
    if (mine) rvec[tid & 7] = r;
    __syncwarp();
    if (mine) {  // x_J = W_J * r_J
      const double* w = wt + J * 64;  // transposed: inv(U_JJ)[row][c] = w(c, row)
      const int row = tid & 7;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        s0 += w[c * 8 + (row ^ tile_sw(c))] * rvec[c];
        s1 += w[(c + 1) * 8 + (row ^ tile_sw(c + 1))] * rvec[c + 1];
      }
      xvec[tid] = s0 + s1;
    }
    if (J == 0) break;
    group_sync<NAMED>(bar_id, NW * 32);  // x_J visible; also orders this step's rvec reads before the next step's writes
    if (tid < 8 * J) {  // r_t -= U[t][8J .. 8J+7] . x_J
      const double* u = tiles + size_t(SM::tidx(tid >> 3, J)) * 64 + (tid & 7) * 8;
      const double* x = xvec + 8 * J;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const int cr = (c + 2 * ((tid >> 1) & 3)) & 7;  // rotate the start pair by row/2: 8 rows -> 8 bank groups
        const double2 uv = *reinterpret_cast<const double2*>(u + (cr ^ tile_sw(tid & 7)));
        s0 += uv.x * x[cr];
        s1 += uv.y * x[cr + 1];
      }
      r -= s0 + s1;
    }
  }
  QMFB_T(tp5);
  QMFB_ACC_IF(tid == 0 && bar_id <= 1, 8, tp4, tp5);
  return ok;
}


// SM is the shared-memory layout (WalsSmem<NT>: tiles in shared memory; WalsSmemBig<NT>, k > 128:
// GT = true and the tiles live in the CTA's L2-resident global workspace `gtiles`): the whole CTA solves.

#ifndef QMFB_DECOUPLED
#define QMFB_DECOUPLED 0   // 1: solve_row_impl_decoupled (chain warp + sweepers on mbarriers), 0: solve_row_impl_blocked
#endif
template <int NT, int NW, int TU, bool NAMED>
__device__ __forceinline__ bool solve_row_impl(double* tiles, double* wt, double* bcopy, double* xvec, double* rvec, double* fscratch,
                                               int warp, int lane, int tid, int bar_id) {
#if QMFB_DECOUPLED
  return solve_row_impl_decoupled<NT, NW, TU, NAMED>(tiles, wt, bcopy, xvec, rvec, fscratch, warp, lane, tid, bar_id);
#else
  return solve_row_impl_blocked<NT, NW, TU, NAMED>(tiles, wt, bcopy, xvec, rvec, fscratch, warp, lane, tid, bar_id);
#endif
}

template <class SM, bool GT>
__device__ __forceinline__ bool solve_row(unsigned char* smem, double* gtiles) {
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  return solve_row_impl<SM::kNT, SM::NWARPS, GT ? 4 : kTU, false>(
    GT ? gtiles : reinterpret_cast<double*>(smem + SM::kOffTiles), reinterpret_cast<double*>(smem + SM::kOffW),
    reinterpret_cast<double*>(smem + SM::kOffB), reinterpret_cast<double*>(smem + SM::kOffX),
    reinterpret_cast<double*>(smem + SM::kOffR), reinterpret_cast<double*>(smem + SM::kOffFs), warp, threadIdx.x & 31,
    threadIdx.x, 0);
}

template <int NT>
__global__ void __launch_bounds__(WalsSmem<NT>::NTHREADS, (NT >= 12 ? 2 : (NT >= 8 ? QMFB_OCC8 : QMFB_OCC4))) wals_solve_kernel(const SolveParams prm) {
  using SM = WalsSmem<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  double* stagebuf = reinterpret_cast<double*>(smem + SM::kOffStage);
  double* tiles = reinterpret_cast<double*>(smem + SM::kOffTiles);
  double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
  double* wt = reinterpret_cast<double*>(smem + SM::kOffW);
  double* bcopy = reinterpret_cast<double*>(smem + SM::kOffB);
  double* xvec = reinterpret_cast<double*>(smem + SM::kOffX);
  double* rvec = reinterpret_cast<double*>(smem + SM::kOffR);
  double* fscratch = reinterpret_cast<double*>(smem + SM::kOffFs);
  double* bhalf = reinterpret_cast<double*>(smem + SM::kOffBh);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + SM::kRing;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, tid = threadIdx.x;
  constexpr int TPR = SM::KP / 2;  // threads per gathered row (16 bytes each); 4 rows per pass

  if (tid == 0) {
    for (int s = 0; s < SM::kRing; ++s) {
      mbar_init(&full[s], 32 + kChunk);  // gathering warp: one deferred cp.async arrival per lane + weight writers
      mbar_init(&empty[s], SM::NWARPS);
    }
    solve_bars_init<SM::NWARPS>(fscratch);
    mbar_fence_init();
  }

  // ---- static serpentine schedule over the longest-first order; the per-row state lives in two
  //      shared-memory slots (current / next) so that it costs no registers during the build ------
  struct RowSlot {
    int64_t p0, p1;
    int32_t row;
    uint32_t base;  // absolute chunk index of the row's chunk 0 (ring stage = index % kStages)
    int64_t pad;
  };
  volatile RowSlot* slots = reinterpret_cast<volatile RowSlot*>(smem + SM::kOffRow);
  const int G = gridDim.x, bid = blockIdx.x;
  auto slot_of = [&](int i) { return i * G + ((i & 1) ? (G - 1 - bid) : bid); };
  const int nlong = prm.long_sum != nullptr ? __ldg(&prm.long_plan->nlong) : 0;  // rows order[0 .. nlong) are prebuilt
  if (tid == 0) {
    const int s0 = slot_of(0);
    const int r0 = s0 < prm.nrows ? prm.order[s0] : -1;
    slots[0].row = r0;
    slots[0].base = 0u;
    slots[0].p0 = r0 >= 0 ? prm.row_ptr[r0] : 0;
    slots[0].p1 = r0 >= 0 ? prm.row_ptr[r0 + 1] : 0;
  }
  int it = 0;
  __syncthreads();  // barriers initialised, first row slot visible

  for (;;) {
    const volatile RowSlot* cs = slots + (it & 1);
    if (cs->row < 0) break;
    // thread 0 walks the next row's index chain (order -> row_ptr) at leisure during this row
    int nrow = -1;
    if (tid == 0) {
      const int sn = slot_of(it + 1);
      if (sn < prm.nrows) nrow = __ldg(prm.order + sn);
    }
    QMFB_T(tp0);
    // an extremely long row at the head of `order` was summed ahead of this kernel: no gather loop
    const int opos = slot_of(it);
    const bool is_long = opos < nlong;
    const int64_t bp1 = is_long ? cs->p0 : cs->p1;
    build_row<SM, false>(smem, prm.Y, prm.ldy, prm.col, prm.val, prm.gram, prm.alpha, prm.lambda, prm.k, cs->p0, bp1, cs->base,
                         is_long ? prm.long_sum + size_t(opos) * LongRow<NT>::kLen : nullptr, tiles, bhalf, nullptr, 0u, false,
                         prm.tma_gather != 0);
    QMFB_T(tp1);
    QMFB_ACC(0, tp0, tp1);
    int64_t np0 = 0, np1 = 0;
    if (tid == 0 && nrow >= 0) {
      np0 = __ldg(prm.row_ptr + nrow);
      np1 = __ldg(prm.row_ptr + nrow + 1);
    }
    __syncthreads();
    if (!solve_row<SM, false>(smem, nullptr) && lane == 0) *prm.error = 1;
    __syncthreads();
    QMFB_T(tp5);
    // ---- loss term: c + x^T B x - 2 x^T b with x^T B x = z^T z - lambda x^T x (WALSEngine.cpp:295-304)
    if (warp == 0) {
      double part = 0.0;
      for (int i = lane; i < prm.k; i += 32) {
        const double z = tiles[size_t(SM::tidx(i >> 3, NT)) * 64 + (i & 7) * 8 + tile_sw(i & 7)];
        const double x = xvec[i];
        part += z * z - prm.lambda * x * x - 2.0 * x * bcopy[i];
      }
      if (lane < SM::NWARPS) part += bhalf[lane];  // sum_s (1 + alpha r_s), WALSEngine.cpp:286
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) prm.row_loss[cs->row] = part;
    }
    if (SM::NWARPS == 1) {
      store_solved_row(prm, xvec, prm.row_offset + cs->row, SM::KP, 0, lane, 0, 1);
    } else if (warp != 0) {
      store_solved_row(prm, xvec, prm.row_offset + cs->row, SM::KP, warp - 1, lane, 0, SM::NWARPS - 1);
    }
    // ---- next row -----------------------------------------------------------------------------------
    if (tid == 0) {
      volatile RowSlot* ns = slots + ((it + 1) & 1);
      ns->row = nrow;
      ns->p0 = np0;
      ns->p1 = np1;
      ns->base = cs->base + uint32_t((bp1 - cs->p0 + kChunk - 1) / kChunk);
    }
    ++it;
    __syncthreads();  // tiles / xvec / bcopy free again, next row slot visible
    QMFB_T(tp6);
    QMFB_ACC(9, tp5, tp6);
    QMFB_ACC(10, tp0, tp6);
#ifdef QMFB_PROFILE_PHASES
    if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[15], 1ull);
#endif
  }
  // peer replicas: make this thread's NVLink stores visible system-wide before it retires (kernel
  // completion implies it; stated explicitly because other ranks read the rows right after the next collective)
  if (prm.npeers > 0) __threadfence_system();
}

// ------------------------------------------------------------------------------------------
// warp-specialised variant (see WalsSmemWs): builders and two solver groups in one CTA per SM
// ------------------------------------------------------------------------------------------
#ifndef QMFB_WS_TU
#define QMFB_WS_TU 2   // trailing-update tiles in flight per solver warp (3 sweep warps instead of 7)
#endif

template <int NT>
__global__ void __launch_bounds__(WalsSmemWs<NT>::NTHREADS, 1) wals_solve_ws_kernel(const SolveParams prm) {
  using SM = WalsSmemWs<NT>;
  constexpr int NB = SM::NWARPS, NS = SM::NSOLVE;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* empty = full + SM::kRing;
  uint64_t* tfull = empty + SM::kRing;   // [2] tiles of buffer b are complete (builders -> solver group b)
  uint64_t* tempty = tfull + 2;          // [2] solver group b is done with buffer b
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < SM::kRing; ++s) {
      mbar_init(&full[s], 32 + kChunk);
      mbar_init(&empty[s], NB);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], NB * 32);
      mbar_init(&tempty[b], NS * 32);
      solve_bars_init<NS>(reinterpret_cast<double*>(smem + SM::kOffFs) + 32 * b);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int G = gridDim.x, bid = blockIdx.x;
  auto slot_of = [&](int i) { return i * G + ((i & 1) ? (G - 1 - bid) : bid); };
  const int nlong = prm.long_sum != nullptr ? __ldg(&prm.long_plan->nlong) : 0;  // rows order[0 .. nlong) are prebuilt

  if (warp < NB) {
    // ================= builders: rows it = 0, 1, 2, ... of this CTA's serpentine schedule =================
    struct RowSlot {
      int64_t p0, p1;
      int32_t row;
      uint32_t base;
      int64_t pad;
    };
    volatile RowSlot* slots = reinterpret_cast<volatile RowSlot*>(smem + SM::kOffRow);
    if (tid == 0) {
      const int s0 = slot_of(0);
      const int r0 = s0 < prm.nrows ? prm.order[s0] : -1;
      slots[0].row = r0;
      slots[0].base = 0u;
      slots[0].p0 = r0 >= 0 ? prm.row_ptr[r0] : 0;
      slots[0].p1 = r0 >= 0 ? prm.row_ptr[r0 + 1] : 0;
    }
    group_sync<true>(3, NB * 32);
    for (int it = 0;; ++it) {
      const volatile RowSlot* cs = slots + (it & 1);
      if (cs->row < 0) break;
      int nrow = -1;
      if (tid == 0) {
        const int sn = slot_of(it + 1);
        if (sn < prm.nrows) nrow = __ldg(prm.order + sn);
      }
      const int opos = slot_of(it);
      const bool is_long = opos < nlong;
      const int64_t bp1 = is_long ? cs->p0 : cs->p1;
      const int b = it & 1;
      QMFB_T(tb0);
      build_row<SM, true>(smem, prm.Y, prm.ldy, prm.col, prm.val, prm.gram, prm.alpha, prm.lambda, prm.k, cs->p0, bp1, cs->base,
                          is_long ? prm.long_sum + size_t(opos) * LongRow<NT>::kLen : nullptr,
                          reinterpret_cast<double*>(smem + SM::kOffTiles + size_t(b) * SM::kTileBytes),
                          reinterpret_cast<double*>(smem + SM::kOffBh) + 8 * b, &tempty[b], uint32_t(((it >> 1) - 1) & 1), it >= 2,
                          prm.tma_gather != 0);
      mbar_arrive(&tfull[b]);  // release: this thread's tile stores are visible to the solver group that acquires
      if (tid == 0) {
        volatile RowSlot* ns = slots + ((it + 1) & 1);
        ns->row = nrow;
        ns->p0 = nrow >= 0 ? __ldg(prm.row_ptr + nrow) : 0;
        ns->p1 = nrow >= 0 ? __ldg(prm.row_ptr + nrow + 1) : 0;
        ns->base = cs->base + uint32_t((bp1 - cs->p0 + kChunk - 1) / kChunk);
      }
      group_sync<true>(3, NB * 32);  // next row slot visible; every builder is past this row's ring
      QMFB_T(tb1);
      QMFB_ACC(10, tb0, tb1);
    }
  } else {
    // ================= solver group g: rows it = g, g + 2, ... out of tile buffer g =================
    const int g = (warp - NB) / NS, gw = (warp - NB) % NS, gtid = gw * 32 + lane;
    double* tiles = reinterpret_cast<double*>(smem + SM::kOffTiles + size_t(g) * SM::kTileBytes);
    double* wt = reinterpret_cast<double*>(smem + SM::kOffW) + size_t(g) * NT * 64;
    double* bcopy = reinterpret_cast<double*>(smem + SM::kOffB) + size_t(g) * SM::KP;
    double* xvec = reinterpret_cast<double*>(smem + SM::kOffX) + size_t(g) * SM::KP;
    double* rvec = reinterpret_cast<double*>(smem + SM::kOffR) + 8 * g;
    double* fscratch = reinterpret_cast<double*>(smem + SM::kOffFs) + 32 * g;
    const double* bhalf = reinterpret_cast<const double*>(smem + SM::kOffBh) + 8 * g;
    for (int it = g;; it += 2) {
      const int sl = slot_of(it);
      if (sl >= prm.nrows) break;
      const int row = __ldg(prm.order + sl);
      QMFB_T(tw0);
      mbar_wait(&tfull[g], uint32_t(it >> 1) & 1u);
      QMFB_T(tw1);
      QMFB_ACC_IF(g == 0 && gtid == 0, 22, tw0, tw1);
      if (!solve_row_impl<NT, NS, QMFB_WS_TU, true>(tiles, wt, bcopy, xvec, rvec, fscratch, gw, lane, gtid, 1 + g) && lane == 0) {
        *prm.error = 1;
      }
      group_sync<true>(1 + g, NS * 32);
      // loss term: c + x^T B x - 2 x^T b with x^T B x = z^T z - lambda x^T x (WALSEngine.cpp:295-304)
      if (gw == 0) {
        double part = 0.0;
        for (int i = lane; i < prm.k; i += 32) {
          const double z = tiles[size_t(SM::tidx(i >> 3, NT)) * 64 + (i & 7) * 8 + tile_sw(i & 7)];
          const double x = xvec[i];
          part += z * z - prm.lambda * x * x - 2.0 * x * bcopy[i];
        }
        if (lane < NB) part += bhalf[lane];  // sum_s (1 + alpha r_s), WALSEngine.cpp:286
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) prm.row_loss[row] = part;
      }
      if (NS == 1) {
        store_solved_row(prm, xvec, prm.row_offset + row, SM::KP, 0, lane, 0, 1);
      } else if (gw != 0) {
        store_solved_row(prm, xvec, prm.row_offset + row, SM::KP, gw - 1, lane, 0, NS - 1);
      }
      group_sync<true>(1 + g, NS * 32);  // every read of the buffer / scratch is done
      mbar_arrive(&tempty[g]);
      QMFB_T(tw2);
      QMFB_ACC_IF(g == 0 && gtid == 0, 23, tw1, tw2);
#ifdef QMFB_PROFILE_PHASES
      if (g == 0 && gtid == 0) atomicAdd(&g_phase_cycles[15], 1ull);
#endif
    }
  }
  if (prm.npeers > 0) __threadfence_system();
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
