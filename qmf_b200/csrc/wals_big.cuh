// WALS half-step kernels for 128 < k <= 256 (NT = KP/8 in {20, 24, 28, 32}).
//
// The upper-tile system of one row (NT(NT+1)/2 + NT tiles of 512 B: 280 KB at k = 256) no longer
// fits in shared memory or in the accumulator registers of one CTA, so each persistent CTA
// (one per SM, 16 warps) keeps it in its own slice of an L2-resident global workspace
// (148 x 280 KB = 41 MB of the 126 MB L2):
//   build   : super-chunks of kBigRows gathered rows are staged in shared memory (cp.async); every
//             warp sweeps its share of the tiles, read-modify-writing each accumulator tile once per
//             super-chunk (DMMA, b as the extra tile column exactly as in wals_kernels.cuh)
//   solve   : the SAME blocked Cholesky / back substitution code as k <= 128 (solve_row<SM, true>),
//             with the tile pointer aimed at the workspace
// Same reference semantics as wals_kernels.cuh (WALSEngine.cpp:246-310, Matrix.cpp:81-96).
// This is the functional path for large k; it is L2-latency bound, not tuned to the DMMA roofline.
#pragma once
#include "wals_kernels.cuh"

namespace qmfb {

constexpr int kBigRows = 32;  // gathered rows per super-chunk

template <int NT>
struct WalsSmemBig {
  static constexpr int kNT = NT;
  static constexpr int KP = NT * 8;
  static constexpr int LD = KP + 4;
  static constexpr int NWARPS = 16;
  static constexpr int NTHREADS = NWARPS * 32;
  static constexpr int NTILE_A = NT * (NT + 1) / 2;
  static constexpr int NTILE = NTILE_A + NT;
  static constexpr size_t kOffStage = 0;                                       // kBigRows x LD doubles
  static constexpr size_t kOffTiles = 0;                                       // unused (tiles are global)
  static constexpr size_t kOffWts = size_t(kBigRows) * LD * 8;                 // 2 x kBigRows doubles
  static constexpr size_t kOffCol = kOffWts + size_t(2) * kBigRows * 8;        // kBigRows int32
  static constexpr size_t kOffW = kOffCol + size_t(kBigRows) * 4;              // NT inverse diagonal tiles
  static constexpr size_t kOffB = kOffW + size_t(NT) * 64 * 8;
  static constexpr size_t kOffX = kOffB + size_t(KP) * 8;
  static constexpr size_t kOffR = kOffX + size_t(KP) * 8;
  static constexpr size_t kOffFs = kOffR + 64;
  static constexpr size_t kOffBh = kOffFs + 256;                               // sum of (1 + alpha r)
  static constexpr size_t kBytes = kOffBh + 64;
  __host__ __device__ static constexpr int tidx(int I, int J) { return I * (NT + 1) - I * (I - 1) / 2 + (J - I); }
  __host__ __device__ static constexpr int gidx(int I, int J) { return I * NT - I * (I - 1) / 2 + (J - I); }
};

__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// One staged super-chunk: tile (I, J) += sum_s (wa_s y_s(8I+m)) y_s(8J+n)   (J < NT)
//                         tile (I, NT)[m][0] += sum_s y_s(8I+m) wb_s        (b column)
// `rowlen(I)` tiles per tile row, flat index t; c starts from `init` on the first super-chunk.
template <class SM, bool WITH_B>
__device__ __forceinline__ void big_accumulate(const double* sb, const double* wts, double* tiles, const double* gram,
                                               bool first, bool last, double lambda, int k, int warp, int lane) {
  constexpr int NT = SM::kNT;
  constexpr int ROWLEN0 = WITH_B ? NT + 1 : NT;
  const int ntile = WITH_B ? SM::NTILE : SM::NTILE_A;
  const int m = lane >> 2, kk = lane & 3;
  int I = 0, rem = warp;
  for (int t = warp; t < ntile; t += SM::NWARPS) {
    while (rem >= ROWLEN0 - I) {
      rem -= ROWLEN0 - I;
      ++I;
    }
    const int J = I + rem;
    // solve path: swizzled accumulator layout (what solve_row reads); Gram path: packed row-major
    double* tp = tiles + size_t(t) * 64 + (WITH_B ? tile_acc_off(lane) : lane * 2);
    double c[2] = {0.0, 0.0};
    if (!first) {
      const double2 v = *reinterpret_cast<const double2*>(tp);
      c[0] = v.x;
      c[1] = v.y;
    } else if (WITH_B && J < NT) {
      const double2 v = *reinterpret_cast<const double2*>(gram + size_t(SM::gidx(I, J)) * 64 + lane * 2);
      c[0] = v.x;
      c[1] = v.y;
    }
    const double* pa = sb + kk * SM::LD + 8 * I + m;
    if (!WITH_B || J < NT) {
      const double* pb = sb + kk * SM::LD + 8 * J + m;
#pragma unroll
      for (int s0 = 0; s0 < kBigRows; s0 += 4) dmma(c, pa[s0 * SM::LD] * wts[s0 + kk], pb[s0 * SM::LD]);
    } else {
#pragma unroll
      for (int s0 = 0; s0 < kBigRows; s0 += 4) dmma(c, pa[s0 * SM::LD], m == 0 ? wts[kBigRows + s0 + kk] : 0.0);
    }
    if (WITH_B && last && I == J) {  // A(i,i) += lambda (WALSEngine.cpp:290-292), unit pivot on padding
      const int gi = 8 * I + m, c0 = 2 * kk;
      if (c0 == m) c[0] = gi < k ? c[0] + lambda : 1.0;
      if (c0 + 1 == m) c[1] = gi < k ? c[1] + lambda : 1.0;
    }
    *reinterpret_cast<double2*>(tp) = make_double2(c[0], c[1]);
    rem += SM::NWARPS;
  }
}

// partial Gram of rows [r0, r1) per CTA into partial[blockIdx.x] (packed upper tiles)
template <int NT>
__global__ void __launch_bounds__(WalsSmemBig<NT>::NTHREADS, 1)
    gram_partial_big_kernel(const double* __restrict__ Y, int64_t ldy, int64_t row_begin, int64_t row_end,
                            double* __restrict__ partial, int part0, int nparts) {
  using SM = WalsSmemBig<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  double* sb = reinterpret_cast<double*>(smem + SM::kOffStage);
  double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  constexpr int PPR = SM::KP / 2;
  const int64_t n = row_end - row_begin;
  const int part = part0 + int(blockIdx.x);  // see gram_partial_kernel
  const int64_t r0 = row_begin + n * part / nparts;
  const int64_t r1 = row_begin + n * (part + 1) / nparts;
  double* out = partial + size_t(part) * SM::NTILE_A * 64;
  if (r1 <= r0) {
    for (int i = tid; i < SM::NTILE_A * 64; i += SM::NTHREADS) out[i] = 0.0;
    return;
  }
  const int nch = int((r1 - r0 + kBigRows - 1) / kBigRows);
  for (int c = 0; c < nch; ++c) {
    __syncthreads();
    const int64_t base = r0 + int64_t(c) * kBigRows;
    if (tid < kBigRows) wts[tid] = base + tid < r1 ? 1.0 : 0.0;
    for (int q = tid; q < kBigRows * PPR; q += SM::NTHREADS) {
      const int row = q / PPR, piece = q % PPR;
      const int64_t p = base + row < r1 ? base + row : r0;
      cp_async16(sb + row * SM::LD + piece * 2, Y + p * ldy + piece * 2);
    }
    cp_async_wait_all();
    __syncthreads();
    big_accumulate<SM, false>(sb, wts, out, nullptr, c == 0, false, 0.0, 0, warp, lane);
  }
}

template <int NT>
__global__ void __launch_bounds__(WalsSmemBig<NT>::NTHREADS, 1) wals_solve_big_kernel(const SolveParams prm, double* __restrict__ ws) {
  using SM = WalsSmemBig<NT>;
  extern __shared__ __align__(128) unsigned char smem[];
  double* sb = reinterpret_cast<double*>(smem + SM::kOffStage);
  double* wts = reinterpret_cast<double*>(smem + SM::kOffWts);
  int32_t* cols = reinterpret_cast<int32_t*>(smem + SM::kOffCol);
  double* bcopy = reinterpret_cast<double*>(smem + SM::kOffB);
  double* xvec = reinterpret_cast<double*>(smem + SM::kOffX);
  double* csum_s = reinterpret_cast<double*>(smem + SM::kOffBh);
  double* tiles = ws + size_t(blockIdx.x) * SM::NTILE * 64;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  constexpr int PPR = SM::KP / 2;
  if (tid == 0) {
    solve_bars_init<SM::NWARPS>(reinterpret_cast<double*>(smem + SM::kOffFs));
    mbar_fence_init();
  }
  __syncthreads();

  for (int slot = blockIdx.x; slot < prm.nrows; slot += gridDim.x) {
    const int row = prm.order[slot];
    const int64_t p0 = prm.row_ptr[row], p1 = prm.row_ptr[row + 1];
    const int nch = p1 > p0 ? int((p1 - p0 + kBigRows - 1) / kBigRows) : 1;  // an empty row still needs A = G + lambda I
    double csum = 0.0;
    for (int c = 0; c < nch; ++c) {
      __syncthreads();  // previous super-chunk consumed (and previous row finished)
      if (tid < kBigRows) {
        const int64_t p = p0 + int64_t(c) * kBigRows + tid;
        const bool valid = p < p1;
        const double v = valid ? prm.val[p] : 0.0;
        cols[tid] = valid ? prm.col[p] : 0;
        const double wb = valid ? 1.0 + prm.alpha * v : 0.0;  // WALSEngine.cpp:280
        wts[tid] = valid ? prm.alpha * v : 0.0;               // WALSEngine.cpp:282
        wts[kBigRows + tid] = wb;
        csum += wb;
      }
      __syncthreads();
      for (int q = tid; q < kBigRows * PPR; q += SM::NTHREADS) {
        const int r = q / PPR, piece = q % PPR;
        cp_async16(sb + r * SM::LD + piece * 2, prm.Y + int64_t(cols[r]) * prm.ldy + piece * 2);
      }
      cp_async_wait_all();
      __syncthreads();
      big_accumulate<SM, true>(sb, wts, tiles, prm.gram, c == 0, c == nch - 1, prm.lambda, prm.k, warp, lane);
    }
    if (warp == 0) {  // kBigRows == 32: the weight writers are exactly warp 0
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
      if (lane == 0) csum_s[0] = csum;
    }
    __syncthreads();
    if (!solve_row<SM, true>(smem, tiles) && lane == 0) *prm.error = 1;
    __syncthreads();
    // loss term: c + x^T B x - 2 x^T b with x^T B x = z^T z - lambda x^T x (WALSEngine.cpp:295-304)
    if (warp == 0) {
      double part = 0.0;
      for (int i = lane; i < prm.k; i += 32) {
        const double z = tiles[size_t(SM::tidx(i >> 3, NT)) * 64 + (i & 7) * 8 + tile_sw(i & 7)];
        const double x = xvec[i];
        part += z * z - prm.lambda * x * x - 2.0 * x * bcopy[i];
      }
      if (lane == 0) part += csum_s[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) prm.row_loss[row] = part;
    } else {
      store_solved_row(prm, xvec, prm.row_offset + row, SM::KP, warp - 1, lane, 0, SM::NWARPS - 1);
    }
  }
  if (prm.npeers > 0) __threadfence_system();
}

}  // namespace qmfb
