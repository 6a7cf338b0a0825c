// C-ABI (include/qmf_b200.h) for the WALS half-step: launchers + the host-buffer engine handle.
#include "qmfb_common.h"
#include "wals_kernels.cuh"
#include "wals_big.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <numeric>
#include <vector>

namespace qmfb {

constexpr int kGramMaxParts = 296;  // 2 CTAs per SM on a 148-SM B200

// Per-DEVICE launch configuration.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the
// current device/context only and the occupancy-derived grid cap depends on the device, so the state
// is kept per device ordinal (a process may drive several GPUs: qmfb_wals_sharded_*, --ngpus) and is
// guarded for concurrent host threads.
constexpr int kMaxDevices = 64;
struct PerDeviceCfg {
  std::mutex mu;
  bool done[kMaxDevices] = {};
  int value[kMaxDevices] = {};
};
// runs init(dev, value) once per device; returns the cached value through *out
template <class Init>
static int per_device(PerDeviceCfg& cfg, int* out, Init&& init) {
  int dev = 0;
  QMFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return set_error(QMFB_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(cfg.mu);
  if (!cfg.done[dev]) {
    int v = 0;
    if (int rc = init(dev, &v)) return rc;
    cfg.value[dev] = v;
    cfg.done[dev] = true;
  }
  if (out) *out = cfg.value[dev];
  return QMFB_OK;
}

// Number of fixed row parts the Gram of an n-row matrix is summed over (one CTA per part)
static int gram_nparts(int64_t n, int kp) {
  if (kp <= 128) return int(std::min<int64_t>(kGramMaxParts, std::max<int64_t>(1, (n + 4 * kChunk - 1) / (4 * kChunk))));
  return int(std::min<int64_t>(kGramMaxParts / 2, std::max<int64_t>(1, (n + 4 * kBigRows - 1) / (4 * kBigRows))));
}

template <int NT>
static int launch_gram_parts(cudaStream_t st, const double* Y, int64_t ldy, int64_t r0, int64_t r1, int p0, int p1, double* ws) {
  using SM = WalsSmem<NT>;
  static PerDeviceCfg cfg;
  if (int rc = per_device(cfg, nullptr, [](int, int*) -> int {
        QMFB_CUDA(cudaFuncSetAttribute(gram_partial_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::kBytes)));
        return QMFB_OK;
      })) return rc;
  if (p1 <= p0) return QMFB_OK;
  gram_partial_kernel<NT><<<p1 - p0, SM::NTHREADS, SM::kBytes, st>>>(Y, ldy, r0, r1, ws, p0, gram_nparts(r1 - r0, SM::KP));
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

// Stream-ordered scratch (cudaMallocAsync) comes from the device's default memory pool, which by
// default gives freed memory back to the driver at the next synchronisation: every call would then pay
// a real allocation.  Keep the pool's memory cached (once per device).
static int keep_pool_cached() {
  static PerDeviceCfg cfg;
  return per_device(cfg, nullptr, [](int dev, int*) -> int {
    cudaMemPool_t pool;
    QMFB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = UINT64_MAX;
    QMFB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    return QMFB_OK;
  });
}

// 0 = automatic (by mean row length), 1 = always the plain kernel, 2 = always the warp-specialised kernel (k <= 128)
static std::atomic<int> g_solve_kernel{[] {
  const char* e = getenv("QMFB_SOLVE");
  return e == nullptr ? 0 : (e[0] == 'c' ? 1 : (e[0] == 'w' ? 2 : 0));
}()};

// rows of the matrix the next solve gathers from (set by the engines right before the launch; 0 = unknown)
static std::atomic<int64_t> g_gather_rows_hint{0};


// nnz_hint: number of signals of the rows being solved (-1: unknown)
template <int NT>
static int launch_solve(cudaStream_t st, const SolveParams& prm, double* loss_sum, int32_t* scratch, int64_t nnz_hint) {
  if (int rc = keep_pool_cached()) return rc;
  using SM = WalsSmem<NT>;
  static PerDeviceCfg cfg;
  int grid_cap = 0;
  if (int rc = per_device(cfg, &grid_cap, [](int dev, int* cap) -> int {
        QMFB_CUDA(cudaFuncSetAttribute(wals_solve_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::kBytes)));
        QMFB_CUDA(cudaFuncSetAttribute(long_row_partial_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::kBytes)));
        int sms = 0, occ = 0;
        QMFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        QMFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wals_solve_kernel<NT>, SM::NTHREADS, SM::kBytes));
        if (occ < 1) return set_error(QMFB_ERR_CUDA, "wals_solve_kernel does not fit on an SM");
        *cap = sms * occ;
        return QMFB_OK;
      })) return rc;
  QMFB_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(int32_t), st));
  if (prm.nrows > 0) {
    static const int grid_env = getenv("QMFB_GRID_CAP") ? atoi(getenv("QMFB_GRID_CAP")) : 0;  // measurement switch
    const int grid = int(std::min<int64_t>(grid_env > 0 ? std::min(grid_env, grid_cap) : grid_cap, prm.nrows));
    // extremely long rows at the head of `order` are summed by many CTAs ahead of the solve kernel
    // (three near-empty launches when there is none); stream-ordered scratch, freed after the kernel
    constexpr int kLen = LongRow<NT>::kLen;
    double* long_buf = nullptr;
    const size_t plan_doubles = (sizeof(LongPlan) + 7) / 8;
    QMFB_CUDA(cudaMallocAsync(&long_buf, (size_t(kLongMaxSegs + kLongMaxRows) * kLen + plan_doubles) * sizeof(double), st));
    LongRowParams lp{prm.Y, prm.ldy, prm.row_ptr, prm.col, prm.val, prm.order, prm.nrows, prm.alpha, long_buf,
                     long_buf + size_t(kLongMaxSegs) * kLen,
                     reinterpret_cast<LongPlan*>(long_buf + size_t(kLongMaxSegs + kLongMaxRows) * kLen)};
    long_row_plan_kernel<<<1, kLongMaxRows, 0, st>>>(lp);
    long_row_partial_kernel<NT><<<grid_cap, SM::NTHREADS, SM::kBytes, st>>>(lp);
    long_row_reduce_kernel<NT><<<dim3((kLen + 255) / 256, 32), 256, 0, st>>>(lp);
    SolveParams run = prm;
    static const bool no_long = getenv("QMFB_NO_LONG_ROWS") != nullptr;  // measurement switch: ignore the sums
    run.long_sum = no_long ? nullptr : lp.sum;
    run.long_plan = lp.plan;
    // Two solve kernels: the plain one (two CTAs per SM, each alternating build and solve) and the warp-specialised
    // one (one CTA per SM: builder warps + two solver groups, three rows in flight).  Measured on the user-shaped
    // half of C4 (gpurun r02_variants.log): 24.75 ms plain vs 24.95 - 27.3 ms warp-specialised - its 4-warp solver
    // groups are slower per row than the 8-warp solve and the build no longer hides them - so the plain kernel is the
    // default for every row length; qmfb_wals_set_solve_kernel(2) / QMFB_SOLVE=ws selects the other (tests, measurements).
    (void)nnz_hint;
    bool ws = false;
    const int mode = g_solve_kernel.load();
    if (mode == 1) ws = false;
    if (mode == 2) ws = true;
    if constexpr (NT < 8) ws = false;  // a solver group needs a chain warp and at least one sweeper (NT / 4 >= 2 warps)
    if constexpr (NT >= 8) if (ws) {
      using WS = WalsSmemWs<NT>;
      static PerDeviceCfg wcfg;
      int wcap = 0;
      if (int rc = per_device(wcfg, &wcap, [](int dev, int* cap) -> int {
            QMFB_CUDA(cudaFuncSetAttribute(wals_solve_ws_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WS::kBytes)));
            int sms = 0, occ = 0;
            QMFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            QMFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wals_solve_ws_kernel<NT>, WS::NTHREADS, WS::kBytes));
            if (occ < 1) return set_error(QMFB_ERR_CUDA, "wals_solve_ws_kernel does not fit on an SM");
            *cap = sms * occ;
            return QMFB_OK;
          })) {
        cudaFreeAsync(long_buf, st);
        return rc;
      }
      wals_solve_ws_kernel<NT><<<int(std::min<int64_t>(wcap, prm.nrows)), WS::NTHREADS, WS::kBytes, st>>>(run);
    }
    if (!ws) wals_solve_kernel<NT><<<grid, SM::NTHREADS, SM::kBytes, st>>>(run);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(long_buf, st);
    QMFB_CUDA(e);
  }
  sum_kernel<<<1, 1024, 0, st>>>(prm.row_loss, prm.nrows, loss_sum);
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

// ---- 128 < k <= 256: tiles in an L2-resident global workspace (wals_big.cuh) ----------------
template <int NT>
static int launch_gram_parts_big(cudaStream_t st, const double* Y, int64_t ldy, int64_t r0, int64_t r1, int p0, int p1, double* ws) {
  using SM = WalsSmemBig<NT>;
  static PerDeviceCfg cfg;
  if (int rc = per_device(cfg, nullptr, [](int, int*) -> int {
        QMFB_CUDA(cudaFuncSetAttribute(gram_partial_big_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::kBytes)));
        return QMFB_OK;
      })) return rc;
  if (p1 <= p0) return QMFB_OK;
  gram_partial_big_kernel<NT><<<p1 - p0, SM::NTHREADS, SM::kBytes, st>>>(Y, ldy, r0, r1, ws, p0, gram_nparts(r1 - r0, SM::KP));
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

template <int NT>
static int launch_solve_big(cudaStream_t st, const SolveParams& prm, double* loss_sum, int32_t* scratch) {
  using SM = WalsSmemBig<NT>;
  static PerDeviceCfg cfg;
  int sms = 0;
  if (int rc = per_device(cfg, &sms, [](int dev, int* out) -> int {
        QMFB_CUDA(cudaFuncSetAttribute(wals_solve_big_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SM::kBytes)));
        QMFB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
        return QMFB_OK;
      })) return rc;
  if (int rc = keep_pool_cached()) return rc;
  QMFB_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(int32_t), st));
  if (prm.nrows > 0) {
    const int grid = int(std::min<int64_t>(sms, prm.nrows));
    double* ws = nullptr;  // stream-ordered: concurrent solves on other streams get their own slice
    QMFB_CUDA(cudaMallocAsync(&ws, size_t(grid) * SM::NTILE * 64 * sizeof(double), st));
    wals_solve_big_kernel<NT><<<grid, SM::NTHREADS, SM::kBytes, st>>>(prm, ws);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    QMFB_CUDA(e);
  }
  sum_kernel<<<1, 1024, 0, st>>>(prm.row_loss, prm.nrows, loss_sum);
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

int det_sum_launch(cudaStream_t st, const double* v, int64_t n, double* out) {
  sum_kernel<<<1, 1024, 0, st>>>(v, n, out);
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

}  // namespace qmfb

using namespace qmfb;

extern "C" {

#ifdef QMFB_PROFILE_PHASES
// debug build only (not part of include/qmf_b200.h): read and reset the phase cycle counters
int qmfb_debug_phase_cycles(unsigned long long* out24) {
  QMFB_CUDA(cudaMemcpyFromSymbol(out24, g_phase_cycles, sizeof(unsigned long long) * 24));
  unsigned long long zero[24] = {0};
  QMFB_CUDA(cudaMemcpyToSymbol(g_phase_cycles, zero, sizeof(zero)));
  return QMFB_OK;
}
#endif

#ifdef QMFB_PROFILE_PHASES
int qmfb_debug_set_flags(int flags) {
  QMFB_CUDA(cudaMemcpyToSymbol(g_debug_flags, &flags, sizeof(int)));
  return QMFB_OK;
}
#endif

int qmfb_wals_set_solve_kernel(int mode) {
  if (mode < 0 || mode > 2) return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_solve_kernel: mode must be 0 (auto), 1 (plain) or 2 (warp-specialised)");
  g_solve_kernel.store(mode);
  return QMFB_OK;
}

int qmfb_padded_k(int k) {
  if (k < 1 || k > 256) return set_error(QMFB_ERR_UNSUPPORTED, "nfactors must be in [1, 256] (got %d)", k);
  return ((k + 31) / 32) * 32;
}

int64_t qmfb_gram_packed_len(int k) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  const int nt = kp / 8;
  return int64_t(nt) * (nt + 1) / 2 * 64;
}

int64_t qmfb_gram_workspace_len(int k) {
  const int64_t n = qmfb_gram_packed_len(k);
  return n < 0 ? n : n * kGramMaxParts;
}

int qmfb_gram_parts_count(int64_t nrows, int k) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  if (nrows < 0) return set_error(QMFB_ERR_INVALID, "qmfb_gram_parts_count: bad argument");
  return gram_nparts(nrows, kp);
}

int qmfb_gram_parts_dev(void* stream, const double* Y, int64_t ldy, int64_t row_begin, int64_t row_end, int k, int part_begin,
                        int part_end, double* workspace) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  if (ldy < kp || (ldy & 1) || (reinterpret_cast<uintptr_t>(Y) & 15) || row_end < row_begin || !Y || !workspace || part_begin < 0 ||
      part_end < part_begin || part_end > gram_nparts(row_end - row_begin, kp)) {
    return set_error(QMFB_ERR_INVALID, "qmfb_gram_parts_dev: bad argument (Y must be 16-byte aligned, ldy even and >= padded k)");
  }
  auto st = static_cast<cudaStream_t>(stream);
  switch (kp / 8) {
    case 4: return launch_gram_parts<4>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 8: return launch_gram_parts<8>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 12: return launch_gram_parts<12>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 16: return launch_gram_parts<16>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 20: return launch_gram_parts_big<20>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 24: return launch_gram_parts_big<24>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 28: return launch_gram_parts_big<28>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
    case 32: return launch_gram_parts_big<32>(st, Y, ldy, row_begin, row_end, part_begin, part_end, workspace);
  }
  return set_error(QMFB_ERR_UNSUPPORTED, "unsupported padded k %d", kp);
}

int qmfb_gram_reduce_parts_dev(void* stream, const double* const* workspaces, const int* part_end, int nsrc, int k,
                               double* gram_packed) {
  const int64_t nelem = qmfb_gram_packed_len(k);
  if (nelem < 0) return int(nelem);
  if (!workspaces || !part_end || nsrc < 1 || nsrc > kMaxGramSrc || !gram_packed) {
    return set_error(QMFB_ERR_INVALID, "qmfb_gram_reduce_parts_dev: bad argument");
  }
  GramReduceParams prm{};
  for (int d = 0; d < nsrc; ++d) {
    if (!workspaces[d] || part_end[d] < (d ? part_end[d - 1] : 0)) return set_error(QMFB_ERR_INVALID, "qmfb_gram_reduce_parts_dev: bad source %d", d);
    prm.src[d] = workspaces[d];
    prm.part_end[d] = part_end[d];
  }
  prm.nsrc = nsrc;
  prm.nelem = int(nelem);
  gram_reduce_kernel<<<(prm.nelem + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(prm, gram_packed);
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

int qmfb_gram_dev(void* stream, const double* Y, int64_t ldy, int64_t row_begin, int64_t row_end, int k,
                  double* workspace, double* gram_packed) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  if (row_end < row_begin || !gram_packed) return set_error(QMFB_ERR_INVALID, "qmfb_gram_dev: bad argument");
  const int parts = gram_nparts(row_end - row_begin, kp);
  if (int rc = qmfb_gram_parts_dev(stream, Y, ldy, row_begin, row_end, k, 0, parts, workspace)) return rc;
  const double* src[1] = {workspace};
  return qmfb_gram_reduce_parts_dev(stream, src, &parts, 1, k, gram_packed);
}

int qmfb_gram_unpack_dev(void* stream, const double* gram_packed, int k, double* out) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  auto st = static_cast<cudaStream_t>(stream);
  const int nb = (k * k + 255) / 256;
  switch (kp / 8) {
    case 4: gram_unpack_kernel<4><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 8: gram_unpack_kernel<8><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 12: gram_unpack_kernel<12><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 16: gram_unpack_kernel<16><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 20: gram_unpack_kernel<20><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 24: gram_unpack_kernel<24><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 28: gram_unpack_kernel<28><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
    case 32: gram_unpack_kernel<32><<<nb, 256, 0, st>>>(gram_packed, k, out); break;
  }
  QMFB_CUDA(cudaGetLastError());
  return QMFB_OK;
}

int qmfb_wals_solve_dev(void* stream, double* X, int64_t ldx, int64_t row_offset, const double* Y, int64_t ldy, int k,
                        const int64_t* row_ptr, const int32_t* col, const double* val, const int32_t* order,
                        int64_t nrows, int64_t nnz, const double* gram_packed, double alpha, double lambda, double* row_loss,
                        double* loss_sum, int32_t* scratch) {
  return qmfb_wals_solve_peers_dev(stream, X, ldx, row_offset, Y, ldy, k, row_ptr, col, val, order, nrows, nnz, gram_packed, alpha,
                                   lambda, row_loss, loss_sum, scratch, nullptr, 0);
}

int qmfb_wals_solve_peers_dev(void* stream, double* X, int64_t ldx, int64_t row_offset, const double* Y, int64_t ldy, int k,
                              const int64_t* row_ptr, const int32_t* col, const double* val, const int32_t* order,
                              int64_t nrows, int64_t nnz_hint, const double* gram_packed, double alpha, double lambda, double* row_loss,
                              double* loss_sum, int32_t* scratch, double* const* peer_X, int npeers) {
  const int kp = qmfb_padded_k(k);
  if (kp < 0) return kp;
  if (ldx < kp || ldy < kp || nrows < 0 || nrows > INT32_MAX || !X || !Y || !row_ptr || !order || !gram_packed || !row_loss ||
      !loss_sum || !scratch || npeers < 0 || npeers > kMaxPeers || (npeers > 0 && !peer_X)) {
    return set_error(QMFB_ERR_INVALID, "qmfb_wals_solve_dev: bad argument");
  }
  // the gather uses 16-byte cp.async on Y + row * ldy
  if ((ldy & 1) || (reinterpret_cast<uintptr_t>(Y) & 15)) {
    return set_error(QMFB_ERR_INVALID, "qmfb_wals_solve_dev: Y must be 16-byte aligned with an even row stride");
  }
  SolveParams prm{X, ldx, row_offset, Y, ldy, k, row_ptr, col, val, order, int(nrows), gram_packed, alpha, lambda,
                  row_loss, scratch + 1, npeers, {}, nullptr, 0};
  {
    // TMA gathers only when the gathered matrix is comfortably L2 resident (see SolveParams::tma_gather);
    // the launcher does not know the row count of Y: the largest column index bounds it from the CSR's side,
    // the caller's hint (ncols_hint) from the engine's.  QMFB_TMA_GATHER=0|1 overrides (measurements).
    static const char* e = getenv("QMFB_TMA_GATHER");
    if (e != nullptr) {
      prm.tma_gather = e[0] == '1';
    } else {
      prm.tma_gather = g_gather_rows_hint.load() > 0 && g_gather_rows_hint.load() * int64_t(kp) * 8 <= (int64_t(48) << 20);
    }
  }
  for (int p = 0; p < npeers; ++p) {
    if (!peer_X[p]) return set_error(QMFB_ERR_INVALID, "qmfb_wals_solve_peers_dev: null peer pointer %d", p);
    prm.peerX[p] = peer_X[p];
  }
  auto st = static_cast<cudaStream_t>(stream);
  switch (kp / 8) {
    case 4: return launch_solve<4>(st, prm, loss_sum, scratch, nnz_hint);
    case 8: return launch_solve<8>(st, prm, loss_sum, scratch, nnz_hint);
    case 12: return launch_solve<12>(st, prm, loss_sum, scratch, nnz_hint);
    case 16: return launch_solve<16>(st, prm, loss_sum, scratch, nnz_hint);
    case 20: return launch_solve_big<20>(st, prm, loss_sum, scratch);
    case 24: return launch_solve_big<24>(st, prm, loss_sum, scratch);
    case 28: return launch_solve_big<28>(st, prm, loss_sum, scratch);
    case 32: return launch_solve_big<32>(st, prm, loss_sum, scratch);
  }
  return set_error(QMFB_ERR_UNSUPPORTED, "unsupported padded k %d", kp);
}

// ------------------------------------------------------------------------------------------
// engine level
// ------------------------------------------------------------------------------------------
struct qmfb_wals {
  int device = 0;
  int64_t n[2] = {0, 0};
  int k = 0, kp = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // device-to-host copy of the user factors, overlapped with the item half-step
  cudaEvent_t ev_user = nullptr, ev_copy = nullptr;
  double* F[2] = {nullptr, nullptr};
  int64_t row_begin[2] = {0, 0}, nrows[2] = {0, 0}, nnz[2] = {0, 0};
  int64_t* row_ptr[2] = {nullptr, nullptr};
  int32_t* col[2] = {nullptr, nullptr};
  double* val[2] = {nullptr, nullptr};
  int32_t* order[2] = {nullptr, nullptr};
  double *gram_ws = nullptr, *gram_packed = nullptr, *gram_full = nullptr, *row_loss = nullptr, *loss_sum = nullptr;
  int32_t* scratch = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  float gram_ms = 0.f, solve_ms = 0.f;
  int64_t launches = 0;
};

static int wals_release(qmfb_wals* h) {
  cudaSetDevice(h->device);
  for (int s = 0; s < 2; ++s) {
    cudaFree(h->F[s]);
    cudaFree(h->row_ptr[s]);
    cudaFree(h->col[s]);
    cudaFree(h->val[s]);
    cudaFree(h->order[s]);
  }
  cudaFree(h->gram_ws);
  cudaFree(h->gram_packed);
  cudaFree(h->gram_full);
  cudaFree(h->row_loss);
  cudaFree(h->loss_sum);
  cudaFree(h->scratch);
  for (auto& e : h->ev) {
    if (e) cudaEventDestroy(e);
  }
  if (h->ev_user) cudaEventDestroy(h->ev_user);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return QMFB_OK;
}

int qmfb_wals_create(int device, int64_t nusers, int64_t nitems, int nfactors, qmfb_wals_t** out) {
  if (!out || nusers < 1 || nitems < 1) return set_error(QMFB_ERR_INVALID, "qmfb_wals_create: bad argument");
  const int kp = qmfb_padded_k(nfactors);
  if (kp < 0) return kp;
  QMFB_CUDA(cudaSetDevice(device));
  auto* h = new qmfb_wals;
  h->device = device;
  h->n[0] = nusers;
  h->n[1] = nitems;
  h->k = nfactors;
  h->kp = kp;
  int rc = [&]() -> int {
    QMFB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    QMFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    QMFB_CUDA(cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
    QMFB_CUDA(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    for (int s = 0; s < 2; ++s) {
      QMFB_CUDA(cudaMalloc(&h->F[s], size_t(h->n[s]) * kp * sizeof(double)));
      QMFB_CUDA(cudaMemsetAsync(h->F[s], 0, size_t(h->n[s]) * kp * sizeof(double), h->stream));
    }
    QMFB_CUDA(cudaMalloc(&h->gram_ws, size_t(qmfb_gram_workspace_len(nfactors)) * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->gram_packed, size_t(qmfb_gram_packed_len(nfactors)) * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->gram_full, size_t(nfactors) * nfactors * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->row_loss, size_t(std::max(nusers, nitems)) * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->loss_sum, sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->scratch, 2 * sizeof(int32_t)));
    for (auto& e : h->ev) QMFB_CUDA(cudaEventCreate(&e));
    QMFB_CUDA(cudaStreamSynchronize(h->stream));
    return QMFB_OK;
  }();
  if (rc != QMFB_OK) {
    wals_release(h);
    return rc;
  }
  *out = h;
  return QMFB_OK;
}

int qmfb_wals_destroy(qmfb_wals_t* h) {
  if (!h) return QMFB_OK;
  return wals_release(h);
}

int qmfb_wals_set_csr(qmfb_wals_t* h, int side, int64_t row_begin, int64_t nrows, const int64_t* row_ptr,
                      const int32_t* col_idx, const double* val) {
  if (!h || side < 0 || side > 1 || !row_ptr || nrows < 0 || row_begin < 0 || row_begin + nrows > h->n[side] || row_ptr[0] != 0) {
    return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_csr: bad argument");
  }
  const int64_t nnz = row_ptr[nrows];
  if (nnz > 0 && (!col_idx || !val)) return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_csr: null col/val");
  const int64_t ncols = h->n[1 - side];
  for (int64_t r = 0; r < nrows; ++r) {
    if (row_ptr[r + 1] < row_ptr[r]) return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_csr: row_ptr not monotone at %lld", (long long)r);
  }
  for (int64_t p = 0; p < nnz; ++p) {
    if (col_idx[p] < 0 || col_idx[p] >= ncols) return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_csr: col_idx[%lld]=%d out of range", (long long)p, col_idx[p]);
  }
  QMFB_CUDA(cudaSetDevice(h->device));
  cudaFree(h->row_ptr[side]);
  cudaFree(h->col[side]);
  cudaFree(h->val[side]);
  cudaFree(h->order[side]);
  h->row_ptr[side] = nullptr; h->col[side] = nullptr; h->val[side] = nullptr; h->order[side] = nullptr;
  // longest rows first: the persistent solve kernel pulls rows through an atomic counter
  std::vector<int32_t> order(static_cast<size_t>(nrows));
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [row_ptr](int32_t a, int32_t b) {
    return row_ptr[a + 1] - row_ptr[a] > row_ptr[b + 1] - row_ptr[b];
  });
  QMFB_CUDA(cudaMalloc(&h->row_ptr[side], size_t(nrows + 1) * sizeof(int64_t)));
  QMFB_CUDA(cudaMalloc(&h->col[side], size_t(std::max<int64_t>(nnz, 1)) * sizeof(int32_t)));
  QMFB_CUDA(cudaMalloc(&h->val[side], size_t(std::max<int64_t>(nnz, 1)) * sizeof(double)));
  QMFB_CUDA(cudaMalloc(&h->order[side], size_t(std::max<int64_t>(nrows, 1)) * sizeof(int32_t)));
  QMFB_CUDA(cudaMemcpyAsync(h->row_ptr[side], row_ptr, size_t(nrows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  if (nnz > 0) {
    QMFB_CUDA(cudaMemcpyAsync(h->col[side], col_idx, size_t(nnz) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    QMFB_CUDA(cudaMemcpyAsync(h->val[side], val, size_t(nnz) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (nrows > 0) {
    QMFB_CUDA(cudaMemcpyAsync(h->order[side], order.data(), size_t(nrows) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  }
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  h->row_begin[side] = row_begin;
  h->nrows[side] = nrows;
  h->nnz[side] = nnz;
  return QMFB_OK;
}

int qmfb_wals_set_signals(qmfb_wals_t* h, const qmfb_signals_t* s) {
  int64_t nu = 0, ni = 0, nnz = 0;
  if (!h || !s || qmfb_signals_dims(s, &nu, &ni, &nnz) != QMFB_OK) return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_signals: bad argument");
  if (nu != h->n[0] || ni != h->n[1]) {
    return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_signals: engine is %lld x %lld, signals are %lld x %lld", (long long)h->n[0],
                     (long long)h->n[1], (long long)nu, (long long)ni);
  }
  QMFB_CUDA(cudaSetDevice(h->device));
  for (int side = 0; side < 2; ++side) {
    const int64_t* rp = nullptr;
    const int32_t* col = nullptr;
    const double* val = nullptr;
    const int32_t* order = nullptr;
    if (int rc = qmfb_signals_device(s, side, &rp, &col, &val, &order)) return rc;
    const int64_t nrows = h->n[side];
    cudaFree(h->row_ptr[side]);
    cudaFree(h->col[side]);
    cudaFree(h->val[side]);
    cudaFree(h->order[side]);
    h->row_ptr[side] = nullptr; h->col[side] = nullptr; h->val[side] = nullptr; h->order[side] = nullptr;
    QMFB_CUDA(cudaMalloc(&h->row_ptr[side], size_t(nrows + 1) * sizeof(int64_t)));
    QMFB_CUDA(cudaMalloc(&h->col[side], size_t(nnz) * sizeof(int32_t)));
    QMFB_CUDA(cudaMalloc(&h->val[side], size_t(nnz) * sizeof(double)));
    QMFB_CUDA(cudaMalloc(&h->order[side], size_t(nrows) * sizeof(int32_t)));
    QMFB_CUDA(cudaMemcpyAsync(h->row_ptr[side], rp, size_t(nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, h->stream));
    QMFB_CUDA(cudaMemcpyAsync(h->col[side], col, size_t(nnz) * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    QMFB_CUDA(cudaMemcpyAsync(h->val[side], val, size_t(nnz) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    QMFB_CUDA(cudaMemcpyAsync(h->order[side], order, size_t(nrows) * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    h->row_begin[side] = 0;
    h->nrows[side] = nrows;
    h->nnz[side] = nnz;
  }
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return QMFB_OK;
}

int qmfb_wals_set_factors(qmfb_wals_t* h, int side, const double* host) {
  if (!h || side < 0 || side > 1 || !host) return set_error(QMFB_ERR_INVALID, "qmfb_wals_set_factors: bad argument");
  QMFB_CUDA(cudaSetDevice(h->device));
  QMFB_CUDA(cudaMemcpy2DAsync(h->F[side], size_t(h->kp) * 8, host, size_t(h->k) * 8, size_t(h->k) * 8, size_t(h->n[side]),
                              cudaMemcpyHostToDevice, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return QMFB_OK;
}

int qmfb_wals_get_factors(qmfb_wals_t* h, int side, double* host) {
  if (!h || side < 0 || side > 1 || !host) return set_error(QMFB_ERR_INVALID, "qmfb_wals_get_factors: bad argument");
  QMFB_CUDA(cudaSetDevice(h->device));
  QMFB_CUDA(cudaMemcpy2DAsync(host, size_t(h->k) * 8, h->F[side], size_t(h->kp) * 8, size_t(h->k) * 8, size_t(h->n[side]),
                              cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return QMFB_OK;
}

int qmfb_wals_gram(qmfb_wals_t* h, int side, double* host_out) {
  if (!h || side < 0 || side > 1 || !host_out) return set_error(QMFB_ERR_INVALID, "qmfb_wals_gram: bad argument");
  QMFB_CUDA(cudaSetDevice(h->device));
  int rc = qmfb_gram_dev(h->stream, h->F[side], h->kp, 0, h->n[side], h->k, h->gram_ws, h->gram_packed);
  if (rc) return rc;
  rc = qmfb_gram_unpack_dev(h->stream, h->gram_packed, h->k, h->gram_full);
  if (rc) return rc;
  h->launches += 3;
  QMFB_CUDA(cudaMemcpyAsync(host_out, h->gram_full, size_t(h->k) * h->k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return QMFB_OK;
}

static int wals_half_step_async(qmfb_wals* h, int side, double alpha, double lambda) {
  const int other = 1 - side;
  if (!h->row_ptr[side]) return set_error(QMFB_ERR_INVALID, "qmfb_wals_half_step: no CSR uploaded for side %d", side);
  // leftData.setFactors(0), WALSEngine.cpp:170-171 (this shard's rows)
  QMFB_CUDA(cudaMemsetAsync(h->F[side] + h->row_begin[side] * h->kp, 0, size_t(h->nrows[side]) * h->kp * sizeof(double), h->stream));
  QMFB_CUDA(cudaEventRecord(h->ev[0], h->stream));
  int rc = qmfb_gram_dev(h->stream, h->F[other], h->kp, 0, h->n[other], h->k, h->gram_ws, h->gram_packed);
  if (rc) return rc;
  QMFB_CUDA(cudaEventRecord(h->ev[1], h->stream));
  rc = qmfb_wals_solve_dev(h->stream, h->F[side], h->kp, h->row_begin[side], h->F[other], h->kp, h->k, h->row_ptr[side],
                           h->col[side], h->val[side], h->order[side], h->nrows[side], h->nnz[side], h->gram_packed, alpha, lambda,
                           h->row_loss, h->loss_sum, h->scratch);
  if (rc) return rc;
  QMFB_CUDA(cudaEventRecord(h->ev[2], h->stream));
  // gram_partial, gram_reduce, [long_row_partial, long_row_reduce (k <= 128)], wals_solve, sum
  h->launches += (h->nrows[side] > 0 && h->kp <= 128) ? 7 : 4;
  return QMFB_OK;
}

static int wals_finish_step(qmfb_wals* h, double* loss_sum) {
  double loss = 0.0;
  int32_t scratch[2] = {0, 0};
  QMFB_CUDA(cudaMemcpyAsync(&loss, h->loss_sum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaMemcpyAsync(scratch, h->scratch, sizeof(scratch), cudaMemcpyDeviceToHost, h->stream));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  QMFB_CUDA(cudaEventElapsedTime(&h->gram_ms, h->ev[0], h->ev[1]));
  QMFB_CUDA(cudaEventElapsedTime(&h->solve_ms, h->ev[1], h->ev[2]));
  if (scratch[1] != 0) return set_error(QMFB_ERR_NOT_SPD, "normal equations not positive definite (reference: dsysv failed)");
  if (loss_sum) *loss_sum = loss;
  return QMFB_OK;
}

int qmfb_wals_half_step(qmfb_wals_t* h, int update_side, double alpha, double lambda, double* loss_sum) {
  if (!h || update_side < 0 || update_side > 1) return set_error(QMFB_ERR_INVALID, "qmfb_wals_half_step: bad argument");
  QMFB_CUDA(cudaSetDevice(h->device));
  int rc = wals_half_step_async(h, update_side, alpha, lambda);
  if (rc) return rc;
  return wals_finish_step(h, loss_sum);
}

int qmfb_wals_epoch_host(qmfb_wals_t* h, double alpha, double lambda, const double* item_factors_in,
                         double* user_factors_out, double* item_factors_out, double* loss_out) {
  if (!h) return set_error(QMFB_ERR_INVALID, "qmfb_wals_epoch_host: null handle");
  QMFB_CUDA(cudaSetDevice(h->device));
  if (item_factors_in) {
    QMFB_CUDA(cudaMemcpy2DAsync(h->F[1], size_t(h->kp) * 8, item_factors_in, size_t(h->k) * 8, size_t(h->k) * 8,
                                size_t(h->n[1]), cudaMemcpyHostToDevice, h->stream));
  }
  int rc = wals_half_step_async(h, QMFB_SIDE_USER, alpha, lambda);
  if (rc) return rc;
  int32_t err_user[2] = {0, 0};
  QMFB_CUDA(cudaMemcpyAsync(err_user, h->scratch, sizeof(err_user), cudaMemcpyDeviceToHost, h->stream));
  if (user_factors_out) {
    // the user factors are final after the user half-step: their copy to the host runs on a second
    // stream underneath the item half-step (which only reads them)
    QMFB_CUDA(cudaEventRecord(h->ev_user, h->stream));
    QMFB_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_user, 0));
    QMFB_CUDA(cudaMemcpy2DAsync(user_factors_out, size_t(h->k) * 8, h->F[0], size_t(h->kp) * 8, size_t(h->k) * 8,
                                size_t(h->n[0]), cudaMemcpyDeviceToHost, h->copy_stream));
    QMFB_CUDA(cudaEventRecord(h->ev_copy, h->copy_stream));
  }
  rc = wals_half_step_async(h, QMFB_SIDE_ITEM, alpha, lambda);
  if (rc) return rc;
  if (user_factors_out) QMFB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
  if (item_factors_out) {
    QMFB_CUDA(cudaMemcpy2DAsync(item_factors_out, size_t(h->k) * 8, h->F[1], size_t(h->kp) * 8, size_t(h->k) * 8,
                                size_t(h->n[1]), cudaMemcpyDeviceToHost, h->stream));
  }
  double loss = 0.0;
  rc = wals_finish_step(h, &loss);
  if (rc) return rc;
  if (err_user[1] != 0) return set_error(QMFB_ERR_NOT_SPD, "normal equations not positive definite (user half-step)");
  // loss / nusers / nitems, WALSEngine.cpp:215
  if (loss_out) *loss_out = loss / double(h->n[0]) / double(h->n[1]);
  return QMFB_OK;
}

int qmfb_wals_eval_rank(qmfb_wals_t* h, const int32_t* test_users, int64_t nT, const int64_t* label_ptr, const int32_t* label_items,
                        int32_t* cnt, double* pos_scores) {
  if (!h) return set_error(QMFB_ERR_INVALID, "qmfb_wals_eval_rank: null handle");
  QMFB_CUDA(cudaSetDevice(h->device));
  QMFB_CUDA(cudaStreamSynchronize(h->stream));
  return eval_rank_resident(h->device, h->F[0], h->kp, h->n[0], h->F[1], h->kp, h->n[1], h->k, nullptr, test_users, nT, label_ptr,
                            label_items, cnt, pos_scores);
}

double* qmfb_wals_factors_device(qmfb_wals_t* h, int side) { return (h && side >= 0 && side <= 1) ? h->F[side] : nullptr; }
void* qmfb_wals_stream(qmfb_wals_t* h) { return h ? h->stream : nullptr; }
int64_t qmfb_wals_launch_count(qmfb_wals_t* h) { return h ? h->launches : 0; }
int qmfb_wals_last_timing(qmfb_wals_t* h, float* gram_ms, float* solve_ms) {
  if (!h) return set_error(QMFB_ERR_INVALID, "null handle");
  if (gram_ms) *gram_ms = h->gram_ms;
  if (solve_ms) *solve_ms = h->solve_ms;
  return QMFB_OK;
}

}  // extern "C"
