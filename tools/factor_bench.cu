// Isolated timing of factor_diag_tile variants (one warp / several warps per SM).
#include <cstdio>
#include "../qmf_b200/csrc/wals_kernels.cuh"
using namespace qmfb;

// variant B: no inverse (E) part, just the fraction-free factor; U scaled at the end
__device__ __noinline__ bool factor_noinv(const double* tile, double* utile, int lane) {
  const int r = lane >> 2, q = lane & 3;
  const double2 a = *reinterpret_cast<const double2*>(tile + lane * 2);
  double a0 = a.x, a1 = a.y, S = 1.0, prS = 1.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int src = 4 * j;
    const double p = __shfl_sync(0xffffffffu, (j & 1) ? a1 : a0, src + (j >> 1));
    const double uc0 = __shfl_sync(0xffffffffu, a0, src + q);
    const double uc1 = __shfl_sync(0xffffffffu, a1, src + q);
    const double t0 = __shfl_sync(0xffffffffu, a0, src + (r >> 1));
    const double t1 = __shfl_sync(0xffffffffu, a1, src + (r >> 1));
    const double ur = (r & 1) ? t1 : t0;
    ok = ok && (p > 0.0);
    const int hi = __double2hiint(p), lo = __double2loint(p);
    const double pn = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, lo);
    const double sc = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
    if (r == j) prS = p * S;
    S *= pn;
    if (r > j) {
      a0 = fma(a0, pn, -((ur * uc0) * sc));
      a1 = fma(a1, pn, -((ur * uc1) * sc));
    }
  }
  const double g = rsqrt(prS);
  *reinterpret_cast<double2*>(utile + lane * 2) = make_double2(a0 * g, a1 * g);
  return ok;
}

// variant C: row j broadcast through shared memory instead of shuffles (with inverse part)
__device__ __noinline__ bool factor_smem(const double* tile, double* wtile, double* scratch, int lane) {
  const int r = lane >> 2, q = lane & 3;
  const double2 a = *reinterpret_cast<const double2*>(tile + lane * 2);
  double a0 = a.x, a1 = a.y;
  double e0 = (2 * q == r) ? 1.0 : 0.0, e1 = (2 * q + 1 == r) ? 1.0 : 0.0;
  double S = 1.0, prS = 1.0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
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
    if (r > j) {
      a0 = fma(a0, pn, -((ur * uc.x) * sc));
      a1 = fma(a1, pn, -((ur * uc.y) * sc));
      e0 = fma(e0, pn, -((ur * ec.x) * sc));
      e1 = fma(e1, pn, -((ur * ec.y) * sc));
    }
  }
  const double g = rsqrt(prS);
  wtile[(2 * q) * 8 + r] = e0 * g;
  wtile[(2 * q + 1) * 8 + r] = e1 * g;
  return ok;
}

template <int V>
__global__ void kern(double* out, long long* cyc, int iters) {
  __shared__ double tile[8][64], w[8][64], scr[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 64; i += 32) tile[warp][i] = (i / 8 == i % 8) ? 8.0 + i * 0.01 : 0.3 + 0.001 * i;
  __syncwarp();
  bool ok = true;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double2 a = *reinterpret_cast<const double2*>(tile[warp] + tile_acc_off(lane));
    if (V == 0) ok = factor_diag_tile(a.x, a.y, w[warp], scr[warp], lane) && ok;
    if (V == 3) ok = factor_diag_tile_mma(a.x, a.y, w[warp], lane) && ok;
    if (V == 1) ok = factor_noinv(tile[warp], w[warp], lane) && ok;
    if (V == 2) ok = factor_smem(tile[warp], w[warp], scr[warp], lane) && ok;
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[V] = (t1 - t0) / iters;
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    out[V * 64 + lane] = w[0][lane] + (ok ? 0.0 : 1e9);
    out[V * 64 + 32 + lane] = w[0][32 + lane];
  }
}

int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  for (int warps : {1, 8}) {
    kern<0><<<148, 32 * warps>>>(out, cyc, 200);
    kern<1><<<148, 32 * warps>>>(out, cyc, 200);
    kern<2><<<148, 32 * warps>>>(out, cyc, 200);
    kern<3><<<148, 32 * warps>>>(out, cyc, 200);
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    double w[256]; cudaMemcpy(w, out, sizeof(w), cudaMemcpyDeviceToHost);
    double d = 0, m = 0;
    for (int i = 0; i < 64; ++i) { d = fmax(d, fabs(w[i] - w[192 + i])); m = fmax(m, fabs(w[i])); }
    printf("warps/CTA=%d: smem-broadcast+inverse (product) %lld  shuffle no-inverse %lld  smem unrolled %lld  DMMA-broadcast+inverse %lld cycles/tile; max |W_mma - W_smem| = %.3e (max |W| %.3e)\n", warps, h[0], h[1], h[2], h[3], d, m);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
