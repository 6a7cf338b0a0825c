// Dependent-chain latencies on B200 (sm_100a): DMMA, DFMA, DMUL, LDS.64 -> use, SHFL, rsqrt, 1/x.
// One warp per SM-sized grid is irrelevant here: a single warp, clock64() around an unrolled chain.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void lat(double* out, long long* cyc, double x, double y) {
  __shared__ double sm[64];
  sm[threadIdx.x] = threadIdx.x * 1e-3 + 1.0;
  sm[threadIdx.x + 32] = 0;
  __syncthreads();
  const int N = 256;
  double c0 = x, c1 = y, a = 1.0000001, b = 0.999999;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma(c0, c1, a, b);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  double f = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fma(f, a, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  double m = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) m = m * a;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // LDS chain: address depends on loaded value
  int idx = threadIdx.x & 31;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx + 32] + (threadIdx.x & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  double s = x + threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) s = __shfl_xor_sync(0xffffffffu, s, 1);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  double r = x + 2.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) r = rsqrt(r) + 2.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  double q = x + 2.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) q = __drcp_rn(q) + 2.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  // two independent DMMA chains interleaved (throughput of a single warp)
  double d0 = x, d1 = y, e0 = y, e1 = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { dmma(d0, d1, a, b); dmma(e0, e1, a, b); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0);
  out[threadIdx.x] = c0 + c1 + f + m + idx + s + r + q + d0 + d1 + e0 + e1;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  lat<<<1, 32>>>(out, cyc, 1.0, 2.0); lat<<<1, 32>>>(out, cyc, 1.0, 2.0);
  long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
  const char* names[8] = {"DMMA dependent", "DFMA dependent", "DMUL dependent", "LDS.64 dependent (+cvt)", "SHFL.f64 dependent", "rsqrt(double)+add", "__drcp_rn+add", "2 indep DMMA chains (per pair)"};
  for (int i = 0; i < 8; ++i) printf("%-34s %7.1f cycles/op\n", names[i], h[i] / 256.0);
  return 0;
}
