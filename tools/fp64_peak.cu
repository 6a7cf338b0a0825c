// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Output: one line per (variant, warps/CTA, CTAs/SM) with TFLOP/s measured by CUDA events.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double x, double y) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double x, double y) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  double a = x + threadIdx.x * 1e-12, b = y;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double x, double y) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[4] = {x, x + 1e-12, x + 2e-12, x + threadIdx.x * 1e-12}, b[2] = {y, y * 0.5};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double x, double y) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-12 + threadIdx.x * 1e-13;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = y + i * 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// smem-fed DMMA: mimics the WALS build inner loop (per k4-step: NB B-frag loads, 2 A scalings, NB+1 mmas)
template <int NB>
__global__ void k_dmma_smem(double* out, int iters, int ld) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 32 * ld; i += blockDim.x) sm[i] = 1e-6 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double c[NB + 1][2];
#pragma unroll
  for (int i = 0; i <= NB; ++i) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s0 = 0; s0 < 32; s0 += 4) {
      const double* row = sm + (s0 + (lane & 3)) * ld + (lane >> 2);
      double b[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) b[j] = row[8 * j];
      double a0 = b[0] * 1.000001, a1 = b[NB - 1] * 0.99999;
#pragma unroll
      for (int j = 0; j < NB; ++j) dmma884(c[j][0], c[j][1], a0, b[j]);
      dmma884(c[NB][0], c[NB][1], a1, b[NB - 1]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i <= NB; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms=%d l2=%d MB smem/blk optin=%zu clock=%d kHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20,
         p.sharedMemPerBlockOptin, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1024));
  const int sms = p.multiProcessorCount;
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int cps : {1, 2}) {
      if (warps * cps > 64) continue;
      dim3 grid(sms * cps), blk(warps * 32);
      double nthreads = (double)grid.x * blk.x, nwarps = nthreads / 32;
      float ms;
      ms = timeit([&] { k_dfma<8><<<grid, blk>>>(out, iters, 1.0000001, 1e-9); });
      printf("dfma        warps/cta=%2d cta/sm=%d  %.2f TF\n", warps, cps, nthreads * iters * 8 * 2 / ms * 1e-9);
      ms = timeit([&] { k_dmma884<8><<<grid, blk>>>(out, iters, 1.0000001, 1e-9); });
      printf("dmma884     warps/cta=%2d cta/sm=%d  %.2f TF\n", warps, cps, nwarps * iters * 8 * 512.0 / ms * 1e-9);
      ms = timeit([&] { k_dmma1688<8><<<grid, blk>>>(out, iters, 1.0000001, 1e-9); });
      printf("dmma1688    warps/cta=%2d cta/sm=%d  %.2f TF\n", warps, cps, nwarps * iters * 8 * 2048.0 / ms * 1e-9);
      ms = timeit([&] { k_dmma16816<8><<<grid, blk>>>(out, iters / 2, 1.0000001, 1e-9); });
      printf("dmma16816   warps/cta=%2d cta/sm=%d  %.2f TF\n", warps, cps, nwarps * (iters / 2) * 8 * 4096.0 / ms * 1e-9);
    }
  }
  // smem-fed variants: 8 warps/CTA, ld = 136 (bank-conflict-free) vs 128 (4-way conflicts)
  for (int ld : {136, 128}) {
    for (int cps : {1, 2}) {
      dim3 grid(sms * cps), blk(256);
      size_t smem = 32 * ld * sizeof(double);
      float ms;
      ms = timeit([&] { k_dmma_smem<16><<<grid, blk, smem>>>(out, 2000, ld); });
      printf("dmma_smem NB=16 ld=%d cta/sm=%d  %.2f TF\n", ld, cps, (double)grid.x * 8 * 2000 * 8 * 17 * 512.0 / ms * 1e-9);
      ms = timeit([&] { k_dmma_smem<9><<<grid, blk, smem>>>(out, 2000, ld); });
      printf("dmma_smem NB=9  ld=%d cta/sm=%d  %.2f TF\n", ld, cps, (double)grid.x * 8 * 2000 * 8 * 10 * 512.0 / ms * 1e-9);
    }
  }
  // HBM copy bandwidth sanity (1 GiB)
  {
    size_t n = 1ull << 30; char *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n));
    float ms = timeit([&] { cudaMemcpyAsync(b, a, n, cudaMemcpyDeviceToDevice); });
    printf("memcpy d2d 1GiB: %.1f GB/s (read+write)\n", 2.0 * n / ms * 1e-6);
  }
  return 0;
}
