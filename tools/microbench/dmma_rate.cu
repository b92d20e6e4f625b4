// Microbenchmark: FP64 FMA throughput per SM of mma.sync.m8n8k4.f64 (DMMA) vs plain DFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int which = 0; which < 2; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) dmma_kernel<<<148, threads>>>(out, iters); else dfma_kernel<<<148, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma_per_thread = which == 0 ? (double)iters * 8 * 8 : (double)iters * 16;   // DMMA: 256 FMA / 32 lanes
      double total = fma_per_thread * threads * 148;
      printf("%s threads/SM=%4d : %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM at %.0f MHz nominal)\n", which == 0 ? "DMMA" : "DFMA",
             threads, ms, total / ms / 1e9, total / 148 / (ms * 1e-3) / (clk_khz * 1e3), clk_khz / 1e3);
    }
  }
  return 0;
}
