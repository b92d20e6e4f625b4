// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, K=8) as a function of N, of where A lives
// (tensor memory vs shared memory) and of the number of distinct accumulators.  One CTA per SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int M, int a_in_tmem, int naccum, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t bdesc = desc(smem_u32(smem), (uint32_t)N * 16u, 128u);
    const uint64_t adesc = desc(smem_u32(smem) + 32768, (uint32_t)M * 16u, 128u);
    const long long t0 = clock64();
    const uint32_t mask = (uint32_t)naccum - 1u;  // naccum is a power of two
    if (a_in_tmem) {
#pragma unroll 4
      for (int i = 0; i < iters; ++i) mma_ts(tmem + ((uint32_t)i & mask) * (uint32_t)N, tmem + 480, bdesc, idesc, 1u);
    } else {
#pragma unroll 4
      for (int i = 0; i < iters; ++i) mma_ss(tmem + ((uint32_t)i & mask) * (uint32_t)N, adesc, bdesc, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      if (clock64() - t0 > 2000000000LL) break;
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  printf("M N a_in_tmem naccum grid : issue_clk/mma total_clk/mma\n");
  for (int grid : {1, 148})
    for (int M : {128, 64})
      for (int a_tm : {1, 0})
        for (int N : {32, 64, 128, 256})
          for (int nacc : {1, 4}) {
            if (nacc * N > 448) continue;
            if (M == 128 && N % 16) continue;
            rate_kernel<<<grid, 128, 64 * 1024>>>(N, M, a_tm, nacc, iters, out);
            long long h[2] = {0, 0};
            cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%3d %3d %d %d %3d : %7.1f %7.1f\n", M, N, a_tm, nacc, grid, (double)h[0] / iters, (double)h[1] / iters);
          }
  return 0;
}
