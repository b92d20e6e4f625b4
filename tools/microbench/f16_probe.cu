// Probe for the 16-bit X sweep (round 2): answers, in ONE run on a B200,
//   1. does tcgen05.mma.kind::f16 accept A = f16 with B = bf16 (mixed operand formats)?
//   2. how are two f16 K-elements packed into a 32-bit tensor-memory column when A comes from TMEM?
//   3. which (LBO, SBO) assignment does an MN-major SWIZZLE_128B shared-memory descriptor for A expect,
//      when the tile was staged by TMA as 64-column x R-row boxes of a row-major [K][M] matrix?
//   4. how fast does TMA stream an fp16 matrix with (a) 256-column no-swizzle boxes, (b) 64-column
//      SWIZZLE_128B boxes of various heights?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a f16_probe.cu -o f16_probe -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000LL) return false;
  return true;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (TmapEncodeFn)fn;
}
static CUtensorMap make_map(const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_cols,
                            uint32_t box_rows, bool swz128) {
  static TmapEncodeFn enc = get_encode();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(1);
  }
  return m;
}

// ---------------------------------------------------------------------------------------------------------
// semantics: D[m][n] = sum_k A[k][m] * B[n][k],  M = 128, N = 32, K = 16 * ksteps
// mode 0: A from TMEM, packed (k even -> low half);  mode 1: A from smem, MN-major SW128 (lbo, sbo given)
// ---------------------------------------------------------------------------------------------------------
constexpr int SEM_N = 32;
__global__ void __launch_bounds__(128, 1)
sem_kernel(const __grid_constant__ CUtensorMap amap, const __half* __restrict__ A, const uint16_t* __restrict__ Bc,
           int mode, int bfmt, uint32_t a_lbo, uint32_t a_sbo, int box_rows, int ksteps, float* __restrict__ D,
           int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* al = raw + (base - smem_u32(raw));
  const uint32_t a_smem = base;                 // A boxes: 2 boxes x box_rows x 128 B
  const uint32_t b_smem = base + 32768;         // B: ksteps x (2 chunks x N x 16 B)
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B: already in the canonical K-major layout in global memory: plain copy
  for (int i = tid; i < ksteps * 2 * SEM_N * 8; i += 128) reinterpret_cast<uint16_t*>(al + 32768)[i] = Bc[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
  bool ok = true;
  if (mode == 0) {
    // every thread = lane m: pack A[k][m] pairs, 8 TMEM columns per K-step, operand columns start at 64
    for (int u = 0; u < ksteps; ++u) {
      uint32_t r[8];
      for (int c = 0; c < 8; ++c) {
        const uint16_t lo = __half_as_ushort(A[(u * 16 + 2 * c) * 128 + tid]);
        const uint16_t hi = __half_as_ushort(A[(u * 16 + 2 * c + 1) * 128 + tid]);
        r[c] = (uint32_t)lo | ((uint32_t)hi << 16);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       tmem + lane_field + 64 + u * 8),
                   "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  } else {
    if (tid == 0) {
      mbar_expect_tx(bar_tma, 2u * box_rows * 128u);
      tma_load_2d(a_smem, &amap, 0, 0, bar_tma);
      tma_load_2d(a_smem + box_rows * 128, &amap, 64, 0, bar_tma);
    }
    ok = mbar_wait(bar_tma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    // D = f32, A = f16 (0), B = bf16 (1) or f16 (0), M = 128, N = 32; a_major = MN for the smem mode
    const uint32_t idesc = (1u << 4) | (0u << 7) | ((uint32_t)bfmt << 10) | ((mode == 0 ? 0u : 1u) << 15) |
                           ((uint32_t)(SEM_N >> 3) << 17) | ((128u >> 4) << 24);
    for (int u = 0; u < ksteps; ++u) {
      const uint64_t bdesc = make_desc(b_smem + u * (2 * SEM_N * 16), SEM_N * 16u, 128u, 0);
      if (mode == 0) {
        mma_f16_ts(tmem, tmem + 64 + u * 8, bdesc, idesc, u > 0);
      } else {
        const uint64_t adesc = make_desc(a_smem + u * 2048, a_lbo, a_sbo, 2);
        mma_f16_ss(tmem, adesc, bdesc, idesc, u > 0);
      }
    }
    tc_commit(bar_mma);
  }
  ok = mbar_wait(bar_mma, 0) && ok;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(tmem + lane_field)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) D[tid * 32 + n] = v[n];
  if (!ok) atomicExch(status, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

static void run_semantics(int only) {
  const int K = 32, M = 128, N = SEM_N, ksteps = K / 16, box_rows = K;
  std::vector<__half> A(K * M);
  std::vector<float> Af(K * M), Bf(N * K);
  srand(7);
  for (int i = 0; i < K * M; ++i) {
    Af[i] = (float)(rand() % 1500);  // exact counts
    A[i] = __float2half(Af[i]);
  }
  for (int i = 0; i < N * K; ++i) Bf[i] = bf16_round(0.001f + (rand() % 10000) / 3000.0f);
  // canonical K-major B per K-step: [chunk = k/8 (2)][n][k%8]
  std::vector<uint16_t> Bc_bf(ksteps * 2 * N * 8), Bc_h(ksteps * 2 * N * 8);
  for (int u = 0; u < ksteps; ++u)
    for (int c = 0; c < 2; ++c)
      for (int n = 0; n < N; ++n)
        for (int e = 0; e < 8; ++e) {
          const int k = u * 16 + c * 8 + e;
          const float v = Bf[n * K + k];
          __nv_bfloat16 b = __float2bfloat16(v);
          __half h = __float2half(v);
          Bc_bf[((u * 2 + c) * N + n) * 8 + e] = *reinterpret_cast<uint16_t*>(&b);
          Bc_h[((u * 2 + c) * N + n) * 8 + e] = *reinterpret_cast<uint16_t*>(&h);
        }
  auto expect = [&](int pack_mode, bool b_half, std::vector<double>& out) {
    // pack_mode 0: column c of a K-step holds (k = 2c, 2c+1); 1: (k = c, c + 8)
    out.assign(M * N, 0.0);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) {
          int ka = k;
          if (pack_mode == 1) {  // what the hardware would read if it took (lo, hi) of column c as k = c, c + 8
            const int u = k / 16, kk = k % 16;
            const int c = kk % 8, half = kk / 8;  // hardware k -> column c, half
            ka = u * 16 + 2 * c + half;           // the element we stored there
          }
          const float b = b_half ? __half2float(__float2half(Bf[n * K + k])) : Bf[n * K + k];
          s += (double)Af[ka * M + m] * (double)b;
        }
        out[m * N + n] = s;
      }
  };
  __half* dA;
  uint16_t *dBbf, *dBh;
  float* dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dBbf, Bc_bf.size() * 2));
  CK(cudaMalloc(&dBh, Bc_h.size() * 2));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBbf, Bc_bf.data(), Bc_bf.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBh, Bc_h.data(), Bc_h.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap amap = make_map(dA, M, K, M * 2, 64, box_rows, true);
  CK(cudaFuncSetAttribute(sem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  struct Case { int mode, bfmt; uint32_t lbo, sbo; const char* name; };
  const Case cases[] = {
      {0, 1, 0, 0, "TS  A=f16 (TMEM, packed pairs) B=bf16"},
      {0, 0, 0, 0, "TS  A=f16 (TMEM, packed pairs) B=f16 "},
      {1, 1, (uint32_t)box_rows * 128u, 1024u, "SS  A MN-major SW128 lbo=box sbo=1024, B=bf16"},
      {1, 1, 1024u, (uint32_t)box_rows * 128u, "SS  A MN-major SW128 lbo=1024 sbo=box, B=bf16"},
      {1, 0, (uint32_t)box_rows * 128u, 1024u, "SS  A MN-major SW128 lbo=box sbo=1024, B=f16 "},
  };
  std::vector<float> D(M * N);
  std::vector<double> e0, e1;
  int idx = -1;
  for (const Case& c : cases) {
    if (++idx != only && only >= 0) continue;
    CK(cudaMemset(dD, 0xff, M * N * 4));
    CK(cudaMemset(dS, 0, 4));
    sem_kernel<<<1, 128, 48 * 1024>>>(amap, dA, c.bfmt ? dBbf : dBh, c.mode, c.bfmt, c.lbo, c.sbo, box_rows, ksteps, dD, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-50s : CUDA error %s (aborting the semantic tests)\n", c.name, cudaGetErrorString(e));
      return;
    }
    int st = 0;
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    expect(0, c.bfmt == 0, e0);
    expect(1, c.bfmt == 0, e1);
    double d0 = 0, d1 = 0, ref = 0;
    for (int i = 0; i < M * N; ++i) {
      d0 = fmax(d0, fabs(D[i] - e0[i]));
      d1 = fmax(d1, fabs(D[i] - e1[i]));
      ref = fmax(ref, fabs(e0[i]));
    }
    printf("%-50s : timeout=%d  max|D-expected| pairs(2c,2c+1)=%.4g  pairs(c,c+8)=%.4g  (max|expected| %.4g)  D[0..3]=%g %g %g %g\n",
           c.name, st, d0, d1, ref, D[0], D[1], D[2], D[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA streaming rate
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, int rows, int box_cols, int box_rows, int boxes, int stages,
              int rows_per_split, int panel_rows, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bars[32];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x;
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[16]);
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(full + 8 * i, 1);
      mbar_init(empty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int col0 = blockIdx.x * box_cols * boxes;
  const int r_begin = blockIdx.y * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > rows) r_end = rows;
  const int nst = (r_end - r_begin + box_rows - 1) / box_rows;
  const uint32_t box_bytes = (uint32_t)box_cols * box_rows * 2u;
  const uint32_t stage_bytes = box_bytes * boxes;
  if (tid == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nst; ++i) {
      if (!mbar_wait(empty + 8 * s, ph ^ 1u)) { atomicExch(status, 1); break; }
      mbar_expect_tx(full + 8 * s, stage_bytes);
      for (int b = 0; b < boxes; ++b) {
        int x = col0 + b * box_cols, y = r_begin + i * box_rows;
        if (panel_rows > 0) {  // panel-major storage: [x / 256][row][256]
          y += (x / 256) * panel_rows;
          x = x % 256;
        }
        tma_load_2d(base + s * stage_bytes + b * box_bytes, &map, x, y, full + 8 * s);
      }
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (tid == 32) {
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nst; ++i) {
      if (!mbar_wait(full + 8 * s, ph)) { atomicExch(status, 2); break; }
      mbar_arrive(empty + 8 * s);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  }
}

static void run_stream(const __half* X, int rows, int64_t cols, int64_t ld, const char* what, bool panel = false,
                       int only_cfg = -1, int splits_override = 0) {
  int* dS;
  CK(cudaMalloc(&dS, 4));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Cfg { int box_cols, box_rows, boxes; bool swz; };
  const Cfg cfgs[] = {
      {256, 32, 2, false},  // 2 x 16 KB no-swizzle boxes (the fp32 kernel's shape, in fp16)
      {256, 64, 1, false},  // 1 x 32 KB
      {64, 32, 8, true},    // 8 x 4 KB  : 512 columns x 32 rows
      {64, 64, 8, true},    // 8 x 8 KB  : 512 columns x 64 rows (64 KB stage)
      {64, 64, 4, true},    // 4 x 8 KB  : 256 columns x 64 rows
      {64, 128, 4, true},   // 4 x 16 KB : 256 columns x 128 rows (64 KB stage)
      {64, 128, 2, true},   // 2 x 16 KB : 128 columns x 128 rows
      {64, 256, 2, true},   // 2 x 32 KB : 128 columns x 256 rows (64 KB stage)
  };
  int ci = -1;
  for (const Cfg& c : cfgs) {
    ++ci;
    if (only_cfg >= 0 && ci != only_cfg) continue;
    if (panel && !c.swz) continue;
    const int stage_bytes = c.box_cols * c.box_rows * 2 * c.boxes;
    int stages = (192 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    const int rows_pad = ((rows + 255) / 256) * 256;
    const int npanels = (int)((cols + 255) / 256);
    CUtensorMap map = panel ? make_map(X, 256, (uint64_t)npanels * rows_pad, 512, c.box_cols, c.box_rows, c.swz)
                            : make_map(X, (uint64_t)cols, (uint64_t)rows, (uint64_t)ld * 2, c.box_cols, c.box_rows, c.swz);
    const int tiles = (int)((cols + c.box_cols * c.boxes - 1) / (c.box_cols * c.boxes));
    for (int waves : {3, 4, 8}) {
      int splits = (148 * waves + tiles - 1) / tiles;
      if (waves == 3) splits = (148 * 4) / tiles > 1 ? (148 * 4) / tiles - 1 : 1;
      if (splits_override) splits = splits_override;
      if (splits < 1) splits = 1;
      int rps = (rows + splits - 1) / splits;
      rps = ((rps + c.box_rows - 1) / c.box_rows) * c.box_rows;
      splits = (rows + rps - 1) / rps;
      dim3 grid(tiles, splits);
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      CK(cudaMemset(dS, 0, 4));
      const size_t smem = (size_t)stages * stage_bytes + 1024;
      stream_kernel<<<grid, 64, smem>>>(map, rows, c.box_cols, c.box_rows, c.boxes, stages, rps, panel ? rows_pad : 0, dS);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      const int reps = 5;
      for (int r = 0; r < reps; ++r)
        stream_kernel<<<grid, 64, smem>>>(map, rows, c.box_cols, c.box_rows, c.boxes, stages, rps, panel ? rows_pad : 0, dS);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      int st = 0;
      CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
      const double bytes = (double)rows * (double)cols * 2.0;
      printf("%s%s box %3d cols x %3d rows x %d (%s, %3d KB stage x %d) grid %4d x %3d : %.3f ms  %.0f GB/s  status %d\n", what,
             panel ? " PANEL" : "      ", c.box_cols, c.box_rows, c.boxes, c.swz ? "SW128" : "none ", stage_bytes / 1024, stages, tiles, splits,
             ms / reps, bytes / (ms / reps * 1e-3) / 1e9, st);
    }
  }
}

// usage: f16_probe sem <case 0..4>   (one process per case: an illegal instruction poisons the context)
//        f16_probe stream
int main(int argc, char** argv) {
  if (argc >= 2 && argv[1][0] == 's' && argv[1][1] == 'e') {
    run_semantics(argc >= 3 ? atoi(argv[2]) : -1);
    return 0;
  }
  printf("== TMA streaming, fp16 ==\n");
  {
    const int rows = 20000;
    const int64_t cols = 100000, ld = 100000;
    __half* X;
    CK(cudaMalloc(&X, (size_t)rows * ld * 2));
    CK(cudaMemset(X, 0, (size_t)rows * ld * 2));
    if (argc < 3) run_stream(X, rows, cols, ld, "X  20000 x 100000");
    CK(cudaFree(X));
    const size_t pbytes = (size_t)((cols + 255) / 256) * (((rows + 255) / 256) * 256) * 512;
    CK(cudaMalloc(&X, pbytes));
    CK(cudaMemset(X, 0, pbytes));
    run_stream(X, rows, cols, ld, "X  20000 x 100000", true);
    CK(cudaFree(X));
  }
  {
    const int rows = 100000;
    const int64_t cols = 20000, ld = 20000;
    __half* X;
    CK(cudaMalloc(&X, (size_t)rows * ld * 2));
    CK(cudaMemset(X, 0, (size_t)rows * ld * 2));
    if (argc < 3) run_stream(X, rows, cols, ld, "XT 100000 x 20000");
    CK(cudaFree(X));
    const size_t pbytes = (size_t)((cols + 255) / 256) * (((rows + 255) / 256) * 256) * 512;
    CK(cudaMalloc(&X, pbytes));
    CK(cudaMemset(X, 0, pbytes));
    run_stream(X, rows, cols, ld, "XT 100000 x 20000", true);
    CK(cudaFree(X));
  }
  return 0;
}
