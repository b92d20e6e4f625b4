// SC3 distances for sm_100a (scRNA/sc3_clustering_impl.py:106-132), all in fp64.
//
//   K9  pairwise_kernel<EUCLID> : D[i][j] = sqrt(sum_f (X[f][i]-X[f][j])^2), summed in gene order with
//                                 separate multiply and add (no FMA) -> bit-identical to
//                                 scipy.spatial.distance.pdist (plain sequential fp64 sum).
//       pairwise_kernel<DOT>    : G[i][j] = sum_f X[f][i]*X[f][j]  (covariance Gram for Pearson/Spearman,
//                                 also V.diag(s).V^T of the PCA stage)
//       The same kernel gives the consensus row distances (K19): the consensus matrix is symmetric,
//       so pdist over its rows equals the pairwise distance over its columns.
//   K10 rank_avg_kernel         : scipy.stats.rankdata(method='average') of every cell's gene vector
//       col_mean / center / corr_epilogue : np.corrcoef in numpy's op order (cov -> /std -> /std -> clip)
//
// B200's FP64 pipe is 1:2 of FP32 (ncu: "ratio of peak float to double is 2:1"), so the exact fp64
// path costs ~0.7 s at 20k cells x 20k genes; the 3xTF32 tcgen05 Gram is a later optimisation.
#include "common.cuh"

namespace scrna {

constexpr int PW_TILE = 64;   // output tile edge
constexpr int PW_FK = 16;     // features per smem stage

// row_begin < 0: the full symmetric N x N matrix (upper tiles computed, mirrored on write);
// row_begin >= 0: the row block [row_begin, row_end) x N of it, written to out[(i - row_begin)][j] (SC3
// distances sharded by row blocks over the GPUs) -- every element by the same instruction
// sequence as in the symmetric launch, so the block is bit-identical to the corresponding rows.
template <int OP>  // 0 = euclid, 1 = dot
__global__ void __launch_bounds__(256)
pairwise_kernel(const double* __restrict__ X, int64_t ld, int F, int N, double* __restrict__ out,
                int64_t ldo, int row_begin, int row_end) {
  const bool rect = row_begin >= 0;
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (!rect && bj < bi) return;  // symmetric: upper tiles only, mirrored on write
  __shared__ double sA[PW_FK][PW_TILE];
  __shared__ double sB[PW_FK][PW_TILE];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = (rect ? row_begin : 0) + bi * PW_TILE, j0 = bj * PW_TILE;
  const int i_end = rect ? row_end : N;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;

  for (int f0 = 0; f0 < F; f0 += PW_FK) {
    __syncthreads();
    for (int e = threadIdx.x; e < PW_FK * PW_TILE; e += 256) {
      const int ff = e / PW_TILE, c = e - ff * PW_TILE;
      const int f = f0 + ff;
      sA[ff][c] = (f < F && i0 + c < i_end) ? X[(int64_t)f * ld + i0 + c] : 0.0;
      sB[ff][c] = (f < F && j0 + c < N) ? X[(int64_t)f * ld + j0 + c] : 0.0;
    }
    __syncthreads();
    const int nf = min(PW_FK, F - f0);
    for (int ff = 0; ff < nf; ++ff) {   // strictly increasing f: the summation order of pdist
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = sA[ff][ty * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = sB[ff][tx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (OP == 0) {
            const double d = __dsub_rn(a[u], b[v]);
            acc[u][v] = __dadd_rn(acc[u][v], __dmul_rn(d, d));
          } else {
            acc[u][v] = fma(a[u], b[v], acc[u][v]);
          }
        }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < i_end && j < N) {
        const double r = (OP == 0) ? sqrt(acc[u][v]) : acc[u][v];
        if (rect) {
          out[(int64_t)(i - row_begin) * ldo + j] = r;
        } else {
          out[(int64_t)i * ldo + j] = r;
          if (bi != bj) out[(int64_t)j * ldo + i] = r;
        }
      }
    }
}

// ---- the Gram G = X^T X on the FP64 tensor cores ---------------------------------------------------------------------
// pairwise_kernel<DOT> is bound by its shared-memory reads (ncu: 87 % LSU wavefronts, FP64 pipe 34 %): 4 x 4 register
// tiles need 8 operand doubles per 16 FMA.  A DMMA fragment (mma.sync.m8n8k4.f64) needs 2 per 8 per lane.  CTA =
// 64 x 128 output tile (8 warps x 32 x 32 as 4 x 4 DMMA tiles), 32 features per stage, the [feature][cell] rows of
// X staged as they lie in memory with 16-byte cp.async (zero fill past N / F), two stages.  Every output element is
// the same chain over the features whatever tile it falls in, so a row block (multi-GPU sharding) is bit-identical to
// the same rows of the full launch.  Symmetric launches skip the tiles entirely below the diagonal; the lower
// triangle is then copied from the upper one (exactly symmetric result).
constexpr int GD_K = 32;
constexpr int GD_TP = 64 + 8;
constexpr int GD_TQ = 128 + 8;
constexpr int GD_STAGE = GD_K * (GD_TP + GD_TQ);   // doubles

__device__ __forceinline__ void gd_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void gd_cp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(256)
gram_dmma_kernel(const double* __restrict__ X, int64_t ld, int F, int N, double* __restrict__ out, int64_t ldo,
                 int row_begin, int row_end) {
  extern __shared__ __align__(16) double gd_smem[];
  const bool rect = row_begin >= 0;
  const int i0 = (rect ? row_begin : 0) + blockIdx.y * 64, j0 = blockIdx.x * 128;
  const int i_end = rect ? row_end : N;
  if (!rect && j0 + 127 < i0) return;        // entirely below the diagonal: mirrored afterwards
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 2, wc = warp & 3, lr = lane >> 2, lk = lane & 3;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(gd_smem);
  const int nk = (F + GD_K - 1) / GD_K;
  auto issue = [&](int kc) {
    const uint32_t pb = sbase + (uint32_t)((kc & 1) * GD_STAGE) * 8u, qb = pb + (uint32_t)(GD_K * GD_TP) * 8u;
    const int f0 = kc * GD_K;
    for (int e = tid; e < GD_K * 32; e += 256) {            // P: GD_K rows x 64 cells = 32 chunks of 2
      const int k = e >> 5, ch = e & 31, col = i0 + 2 * ch;
      const int valid = (f0 + k < F) ? max(0, min(2, N - col)) : 0;
      gd_cp16(pb + (uint32_t)(k * GD_TP + 2 * ch) * 8u, X + (int64_t)(valid ? f0 + k : 0) * ld + (valid ? col : 0),
              (uint32_t)valid * 8u);
    }
    for (int e = tid; e < GD_K * 64; e += 256) {            // Q: GD_K rows x 128 cells = 64 chunks of 2
      const int k = e >> 6, ch = e & 63, col = j0 + 2 * ch;
      const int valid = (f0 + k < F) ? max(0, min(2, N - col)) : 0;
      gd_cp16(qb + (uint32_t)(k * GD_TQ + 2 * ch) * 8u, X + (int64_t)(valid ? f0 + k : 0) * ld + (valid ? col : 0),
              (uint32_t)valid * 8u);
    }
  };
  double acc[4][4][2];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
  issue(0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // stage kc landed; stage kc - 1 fully consumed by every warp
    if (kc + 1 < nk) issue(kc + 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* P = gd_smem + (kc & 1) * GD_STAGE;
    const double* Q = P + GD_K * GD_TP;
#pragma unroll
    for (int kk = 0; kk < GD_K / 4; ++kk) {
      const int k = 4 * kk + lk;
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = P[k * GD_TP + 32 * wr + 8 * u + lr];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = Q[k * GD_TQ + 32 * wc + 8 * v + lr];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) gd_dmma(acc[u][v], a[u], b[v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = i0 + 32 * wr + 8 * u + lr;
    if (r >= i_end) continue;
    double* orow = out + (int64_t)(rect ? r - row_begin : r) * ldo;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int c = j0 + 32 * wc + 8 * v + 2 * lk;
      if (c < N) orow[c] = acc[u][v][0];
      if (c + 1 < N) orow[c + 1] = acc[u][v][1];
    }
  }
}

// out[i][j] = out[j][i] for i > j (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256)
mirror_upper_kernel(double* __restrict__ out, int64_t ldo, int N) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;                                  // destination tiles on or below the diagonal
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int y = ty; y < 32; y += 8) {                    // source tile (bj, bi): rows of block bj, columns of block bi
    const int r = bj * 32 + y, c = bi * 32 + tx;
    t[y][tx] = (r < N && c < N) ? out[(int64_t)r * ldo + c] : 0.0;
  }
  __syncthreads();
  for (int y = ty; y < 32; y += 8) {
    const int i = bi * 32 + y, j = bj * 32 + tx;
    if (i < N && j < N && i > j) out[(int64_t)i * ldo + j] = t[tx][y];
  }
}

// column means over the rows of an F x N matrix: mean[j] = (sum_f X[f][j]) / F  (fixed order per column)
__global__ void col_mean_kernel(const double* __restrict__ X, int64_t ld, int F, int N,
                                double* __restrict__ mean) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  // pairwise-free plain sum in blocks of 128 rows then across blocks: mirrors numpy's blocked
  // pairwise summation closely enough (differences ~1e-16 relative; Pearson is not bit-pinned)
  double total = 0.0;
  for (int f0 = 0; f0 < F; f0 += 128) {
    double s = 0.0;
    const int f1 = min(F, f0 + 128);
    for (int f = f0; f < f1; ++f) s += X[(int64_t)f * ld + j];
    total += s;
  }
  mean[j] = total / (double)F;
}

__global__ void center_cols_kernel(double* __restrict__ X, int64_t ld, int F, int N,
                                   const double* __restrict__ mean) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const double m = mean[j];
  for (int f = blockIdx.y; f < F; f += gridDim.y) X[(int64_t)f * ld + j] -= m;
}

// G (dot Gram of centred columns) -> 1 - corrcoef, numpy op order
__global__ void corr_epilogue_kernel(const double* __restrict__ G, int64_t ldg, int N, int F,
                                     double* __restrict__ out, int64_t ldo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= N || i >= N) return;
  const double fact = 1.0 / (double)(F - 1);
  const double cij = G[(int64_t)i * ldg + j] * fact;
  const double si = sqrt(G[(int64_t)i * ldg + i] * fact);
  const double sj = sqrt(G[(int64_t)j * ldg + j] * fact);
  double r = (cij / si) / sj;
  r = fmin(fmax(r, -1.0), 1.0);
  out[(int64_t)i * ldo + j] = 1.0 - r;
}

// K10: average ranks (1-based) of each ROW of Xt (N cells x F genes, row-major): one CTA per cell,
// rank_i = #{x < x_i} + (#{x == x_i} + 1) / 2.  The cell's vector is staged in shared memory in chunks of `chunk`
// genes (one chunk for up to 27 648 genes); the counts against successive chunks are accumulated in the output
// row itself -- half-integers, exact in fp64 -- so the gene count is not limited by shared memory.
__global__ void __launch_bounds__(1024)
rank_avg_kernel(const double* __restrict__ Xt, int64_t ld, int N, int F, int chunk, double* __restrict__ R,
                int64_t ldr) {
  extern __shared__ double sv[];
  const int cell = blockIdx.x;
  if (cell >= N) return;
  const double* row = Xt + (int64_t)cell * ld;
  double* out = R + (int64_t)cell * ldr;
  for (int c0 = 0; c0 < F; c0 += chunk) {
    const int nc = min(chunk, F - c0);
    __syncthreads();
    for (int f = threadIdx.x; f < nc; f += blockDim.x) sv[f] = row[c0 + f];
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      const double x = (f >= c0 && f < c0 + nc) ? sv[f - c0] : row[f];
      int less = 0, equal = 0;
      for (int m = 0; m < nc; ++m) {
        const double y = sv[m];
        less += (y < x);
        equal += (y == x);
      }
      const double part = (double)less + 0.5 * (double)equal;
      out[f] = (c0 == 0 ? 0.5 : out[f]) + part;
    }
  }
}

__global__ void __launch_bounds__(256)
transpose_f64_kernel(const double* __restrict__ src, int64_t ld_src, int64_t rows, int64_t cols,
                     double* __restrict__ dst, int64_t ld_dst) {
  __shared__ double tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;
    if (c < cols && r < rows) dst[c * ld_dst + r] = tile[tx][ty + i];
  }
}

}  // namespace scrna

using namespace scrna;

static constexpr size_t GD_SMEM = (size_t)2 * GD_STAGE * sizeof(double);
// 16-byte staging needs even leading dimension, an aligned base and an even first row of the block
static bool gram_dmma_ok(const double* X, int64_t ld, int row_begin) {
  return !(ld & 1) && !(reinterpret_cast<uintptr_t>(X) & 15) && !(row_begin & 1);
}
static int gram_dmma_configure() {
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(gram_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GD_SMEM) != cudaSuccess)
      return SCRNA_ERR_CUDA;
    done = true;
  }
  return SCRNA_OK;
}

extern "C" int scrna_pairwise_f64(const double* X, int64_t ld, int F, int N, int op, double* out, int64_t ldo,
                                  void* stream) {
  if (!X || !out || F <= 0 || N <= 0 || ld < N || ldo < N || (op != 0 && op != 1)) return SCRNA_ERR_BADARG;
  const int t = (int)ceil_div(N, PW_TILE);
  if (t > 65535) return SCRNA_ERR_UNSUPPORTED;
  dim3 grid(t, t);
  if (op == 1 && gram_dmma_ok(X, ld, 0)) {
    if (int rc = gram_dmma_configure()) return rc;
    const dim3 gg((unsigned)ceil_div(N, 128), (unsigned)ceil_div(N, 64));
    gram_dmma_kernel<<<gg, 256, GD_SMEM, (cudaStream_t)stream>>>(X, ld, F, N, out, ldo, -1, -1);
    const unsigned tm = (unsigned)ceil_div(N, 32);
    mirror_upper_kernel<<<dim3(tm, tm), 256, 0, (cudaStream_t)stream>>>(out, ldo, N);
    return last_launch_status();
  }
  if (op == 0)
    pairwise_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(X, ld, F, N, out, ldo, -1, -1);
  else
    pairwise_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(X, ld, F, N, out, ldo, -1, -1);
  return last_launch_status();
}

extern "C" int scrna_pairwise_rows_f64(const double* X, int64_t ld, int F, int N, int op, int row_begin, int row_end,
                                       double* out, int64_t ldo, void* stream) {
  if (!X || !out || F <= 0 || N <= 0 || ld < N || ldo < N || (op != 0 && op != 1) || row_begin < 0 ||
      row_end > N || row_begin > row_end)
    return SCRNA_ERR_BADARG;
  if (row_begin == row_end) return SCRNA_OK;
  const int tx = (int)ceil_div(N, PW_TILE), ty = (int)ceil_div(row_end - row_begin, PW_TILE);
  if (tx > 65535 || ty > 65535) return SCRNA_ERR_UNSUPPORTED;
  dim3 grid(tx, ty);
  if (op == 1 && gram_dmma_ok(X, ld, row_begin)) {
    if (int rc = gram_dmma_configure()) return rc;
    const dim3 gg((unsigned)ceil_div(N, 128), (unsigned)ceil_div(row_end - row_begin, 64));
    gram_dmma_kernel<<<gg, 256, GD_SMEM, (cudaStream_t)stream>>>(X, ld, F, N, out, ldo, row_begin, row_end);
    return last_launch_status();
  }
  if (op == 0)
    pairwise_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(X, ld, F, N, out, ldo, row_begin, row_end);
  else
    pairwise_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(X, ld, F, N, out, ldo, row_begin, row_end);
  return last_launch_status();
}

extern "C" int scrna_col_center_f64(double* X, int64_t ld, int F, int N, double* mean, void* stream) {
  if (!X || !mean || F <= 0 || N <= 0 || ld < N) return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  col_mean_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, st>>>(X, ld, F, N, mean);
  int gy = (int)((int64_t)num_sms() * 16 / ceil_div(N, 128));
  if (gy < 1) gy = 1;
  if (gy > F) gy = F;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)gy);
  center_cols_kernel<<<grid, 128, 0, st>>>(X, ld, F, N, mean);
  return last_launch_status();
}

extern "C" int scrna_corr_epilogue_f64(const double* G, int64_t ldg, int N, int F, double* out, int64_t ldo,
                                       void* stream) {
  if (!G || !out || N <= 0 || F <= 1 || ldg < N || ldo < N || N > 65535) return SCRNA_ERR_BADARG;
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)N);
  corr_epilogue_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(G, ldg, N, F, out, ldo);
  return last_launch_status();
}

extern "C" int scrna_rank_avg_f64(const double* Xt, int64_t ld, int N, int F, double* R, int64_t ldr,
                                  void* stream) {
  if (!Xt || !R || N <= 0 || F <= 0 || ld < F || ldr < F) return SCRNA_ERR_BADARG;
  const int chunk = F < 27648 ? F : 27648;
  const size_t smem = (size_t)chunk * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(rank_avg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr = true;
  }
  int threads = F >= 1024 ? 1024 : (int)(ceil_div(F, 32) * 32);
  rank_avg_kernel<<<N, threads, smem, (cudaStream_t)stream>>>(Xt, ld, N, F, chunk, R, ldr);
  return last_launch_status();
}

extern "C" int scrna_transpose_f64(const double* src, int64_t ld_src, int64_t rows, int64_t cols, double* dst,
                                   int64_t ld_dst, void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || cols > ld_src || rows > ld_dst) return SCRNA_ERR_BADARG;
  const int64_t by = ceil_div(rows, 32);
  if (by > 65535) return SCRNA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)by);
  transpose_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, dst, ld_dst);
  return last_launch_status();
}
