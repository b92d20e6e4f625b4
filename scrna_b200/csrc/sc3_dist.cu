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

extern "C" int scrna_pairwise_f64(const double* X, int64_t ld, int F, int N, int op, double* out, int64_t ldo,
                                  void* stream) {
  if (!X || !out || F <= 0 || N <= 0 || ld < N || ldo < N || (op != 0 && op != 1)) return SCRNA_ERR_BADARG;
  const int t = (int)ceil_div(N, PW_TILE);
  if (t > 65535) return SCRNA_ERR_UNSUPPORTED;
  dim3 grid(t, t);
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
