// SC3 transformations for sm_100a (scRNA/sc3_clustering_impl.py:135-203), fp64.
//
//   K11 pca_prepare   : dm - colmean, inf -> nan, / nanstd(col) unless every column std is zero (:161-173)
//   K12 laplacian     : the reference's "spectral" matrix with its broadcasting quirks (:143-150)
//   K13 jacobi_*      : right singular vectors of the n x n matrix by one-sided (Hestenes) Jacobi
//                       rotations in fp64 -- replaces scipy.linalg.svd (:178); only the right vectors are
//                       used by the caller (sc3_clustering.py:95-98), ordered by descending singular value.
#include "common.cuh"

namespace scrna {

// ---- K11 ----------------------------------------------------------------------------------------
// one thread per column; rows are visited in order so the sums follow numpy's axis-0 reduction order
__global__ void pca_center_kernel(const double* __restrict__ dm, int64_t ld, int n, double* __restrict__ out,
                                  int64_t ldo, double* __restrict__ sd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += dm[(int64_t)i * ld + j];
  const double mean = s / (double)n;
  // centred values, +-inf -> nan; nan-aware population std (np.nanstd: _nanvar)
  double cnt = 0.0, tot = 0.0;
  for (int i = 0; i < n; ++i) {
    double v = dm[(int64_t)i * ld + j] - mean;
    if (isinf(v)) v = nan("");
    out[(int64_t)i * ldo + j] = v;
    if (v == v) { tot += v; cnt += 1.0; }
  }
  const double avg = tot / cnt;
  double sq = 0.0;
  for (int i = 0; i < n; ++i) {
    const double v = out[(int64_t)i * ldo + j];
    if (v == v) { const double d = v - avg; sq += d * d; }
  }
  sd[j] = sqrt(sq / cnt);
}

__global__ void count_nonzero_kernel(const double* __restrict__ v, int n, int* __restrict__ out) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += (v[i] != 0.0);
  atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) out[0] = cnt;
}

__global__ void pca_scale_kernel(double* __restrict__ m, int64_t ld, int n, const double* __restrict__ sd,
                                 const int* __restrict__ nonzero) {
  if (nonzero[0] == 0) return;  // "All values are Zero!": the reference skips the division
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double s = sd[j];
  for (int i = blockIdx.y; i < n; i += gridDim.y) m[(int64_t)i * ld + j] /= s;
}

// ---- K12 ----------------------------------------------------------------------------------------
__global__ void row_sum_max_kernel(const double* __restrict__ dm, int64_t ld, int n, double* __restrict__ rowsum,
                                   double* __restrict__ blockmax) {
  // one warp per row: sum (fixed lane order + butterfly) and max
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double s = 0.0, mx = -INFINITY;
  if (row < n) {
    for (int j = lane; j < n; j += 32) {
      const double v = dm[(int64_t)row * ld + j];
      s += v;
      mx = fmax(mx, v);
    }
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (row < n && lane == 0) {
    rowsum[row] = s;
    blockmax[row] = mx;
  }
}

__global__ void max_reduce_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double sm[32];
  double mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) out[0] = mx;
  }
}

// out[i][j] = Dg[i]^-0.5 * ((Dg[j] - exp(-dm[i][j]/max)) * Dg[j]^-0.5)
__global__ void laplacian_kernel(const double* __restrict__ dm, int64_t ld, int n,
                                 const double* __restrict__ Dg, const double* __restrict__ mx,
                                 double* __restrict__ out, int64_t ldo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n || i >= n) return;
  const double a = exp(-dm[(int64_t)i * ld + j] / mx[0]);
  const double l = Dg[j] - a;
  const double di = pow(Dg[i], -0.5), dj = pow(Dg[j], -0.5);
  out[(int64_t)i * ldo + j] = di * (l * dj);
}

// ---- K13: one-sided Jacobi --------------------------------------------------------------------------
// AT holds the working matrix transposed (row j = column j of A = M.V), VT likewise for V.
// Round-robin pairing (circle method) over m = n rounded up to even "players"; player >= n is a bye.
__device__ __forceinline__ void rr_pair(int m, int round, int i, int& p, int& q) {
  const int mm = m - 1;
  if (i == 0) { p = round % mm; q = mm; }
  else { p = (round + i) % mm; q = (round - i + mm) % mm; }
  if (p > q) { const int t = p; p = q; q = t; }
}

__global__ void __launch_bounds__(256)
jacobi_round_kernel(double* __restrict__ AT, double* __restrict__ VT, int64_t ld, int n, int m, int round,
                    double tol, int* __restrict__ rotated) {
  int p, q;
  rr_pair(m, round, blockIdx.x, p, q);
  if (q >= n) return;  // bye
  __shared__ double red[3][8];
  __shared__ double cs[2];
  double* ap = AT + (int64_t)p * ld;
  double* aq = AT + (int64_t)q * ld;
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double x = ap[i], y = aq[i];
    alpha = fma(x, x, alpha);
    beta = fma(y, y, beta);
    gamma = fma(x, y, gamma);
  }
  alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = alpha; red[1][warp] = beta; red[2][warp] = gamma; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, g = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; g += red[2][w]; }
    double c = 1.0, s = 0.0;
    if (a > 0.0 && b > 0.0 && fabs(g) > tol * sqrt(a * b)) {
      const double zeta = (b - a) / (2.0 * g);
      const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      c = 1.0 / sqrt(1.0 + t * t);
      s = c * t;
      // rotations below 1e-13 rad still get applied but do not keep the sweep loop alive: near-null
      // columns (the centred matrix is rank deficient) otherwise rotate noise forever
      if (fabs(s) > 1e-13) atomicAdd(rotated, 1);
    }
    cs[0] = c; cs[1] = s;
  }
  __syncthreads();
  const double c = cs[0], s = cs[1];
  if (s == 0.0) return;
  double* vp = VT + (int64_t)p * ld;
  double* vq = VT + (int64_t)q * ld;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double x = ap[i], y = aq[i];
    ap[i] = c * x - s * y;
    aq[i] = s * x + c * y;
    const double u = vp[i], v = vq[i];
    vp[i] = c * u - s * v;
    vq[i] = s * u + c * v;
  }
}

__global__ void set_identity_kernel(double* __restrict__ V, int64_t ld, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < n && i < n) V[(int64_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
}

// squared norms of the rows of AT (= singular values squared)
__global__ void row_sqnorm_kernel(const double* __restrict__ AT, int64_t ld, int n, double* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  if (row < n)
    for (int j = lane; j < n; j += 32) { const double v = AT[(int64_t)row * ld + j]; s = fma(v, v, s); }
  s = warp_sum(s);
  if (row < n && lane == 0) out[row] = s;
}

// out[i][r] = VT[order[r]][i], r < c : the first c right singular vectors as columns (n x c, ldo)
__global__ void gather_vectors_kernel(const double* __restrict__ VT, int64_t ld, int n, const int* __restrict__ order,
                                      int c, double* __restrict__ out, int64_t ldo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i >= n || r >= c) return;
  out[(int64_t)i * ldo + r] = VT[(int64_t)order[r] * ld + i];
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_pca_prepare_f64(const double* dm, int64_t ld, int n, double* out, int64_t ldo, double* sd,
                                     int* scratch, void* stream) {
  if (!dm || !out || !sd || !scratch || n <= 0 || ld < n || ldo < n) return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  pca_center_kernel<<<(unsigned)ceil_div(n, 64), 64, 0, st>>>(dm, ld, n, out, ldo, sd);
  count_nonzero_kernel<<<1, 256, 0, st>>>(sd, n, scratch);
  int gy = (int)((int64_t)num_sms() * 8 / ceil_div(n, 128));
  if (gy < 1) gy = 1;
  if (gy > n) gy = n;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)gy);
  pca_scale_kernel<<<grid, 128, 0, st>>>(out, ldo, n, sd, scratch);
  return last_launch_status();
}

extern "C" int scrna_laplacian_f64(const double* dm, int64_t ld, int n, double* out, int64_t ldo, double* work,
                                   void* stream) {
  // work: 2*n + 1 doubles (row sums, row maxima, global max)
  if (!dm || !out || !work || n <= 0 || ld < n || ldo < n || n > 65535) return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  double* rowsum = work;
  double* rowmax = work + n;
  double* gmax = work + 2 * (int64_t)n;
  row_sum_max_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(dm, ld, n, rowsum, rowmax);
  max_reduce_kernel<<<1, 1024, 0, st>>>(rowmax, n, gmax);
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)n);
  laplacian_kernel<<<grid, 128, 0, st>>>(dm, ld, n, rowsum, gmax, out, ldo);
  return last_launch_status();
}

extern "C" int scrna_jacobi_init_f64(const double* M, int64_t ldm, int n, double* AT, double* VT, int64_t ld,
                                     void* stream) {
  // AT = M^T (rows of AT are the columns of M), VT = I
  if (!M || !AT || !VT || n <= 0 || ldm < n || ld < n || n > 65535) return SCRNA_ERR_BADARG;
  int rc = scrna_transpose_f64(M, ldm, n, n, AT, ld, stream);
  if (rc != SCRNA_OK) return rc;
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)n);
  set_identity_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(VT, ld, n);
  return last_launch_status();
}

extern "C" int scrna_jacobi_sweep_f64(double* AT, double* VT, int64_t ld, int n, double tol, int* rotated,
                                      void* stream) {
  // one full sweep = m-1 rounds of m/2 disjoint column pairs; `rotated` accumulates the rotation count
  if (!AT || !VT || !rotated || n <= 1 || ld < n) return SCRNA_ERR_BADARG;
  const int m = (n + 1) & ~1;
  cudaStream_t st = (cudaStream_t)stream;
  for (int round = 0; round < m - 1; ++round)
    jacobi_round_kernel<<<m / 2, 256, 0, st>>>(AT, VT, ld, n, m, round, tol, rotated);
  return last_launch_status();
}

extern "C" int scrna_row_sqnorm_f64(const double* AT, int64_t ld, int n, double* out, void* stream) {
  if (!AT || !out || n <= 0 || ld < n) return SCRNA_ERR_BADARG;
  row_sqnorm_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(AT, ld, n, out);
  return last_launch_status();
}

extern "C" int scrna_gather_vectors_f64(const double* VT, int64_t ld, int n, const int32_t* order, int c,
                                        double* out, int64_t ldo, void* stream) {
  if (!VT || !order || !out || n <= 0 || c <= 0 || c > n || ld < n || ldo < c || c > 65535) return SCRNA_ERR_BADARG;
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)c);
  gather_vectors_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(VT, ld, n, order, c, out, ldo);
  return last_launch_status();
}
