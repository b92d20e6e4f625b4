// K13 at the BASELINE size: the leading right singular vectors of the n x n SC3 matrix M as the leading
// eigenvectors of the Gram A = M^T M, by Householder tridiagonalisation + bisection + inverse iteration.
//
// Why a second solver.  The reference takes a full LAPACK SVD (scipy.linalg.svd, sc3_clustering_impl.py:178) and
// keeps the first ceil(0.07 n) right singular vectors (sc3_clustering.py:95-98).  The block one-sided Jacobi solver
// (sc3_svd_block.cu) is the accuracy reference of this package -- it works on M itself -- but needs 16-17 sweeps of
// 4 n^3 FMA: 46 s of the 62 s SC3 run at n = 20 000.  The classical dense route costs a fraction of ONE sweep:
//      A = M^T M                      n^3 FMA        (scrna_pairwise_f64, op 1)
//      A = Q T Q^T                    2/3 n^3 FMA    half of it a matrix-vector product per column: HBM-bound
//      T z = lambda z, top c          O(n c)         bisection on Sturm counts + inverse iteration
//      v = Q z                        2 n^2 c        the n - 2 reflectors applied to c vectors
// Squaring costs accuracy only where it does not matter here: eigenvector errors are eps |A| / gap, and the kept
// eigenvalues (sigma^2 from 2.4e6 down to ~9e2 at n = 20 000) have gaps ~1e0, i.e. 1e-10 -- two digits inside the
// 1e-8 the pipeline is held to.  Used for n >= 6000 (Sc3Ops.right_singular_vectors); smaller matrices, and with them
// every golden-vector parity test, stay on the Jacobi solver.
//
// Tridiagonalisation (LAPACK dsytrd / dlatrd, lower variant, restated for the device).  A is kept FULL and symmetric
// (both triangles), row-major, so "column c" is read as the contiguous row c.  Panels of 32 columns; per column c:
//   k1  x = A[c][c+1:] - (panel corrections)                                 -> unscaled reflector, partial |x|^2
//   k2  beta, tau, scale (one warp);  d[c], e[c], tau[c]
//   k3  y = A22 . v                  (the HBM-bound symmetric matrix-vector product over the trailing matrix; reads one
//                                     triangle by 128 x 128 tiles, eh_symv_upper_kernel)
//   k4  t1 = Wp^T v, t2 = Vp^T v     (panel corrections of y), v stored: Vp[i] and, for the back-transformation, A[c]
//   k5  w' = y - Vp^T t1 - Wp^T t2,  partial w'.v
//   k6  w = tau w' - (tau^2 (w'.v) / 2) v  -> Wp[i]
// and per panel the symmetric rank-64 update  A22 -= Vp^T Wp + Wp^T Vp  as ONE K = 64 GEMM on the FP64 tensor
// cores (DMMA).  All reductions are two-level with a fixed order: results are run-to-run identical.
#include "common.cuh"

namespace scrna {

constexpr int EH_NB = 32;          // panel width
constexpr int EH_RED = 256;        // CTAs of the vector kernels = partial sums per reduction

__device__ __forceinline__ void eh_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// scal: [0] scale, [1] tau, [2] alpha (x[c+1]), [3] w'.v
// ---- k1: unscaled reflector of column c = j + i ----------------------------------------------------------------
__global__ void __launch_bounds__(256)
eh_col_kernel(const double* __restrict__ A, int64_t ld, int n, int c, int i, const double* __restrict__ Vp,
              const double* __restrict__ Wp, double* __restrict__ x, double* __restrict__ part, double* __restrict__ d,
              double* __restrict__ scal) {
  __shared__ double vc[EH_NB], wc[EH_NB], scratch[32];
  if (threadIdx.x < i) {
    vc[threadIdx.x] = Vp[(int64_t)threadIdx.x * n + c];
    wc[threadIdx.x] = Wp[(int64_t)threadIdx.x * n + c];
  }
  __syncthreads();
  double ss = 0.0;
  for (int r = c + blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    double v = A[(int64_t)c * ld + r];
#pragma unroll 8
    for (int p = 0; p < i; ++p) v -= Vp[(int64_t)p * n + r] * wc[p] + Wp[(int64_t)p * n + r] * vc[p];
    if (r == c) d[c] = v;
    else {
      x[r] = v;
      if (r == c + 1) scal[2] = v;
      else ss += v * v;
    }
  }
  const double b = block_sum(ss, scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = b;
}

// ---- k2: Householder scalars (dlarfg) ----------------------------------------------------------------------------
__global__ void eh_house_kernel(const double* __restrict__ part, int nparts, int c, double* __restrict__ e,
                                double* __restrict__ tau, double* __restrict__ scal) {
  double s = 0.0;
  for (int p = threadIdx.x; p < nparts; p += 32) s += part[p];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const double alpha = scal[2];
    double beta = alpha, t = 0.0, sc = 0.0;
    if (s > 0.0) {
      beta = -copysign(sqrt(alpha * alpha + s), alpha);
      t = (beta - alpha) / beta;
      sc = 1.0 / (alpha - beta);
    }
    e[c] = beta;
    tau[c] = t;
    scal[0] = sc;
    scal[1] = t;
  }
}

// ---- k3: y = A22 . v over rows / columns (c, n): one warp per row -------------------------------------------------
__global__ void __launch_bounds__(256)
eh_symv_kernel(const double* __restrict__ A, int64_t ld, int n, int c, const double* __restrict__ x,
               const double* __restrict__ scal, double* __restrict__ y) {
  const int r = c + 1 + blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const double sc = scal[0];
  const double* row = A + (int64_t)r * ld;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int s = c + 1 + lane;
  if (s < n) {                                          // first chunk: v[c+1] = 1
    a0 = row[s] * (lane == 0 ? 1.0 : x[s] * sc);
    s += 32;
  }
  for (; s + 96 < n; s += 128) {
    a0 = fma(row[s], x[s] * sc, a0);
    a1 = fma(row[s + 32], x[s + 32] * sc, a1);
    a2 = fma(row[s + 64], x[s + 64] * sc, a2);
    a3 = fma(row[s + 96], x[s + 96] * sc, a3);
  }
  for (; s < n; s += 32) a0 = fma(row[s], x[s] * sc, a0);
  const double t = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) y[r] = t;
}

// ---- k3 (shipped): the same product reading ONE triangle.  A22 is symmetric, so every 128 x 128 tile (I, J) of the
// upper triangle (absolute 128-aligned blocks, J >= I) serves two pieces of y: rows of block I (A_IJ v_J) and, for
// J > I, rows of block J (A_IJ^T v_I) -- half the HBM traffic of the row-per-warp kernel above.  The pieces go to slabs
// (PR[J][r in block I], PC[I][s in block J]) that k5 adds up in a fixed order: deterministic, no atomics.
constexpr int EH_SB = 128;
__global__ void __launch_bounds__(256)
eh_symv_upper_kernel(const double* __restrict__ A, int64_t ld, int n, int c, const double* __restrict__ x,
                     const double* __restrict__ scal, double* __restrict__ PR, double* __restrict__ PC) {
  __shared__ double vI[EH_SB], vJ[EH_SB], cpart[8][EH_SB];
  const int lo = c + 1;
  const int b0 = lo / EH_SB, T = (n - 1) / EH_SB - b0 + 1;
  // linear tile index -> (I, J) of the upper triangle, row by row: row a holds T - a tiles
  const int t = blockIdx.x;
  int a = (int)(((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)t)) * 0.5);
  while (a > 0 && (int64_t)a * T - (int64_t)a * (a - 1) / 2 > t) --a;
  while ((int64_t)(a + 1) * T - (int64_t)(a + 1) * a / 2 <= t) ++a;
  const int I = b0 + a, J = I + (t - (int)((int64_t)a * T - (int64_t)a * (a - 1) / 2));
  const int I0 = I * EH_SB, J0 = J * EH_SB;
  const double sc = scal[0];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < EH_SB) {
    const int s = I0 + tid;
    vI[tid] = (s < lo || s >= n) ? 0.0 : (s == lo ? 1.0 : x[s] * sc);
  } else {
    const int s = J0 + tid - EH_SB;
    vJ[tid - EH_SB] = (s < lo || s >= n) ? 0.0 : (s == lo ? 1.0 : x[s] * sc);
  }
  __syncthreads();
  const bool offdiag = J > I;
  double cacc[4] = {0.0, 0.0, 0.0, 0.0};
  double vj[4];
  bool cok[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    vj[q] = vJ[lane + 32 * q];
    const int s = J0 + lane + 32 * q;
    cok[q] = s >= lo && s < n;
  }
  double* pr = PR + (int64_t)(J - b0) * n;
  for (int rb = 0; rb < 16; rb += 4) {                   // four rows = sixteen loads per thread in flight
    double av[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = I0 + warp * 16 + rb + u;
      const bool rok = r >= lo && r < n;                 // uniform over the warp
      const double* row = A + (int64_t)(rok ? r : lo) * ld + J0 + lane;
#pragma unroll
      for (int q = 0; q < 4; ++q) av[u][q] = (rok && cok[q]) ? row[32 * q] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rl = warp * 16 + rb + u, r = I0 + rl;
      const double vi = vI[rl];
      double rs = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        rs = fma(av[u][q], vj[q], rs);
        cacc[q] = fma(av[u][q], vi, cacc[q]);
      }
      rs = warp_sum(rs);
      if (lane == 0 && r >= lo && r < n) pr[r] = rs;
    }
  }
  if (offdiag) {
#pragma unroll
    for (int q = 0; q < 4; ++q) cpart[warp][lane + 32 * q] = cacc[q];
    __syncthreads();
    if (tid < EH_SB) {
      double s8 = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s8 += cpart[w][tid];    // fixed order
      const int sabs = J0 + tid;
      if (sabs < n) PC[(int64_t)(I - b0) * n + sabs] = s8;
    }
  }
}

// ---- k4: t1 = Wp . v, t2 = Vp . v (rows of the panel buffers against v); v is stored scaled -------------------------
__global__ void __launch_bounds__(256)
eh_panel_dots_kernel(int n, int c, int i, double* __restrict__ Vp, const double* __restrict__ Wp,
                     const double* __restrict__ x, const double* __restrict__ scal, double* __restrict__ Arow,
                     double* __restrict__ part) {
  __shared__ double sm[8][2 * EH_NB];
  const double sc = scal[0];
  double* vi = Vp + (int64_t)i * n;
  double acc1[EH_NB], acc2[EH_NB];
#pragma unroll
  for (int p = 0; p < EH_NB; ++p) acc1[p] = acc2[p] = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const double v = r <= c ? 0.0 : (r == c + 1 ? 1.0 : x[r] * sc);
    vi[r] = v;
    if (r > c) Arow[r] = v;   // reflector kept in row c of A for the back-transformation
#pragma unroll
    for (int p = 0; p < EH_NB; ++p)
      if (p < i) {
        acc1[p] = fma(Wp[(int64_t)p * n + r], v, acc1[p]);
        acc2[p] = fma(Vp[(int64_t)p * n + r], v, acc2[p]);
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < EH_NB; ++p)
    if (p < i) {
      const double a = warp_sum(acc1[p]);
      const double b = warp_sum(acc2[p]);
      if (lane == 0) { sm[warp][p] = a; sm[warp][EH_NB + p] = b; }
    }
  __syncthreads();
  if (threadIdx.x < 2 * EH_NB) {
    const int p = threadIdx.x % EH_NB;
    double t = 0.0;
    if (p < i)
      for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];   // fixed order
    part[(int64_t)blockIdx.x * 2 * EH_NB + threadIdx.x] = t;
  }
}

// ---- k5: w' = y - Vp^T t1 - Wp^T t2 ; partial w'.v ------------------------------------------------------------------
// Eight threads per row: thread `sub` of a row sums the slabs (and the panel corrections) with index = sub mod 8, the
// eight partial sums are combined by shuffles in a fixed order -- the ~150 slab reads of a row at n = 20 000 are
// spread over eight independent chains in eight threads instead of one latency-bound chain.
__global__ void __launch_bounds__(256)
eh_w_kernel(int n, int c, int i, const double* __restrict__ Vp, const double* __restrict__ Wp,
            const double* __restrict__ dots_part, int nparts, const double* __restrict__ y,
            const double* __restrict__ PR, const double* __restrict__ PC, double* __restrict__ w,
            double* __restrict__ part) {
  __shared__ double t1[EH_NB], t2[EH_NB], scratch[32];
  if (threadIdx.x < 2 * i) {
    const int p = threadIdx.x % i, which = threadIdx.x / i;
    double s = 0.0;
    for (int b = 0; b < nparts; ++b) s += dots_part[(int64_t)b * 2 * EH_NB + which * EH_NB + p];   // fixed order
    (which ? t2 : t1)[p] = s;
  }
  __syncthreads();
  const double* vi = Vp + (int64_t)i * n;
  double dot = 0.0;
  const int b0 = (c + 1) / EH_SB, B = (n - 1) / EH_SB;
  const int sub = threadIdx.x & 7, rows_per_cta = blockDim.x >> 3;
  for (int r0 = c + 1 + blockIdx.x * rows_per_cta; r0 < n; r0 += gridDim.x * rows_per_cta) {   // uniform per CTA
    const int r = r0 + (threadIdx.x >> 3);
    const bool ok = r < n;
    double v = 0.0;
    if (ok) {
      if (PR) {   // y[r] from the slabs of eh_symv_upper_kernel: tiles (I(r), J >= I(r)), then tiles (I < I(r), I(r))
        const int Ir = r / EH_SB;
        const double* pr = PR + (int64_t)(Ir - b0) * n + r;
        const int nr = B - Ir + 1;
        for (int q = sub; q < nr; q += 8) v += pr[(int64_t)q * n];
        const double* pc = PC + r;
        const int nc = Ir - b0;
        for (int q = sub; q < nc; q += 8) v += pc[(int64_t)q * n];
      } else if (sub == 0) {
        v = y[r];
      }
      for (int p = sub; p < i; p += 8) v -= Vp[(int64_t)p * n + r] * t1[p] + Wp[(int64_t)p * n + r] * t2[p];
    }
    // fixed-order combination of the eight partials of the row (lanes 8a .. 8a+7 of the warp)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (ok && sub == 0) {
      w[r] = v;
      dot = fma(v, vi[r], dot);
    }
  }
  const double b = block_sum(dot, scratch);
  if (threadIdx.x == 0) part[blockIdx.x] = b;
}

// ---- k6: w = tau w' - (tau^2 (w'.v) / 2) v -> Wp[i] -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
eh_wfin_kernel(int n, int c, int i, const double* __restrict__ Vp, double* __restrict__ Wp,
               const double* __restrict__ w, const double* __restrict__ part, int nparts,
               const double* __restrict__ scal) {
  __shared__ double sdot;
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int p = threadIdx.x; p < nparts; p += 32) s += part[p];
    s = warp_sum(s);
    if (threadIdx.x == 0) sdot = s;
  }
  __syncthreads();
  const double tau = scal[1];
  const double alpha2 = -0.5 * tau * tau * sdot;
  const double* vi = Vp + (int64_t)i * n;
  double* wi = Wp + (int64_t)i * n;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    wi[r] = r <= c ? 0.0 : tau * w[r] + alpha2 * vi[r];
}

// ---- trailing update: A[r][s] -= sum_p Vp[p][r] Wp[p][s] + Wp[p][r] Vp[p][s],  r, s >= lo --------------------------
// One CTA = 64 rows x 128 columns, K = 64 ([Vp | Wp] against [Wp | Vp]); 8 warps, each 32 x 32 as 4 x 4 DMMA tiles.
constexpr int EH_TP = 64 + 8;      // pitch of the staged 64 x 64 row operand   [k][r]
constexpr int EH_TQ = 128 + 8;     // pitch of the staged 64 x 128 column operand [k][s]
__global__ void __launch_bounds__(256)
eh_syr2k_kernel(double* __restrict__ A, int64_t ld, int n, int lo, int nb, const double* __restrict__ Vp,
                const double* __restrict__ Wp, int upper_only) {
  extern __shared__ __align__(16) double eh_smem[];
  double* P = eh_smem;                 // [k < 64][r < 64]
  double* Q = P + 64 * EH_TP;          // [k < 64][s < 128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 2, wc = warp & 3;
  const int lr = lane >> 2, lk = lane & 3;
  const int r0 = lo + blockIdx.y * 64, s0 = lo + blockIdx.x * 128;
  // only the upper triangle by 128-aligned blocks (columns in a block >= the row's block) is read again (k1 reads
  // row c to the right of the diagonal, eh_symv_upper_kernel tiles J >= I): tiles entirely left of it are skipped
  if (upper_only && s0 + 127 < (r0 / EH_SB) * EH_SB) return;
  // the tile of A goes straight into the accumulators (its loads fly while the operands are staged) and the column
  // operand is staged negated: acc = A - V^T W - W^T V comes out of the DMMA chain, the epilogue only stores
  double acc[4][4][2];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + 32 * wr + 8 * u + lr;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int s = s0 + 32 * wc + 8 * v + 2 * lk;
      const double* p = A + (int64_t)r * ld + s;
      acc[u][v][0] = (r < n && s < n) ? p[0] : 0.0;
      acc[u][v][1] = (r < n && s + 1 < n) ? p[1] : 0.0;
    }
  }
  for (int e = tid; e < 64 * 64; e += 256) {
    const int k = e >> 6, r = e & 63;
    const double* src = (k < 32 ? Vp + (int64_t)k * n : Wp + (int64_t)(k - 32) * n);
    P[k * EH_TP + r] = (k % 32 < nb && r0 + r < n) ? src[r0 + r] : 0.0;
  }
  for (int e = tid; e < 64 * 128; e += 256) {
    const int k = e >> 7, s = e & 127;
    const double* src = (k < 32 ? Wp + (int64_t)k * n : Vp + (int64_t)(k - 32) * n);
    Q[k * EH_TQ + s] = (k % 32 < nb && s0 + s < n) ? -src[s0 + s] : 0.0;
  }
  __syncthreads();
#pragma unroll 4
  for (int kk = 0; kk < 16; ++kk) {
    const int k = 4 * kk + lk;
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = P[k * EH_TP + 32 * wr + 8 * u + lr];
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = Q[k * EH_TQ + 32 * wc + 8 * u + lr];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) eh_dmma(acc[u][v], a[u], b[v]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + 32 * wr + 8 * u + lr;
    if (r >= n) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int s = s0 + 32 * wc + 8 * v + 2 * lk;
      double* p = A + (int64_t)r * ld + s;
      if (s < n) p[0] = acc[u][v][0];
      if (s + 1 < n) p[1] = acc[u][v][1];
    }
  }
}

// ---- eigenvalues: the `c` largest by bisection on Sturm counts (LAPACK dstebz idea).  One WARP per eigenvalue: the 32
// lanes count at 32 interior points of the bracket at once (a 33-section step narrows it 33x for the price of one
// n-term recurrence), then plain bisection finishes the last few ulps with the original stopping rule.
__device__ __forceinline__ int eh_sturm_count(const double* __restrict__ d, const double* __restrict__ e, int n, double x,
                                              double pivmin) {
  int cnt = 0;                                         // eigenvalues < x
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += q < 0.0;
  for (int k = 1; k < n; ++k) {
    const double ek = e[k - 1];
    q = d[k] - x - ek * ek / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;
}

__global__ void __launch_bounds__(128)
eh_bisect_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, int c, double gl,
                 double gu, double pivmin, double* __restrict__ lam) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // j-th largest, uniform over the warp
  if (j >= c) return;
  const int lane = threadIdx.x & 31;
  const int want = n - 1 - j;                            // ascending index of the wanted eigenvalue
  double lo = gl, hi = gu;                               // invariant: count(lo) <= want < count(hi)
  for (int it = 0; it < 40; ++it) {
    const double w = (hi - lo) / 33.0;
    const double x0 = lo + w, x31 = lo + 32.0 * w;
    if (!(w > 0.0) || x0 <= lo || x31 >= hi) break;       // the bracket is a few ulps wide: finish by bisection
    const double x = lo + (double)(lane + 1) * w;
    const int cnt = eh_sturm_count(d, e, n, x, pivmin);
    const unsigned below = __ballot_sync(0xffffffffu, cnt <= want);
    const int m = below == 0xffffffffu ? 32 : __ffs(~below) - 1;   // leading lanes whose point is still a lower bound
    const double nlo = m > 0 ? lo + (double)m * w : lo;
    const double nhi = m < 32 ? lo + (double)(m + 1) * w : hi;
    lo = nlo;
    hi = nhi;
  }
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    const int cnt = eh_sturm_count(d, e, n, mid, pivmin);   // every lane the same value
    if (cnt <= want) lo = mid; else hi = mid;
  }
  if (lane == 0) lam[j] = 0.5 * (lo + hi);
}

// ---- eigenvectors of T by inverse iteration (LAPACK dstein with dlagtf / dlagts), one thread per vector; the work
// arrays are [row][vector] so that the threads of a warp touch consecutive addresses ---------------------------------
__global__ void __launch_bounds__(128)
eh_invit_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, int c, const double* __restrict__ lam,
                double tnorm, double* __restrict__ wa, double* __restrict__ wb, double* __restrict__ wc,
                double* __restrict__ wd, int8_t* __restrict__ win, double* __restrict__ X, int iters) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  const double lm = lam[j];
  const double eps = 2.220446049250313e-16;
  const double tol = fmax(eps * tnorm, 1e-300);
  const int64_t C = c;
  // LU with partial pivoting of T - lambda I (dlagtf)
  double ak = d[0] - lm;                 // a[k], carried
  double scale1 = fabs(ak) + (n > 1 ? fabs(e[0]) : 0.0);
  double bk = n > 1 ? e[0] : 0.0;        // b[k] (super-diagonal), may be modified by the previous step
  for (int k = 0; k < n - 1; ++k) {
    double ak1 = d[k + 1] - lm;
    const double ck = e[k];
    const double bk1 = k < n - 2 ? e[k + 1] : 0.0;
    const double scale2 = fabs(ck) + fabs(ak1) + fabs(bk1);
    const double piv1 = ak == 0.0 ? 0.0 : fabs(ak) / scale1;
    double cout, dout = 0.0, bnext = bk1;
    int8_t in = 0;
    if (ck == 0.0) {
      scale1 = scale2;
      cout = 0.0;
    } else {
      const double piv2 = fabs(ck) / scale2;
      if (piv2 <= piv1) {
        scale1 = scale2;
        cout = ck / ak;
        ak1 -= cout * bk;
      } else {
        in = 1;
        const double mult = ak / ck;
        const double temp = ak1;
        wa[(int64_t)k * C + j] = ck;      // a[k] = c[k]
        ak1 = bk - mult * temp;
        dout = bk1;
        bnext = -mult * dout;
        bk = temp;                        // b[k] = temp
        cout = mult;
        wb[(int64_t)k * C + j] = bk;
        wc[(int64_t)k * C + j] = cout;
        wd[(int64_t)k * C + j] = dout;
        win[(int64_t)k * C + j] = in;
        ak = ak1;
        bk = bnext;
        continue;
      }
    }
    wa[(int64_t)k * C + j] = ak;
    wb[(int64_t)k * C + j] = bk;
    wc[(int64_t)k * C + j] = cout;
    wd[(int64_t)k * C + j] = dout;
    win[(int64_t)k * C + j] = in;
    ak = ak1;
    bk = bnext;
  }
  wa[(int64_t)(n - 1) * C + j] = ak;
  // start vector: deterministic pseudo-random in (-1, 1)
  for (int k = 0; k < n; ++k) {
    unsigned int h = (unsigned int)k * 2654435761u ^ ((unsigned int)j * 40503u + 0x9e3779b9u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    X[(int64_t)k * C + j] = (double)(h >> 8) * (2.0 / 16777216.0) - 1.0;
  }
  for (int it = 0; it < iters; ++it) {
    // forward substitution with the recorded interchanges (dlagts)
    double yprev = X[j];
    for (int k = 1; k < n; ++k) {
      double yk = X[(int64_t)k * C + j];
      const double cc = wc[(int64_t)(k - 1) * C + j];
      if (win[(int64_t)(k - 1) * C + j] == 0) {
        yk -= cc * yprev;
      } else {
        const double temp = yprev;
        yprev = yk;
        yk = temp - cc * yk;
      }
      X[(int64_t)(k - 1) * C + j] = yprev;
      yprev = yk;
    }
    X[(int64_t)(n - 1) * C + j] = yprev;
    // back substitution, tiny pivots perturbed to +-tol (dlagts job = -1)
    double y1 = 0.0, y2 = 0.0, nrm = 0.0;
    for (int k = n - 1; k >= 0; --k) {
      double t = X[(int64_t)k * C + j];
      if (k < n - 1) t -= wb[(int64_t)k * C + j] * y1;
      if (k < n - 2) t -= wd[(int64_t)k * C + j] * y2;
      double a = wa[(int64_t)k * C + j];
      if (fabs(a) < tol) a = copysign(tol, a == 0.0 ? 1.0 : a);
      t /= a;
      X[(int64_t)k * C + j] = t;
      y2 = y1;
      y1 = t;
      nrm = fmax(nrm, fabs(t));
    }
    // normalise (max norm between iterations, 2-norm at the end)
    double s2 = 0.0;
    const double inv = nrm > 0.0 ? 1.0 / nrm : 1.0;
    for (int k = 0; k < n; ++k) {
      const double t = X[(int64_t)k * C + j] * inv;
      X[(int64_t)k * C + j] = t;
      s2 = fma(t, t, s2);
    }
    if (it + 1 == iters) {
      const double inv2 = s2 > 0.0 ? rsqrt(s2) : 1.0;
      for (int k = 0; k < n; ++k) X[(int64_t)k * C + j] *= inv2;
    }
  }
}

// ---- back-transformation: the columns of Z (stored as rows of ZT, c x n) <- H_0 H_1 ... H_{n-3} z ------------------
// A CTA owns EH_BC vectors and applies ALL reflectors to them, last to first, in one launch: the vectors stream
// through L2 / HBM, the reflector (row of A) is shared by every CTA.
constexpr int EH_BC = 8;
__global__ void __launch_bounds__(256)
eh_backtransform_kernel(const double* __restrict__ A, int64_t ld, int n, const double* __restrict__ tau,
                        double* __restrict__ ZT, int c) {
  __shared__ double red[8][EH_BC];
  __shared__ double dots[EH_BC];
  const int v0 = blockIdx.x * EH_BC;
  const int nv = min(EH_BC, c - v0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int col = n - 3; col >= 0; --col) {
    const double t = tau[col];
    if (t == 0.0) continue;                       // uniform across the grid
    const double* v = A + (int64_t)col * ld;      // v[r], r in (col, n), v[col+1] = 1
    double acc[EH_BC];
#pragma unroll
    for (int u = 0; u < EH_BC; ++u) acc[u] = 0.0;
    for (int r = col + 1 + tid; r < n; r += 256) {
      const double vr = v[r];
#pragma unroll
      for (int u = 0; u < EH_BC; ++u)
        if (u < nv) acc[u] = fma(vr, ZT[(int64_t)(v0 + u) * n + r], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < EH_BC; ++u) {
      const double s = warp_sum(acc[u]);
      if (lane == 0) red[warp][u] = s;
    }
    __syncthreads();
    if (tid < EH_BC) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w][tid];
      dots[tid] = s * t;
    }
    __syncthreads();
    for (int r = col + 1 + tid; r < n; r += 256) {
      const double vr = v[r];
#pragma unroll
      for (int u = 0; u < EH_BC; ++u)
        if (u < nv) ZT[(int64_t)(v0 + u) * n + r] -= dots[u] * vr;
    }
    __syncthreads();
  }
}

// The same with ONE vector per CTA held in shared memory for the whole walk over the reflectors (n <= 20 480): the
// vector never leaves the SM, a reflector row is read once from L2 into registers for both of its uses, one CTA
// barrier per reflector (double-buffered partial sums).  5.7 s -> 0.5 s at n = 20 000, c = 1 401.
constexpr int EH_BT = 1024;
constexpr int EH_BT_MAXN = 20480;
constexpr int EH_BT_PER = (EH_BT_MAXN + EH_BT - 1) / EH_BT;   // reflector entries per thread
__global__ void __launch_bounds__(EH_BT, 1)
eh_backtransform_smem_kernel(const double* __restrict__ A, int64_t ld, int n, const double* __restrict__ tau,
                             double* __restrict__ ZT, int c) {
  extern __shared__ double z[];
  __shared__ double red[2][EH_BT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int vec = blockIdx.x; vec < c; vec += gridDim.x) {
    double* zg = ZT + (int64_t)vec * n;
    __syncthreads();
    for (int r = tid; r < n; r += EH_BT) z[r] = zg[r];
    __syncthreads();
    int buf = 0;
    for (int col = n - 3; col >= 0; --col) {
      const double t = tau[col];
      if (t == 0.0) continue;                       // uniform
      const double* v = A + (int64_t)col * ld;      // v[r], r in (col, n), v[col+1] = 1
      double vr[EH_BT_PER];
      double acc = 0.0;
#pragma unroll
      for (int u = 0; u < EH_BT_PER; ++u) {
        const int r = col + 1 + tid + u * EH_BT;
        vr[u] = r < n ? v[r] : 0.0;
        if (r < n) acc = fma(vr[u], z[r], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) red[buf][warp] = acc;
      __syncthreads();
      double dot = 0.0;
#pragma unroll
      for (int w = 0; w < EH_BT / 32; ++w) dot += red[buf][w];   // fixed order, every thread the same value
      dot *= t;
#pragma unroll
      for (int u = 0; u < EH_BT_PER; ++u) {
        const int r = col + 1 + tid + u * EH_BT;
        if (r < n) z[r] -= dot * vr[u];
      }
      buf ^= 1;   // the next reflector's partial sums go to the other buffer: no second barrier
    }
    __syncthreads();
    for (int r = tid; r < n; r += EH_BT) zg[r] = z[r];
  }
}

// ---- blocked back-transformation (compact WY, LAPACK dlarft / dlarfb restated) -------------------------------------
// 32 consecutive reflectors H_j ... H_{j+31} = I - V T V^T (T upper triangular from the Gram of the reflectors), so a
// block is applied to all vectors by two GEMMs on the FP64 tensor cores instead of 32 dot / axpy passes:
//      W1 = ZT . V        (vectors x 32, reduction over the rows: split into row chunks, partials added in order)
//      W2 = W1 . T^T
//      ZT -= W2 . V^T     (rank-32 update, the tile of ZT loaded straight into the DMMA accumulators)
// V[p][r] = A[j + p][r] for r > j + p (leading 1 stored), else 0.
constexpr int OB = 32;             // reflectors per block
constexpr int OB_RC = 2048;        // rows per split of the W1 reduction
constexpr int OB_KS = 16;          // rows per stage of the W1 kernel
constexpr int OB_SP = OB_KS + 4;   // its shared-memory pitch

__device__ __forceinline__ double ob_v(const double* __restrict__ A, int64_t ld, int j, int nb, int p, int r, int n) {
  return (p < nb && r > j + p && r < n) ? A[(int64_t)(j + p) * ld + r] : 0.0;
}

// T of every block: grid = blocks
__global__ void __launch_bounds__(256)
ob_t_kernel(const double* __restrict__ A, int64_t ld, int n, const double* __restrict__ tau, double* __restrict__ Tg) {
  __shared__ double sV[OB][65];
  __shared__ double S[OB][OB + 1];
  __shared__ double T[OB][OB + 1];
  const int j = blockIdx.x * OB, nref = n - 2, nb = min(OB, nref - j);
  const int tid = threadIdx.x;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r0 = j + 1; r0 < n; r0 += 64) {
    __syncthreads();
    for (int e = tid; e < OB * 64; e += 256) {
      const int p = e >> 6, rr = e & 63;
      sV[p][rr] = ob_v(A, ld, j, nb, p, r0 + rr, n);
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid * 4 + e, p = idx >> 5, i = idx & 31;
      double a = acc[e];
#pragma unroll 8
      for (int rr = 0; rr < 64; ++rr) a = fma(sV[p][rr], sV[i][rr], a);
      acc[e] = a;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int idx = tid * 4 + e;
    S[idx >> 5][idx & 31] = acc[e];
  }
  for (int e = tid; e < OB * (OB + 1); e += 256) (&T[0][0])[e] = 0.0;
  __syncthreads();
  if (tid < nb) {               // row a of T depends on its own earlier entries only
    const int a = tid;
    T[a][a] = tau[j + a];
    for (int i = a + 1; i < nb; ++i) {
      double sum = 0.0;
      for (int q = a; q < i; ++q) sum = fma(T[a][q], S[q][i], sum);
      T[a][i] = -tau[j + i] * sum;
    }
  }
  __syncthreads();
  for (int e = tid; e < OB * OB; e += 256) Tg[(int64_t)blockIdx.x * OB * OB + e] = T[e >> 5][e & 31];
}

// W1 partials: grid (ceil(c / 64), row chunks); 128 threads; out W1p[(chunk * cpad + vector) * 32 + p]
__global__ void __launch_bounds__(128)
ob_w1_kernel(const double* __restrict__ A, int64_t ld, int n, int j, int nb, int rbase, const double* __restrict__ ZT,
             int c, int cpad, double* __restrict__ W1p) {
  __shared__ __align__(16) double zs[2][64 * OB_SP];
  __shared__ __align__(16) double vs[2][OB * OB_SP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  const int v0 = blockIdx.x * 64;
  const int r_lo = rbase + blockIdx.y * OB_RC, r_hi = min(n, r_lo + OB_RC);
  const uint32_t zs_a = (uint32_t)__cvta_generic_to_shared(&zs[0][0]);
  const int nk = (r_hi - r_lo + OB_KS - 1) / OB_KS;
  auto issue = [&](int kc) {
    const int st = kc & 1, r0 = r_lo + kc * OB_KS;
    for (int e = tid; e < 64 * (OB_KS / 2); e += 128) {
      const int row = e >> 3, ch = e & 7, r = r0 + 2 * ch, v = v0 + row;
      const int valid = (v < c) ? max(0, min(2, r_hi - r)) : 0;
      const uint32_t dst = zs_a + (uint32_t)((st * 64 + row) * OB_SP + 2 * ch) * 8u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                   "l"(ZT + (int64_t)(valid ? v : 0) * n + (valid ? r : 0)), "r"((uint32_t)valid * 8u)
                   : "memory");
    }
    for (int e = tid; e < OB * OB_KS; e += 128) {
      const int p = e >> 4, rr = e & 15, r = r0 + rr;
      vs[st][p * OB_SP + rr] = r < r_hi ? ob_v(A, ld, j, nb, p, r, n) : 0.0;
    }
  };
  double acc[2][4][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[u][t][0] = acc[u][t][1] = 0.0;
  issue(0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (kc + 1 < nk) issue(kc + 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* xt = &zs[kc & 1][(warp * 16 + lr) * OB_SP + lk];
    const double* ct = &vs[kc & 1][lr * OB_SP + lk];
#pragma unroll
    for (int ks = 0; ks < OB_KS / 4; ++ks) {
      const double a0 = xt[4 * ks], a1 = xt[8 * OB_SP + 4 * ks];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double b = ct[t * 8 * OB_SP + 4 * ks];
        eh_dmma(acc[0][t], a0, b);
        eh_dmma(acc[1][t], a1, b);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int v = v0 + warp * 16 + u * 8 + lr;
    if (v >= c) continue;
    double* o = W1p + ((int64_t)blockIdx.y * cpad + v) * OB;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      o[t * 8 + 2 * lk] = acc[u][t][0];
      o[t * 8 + 2 * lk + 1] = acc[u][t][1];
    }
  }
}

// W2k[q][vector] = sum_p (sum_chunks W1p[chunk][vector][p]) T[q][p]; 8 vectors per CTA of 256 threads
__global__ void __launch_bounds__(256)
ob_w2_kernel(const double* __restrict__ W1p, int nchunks, int c, int cpad, const double* __restrict__ Tb,
             double* __restrict__ W2k) {
  __shared__ double w1[8][OB + 1];
  __shared__ double T[OB][OB + 1];
  const int tid = threadIdx.x, vl = tid >> 5, p = tid & 31;
  const int v = blockIdx.x * 8 + vl;
  for (int e = tid; e < OB * OB; e += 256) T[e >> 5][e & 31] = Tb[e];
  double s = 0.0;
  if (v < c)
    for (int ch = 0; ch < nchunks; ++ch) s += W1p[((int64_t)ch * cpad + v) * OB + p];   // fixed order
  w1[vl][p] = s;
  __syncthreads();
  const int q = p;
  double o = 0.0;
#pragma unroll 8
  for (int pp = 0; pp < OB; ++pp) o = fma(w1[vl][pp], T[q][pp], o);
  if (v < cpad) W2k[(int64_t)q * cpad + v] = v < c ? o : 0.0;
}

// ZT[vector][r] -= sum_q W2k[q][vector] V[q][r]: CTA = 64 vectors x 128 rows, K = 32
__global__ void __launch_bounds__(256)
ob_update_kernel(const double* __restrict__ A, int64_t ld, int n, int j, int nb, int rbase, double* __restrict__ ZT,
                 int c, int cpad, const double* __restrict__ W2k) {
  extern __shared__ __align__(16) double ob_smem[];
  double* P = ob_smem;                 // [q][vector < 64]
  double* Q = P + OB * EH_TP;          // [q][row < 128], negated
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 2, wc = warp & 3, lr = lane >> 2, lk = lane & 3;
  const int v0 = blockIdx.y * 64, r0 = rbase + blockIdx.x * 128;
  double acc[4][4][2];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int v = v0 + 32 * wr + 8 * u + lr;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int r = r0 + 32 * wc + 8 * w + 2 * lk;
      const double* zp = ZT + (int64_t)v * n + r;
      acc[u][w][0] = (v < c && r < n) ? zp[0] : 0.0;
      acc[u][w][1] = (v < c && r + 1 < n) ? zp[1] : 0.0;
    }
  }
  for (int e = tid; e < OB * 64; e += 256) {
    const int q = e >> 6, vv = e & 63;
    P[q * EH_TP + vv] = (v0 + vv < cpad) ? W2k[(int64_t)q * cpad + v0 + vv] : 0.0;
  }
  for (int e = tid; e < OB * 128; e += 256) {
    const int q = e >> 7, rr = e & 127;
    Q[q * EH_TQ + rr] = -ob_v(A, ld, j, nb, q, r0 + rr, n);
  }
  __syncthreads();
#pragma unroll 4
  for (int kk = 0; kk < OB / 4; ++kk) {
    const int k = 4 * kk + lk;
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = P[k * EH_TP + 32 * wr + 8 * u + lr];
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = Q[k * EH_TQ + 32 * wc + 8 * u + lr];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int w = 0; w < 4; ++w) eh_dmma(acc[u][w], a[u], b[w]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int v = v0 + 32 * wr + 8 * u + lr;
    if (v >= c) continue;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int r = r0 + 32 * wc + 8 * w + 2 * lk;
      double* zp = ZT + (int64_t)v * n + r;
      if (r < n) zp[0] = acc[u][w][0];
      if (r + 1 < n) zp[1] = acc[u][w][1];
    }
  }
}

}  // namespace scrna

using namespace scrna;

// work: 2 * 32 * n (Vp, Wp) + 3 * n (x, y, w) + 256 * 66 (partials) + 8 doubles + 2 * ceil(n / 128) * n (symv slabs)
extern "C" int64_t scrna_sytrd_work_elems(int n) {
  return n > 0 ? (int64_t)2 * EH_NB * n + 3 * (int64_t)n + (int64_t)EH_RED * (2 * EH_NB + 2) + 8 +
                     2 * ceil_div(n, EH_SB) * (int64_t)n   // slabs of the one-triangle symv
               : SCRNA_ERR_BADARG;
}

extern "C" int scrna_sytrd_f64(double* A, int64_t ld, int n, double* d, double* e, double* tau, double* work,
                               void* stream) {
  if (!A || !d || !e || !tau || !work || n < 3 || ld < n) return SCRNA_ERR_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* Vp = work;
  double* Wp = Vp + (int64_t)EH_NB * n;
  double* x = Wp + (int64_t)EH_NB * n;
  double* y = x + n;
  double* w = y + n;
  double* part = w + n;                                   // EH_RED
  double* dpart = part + EH_RED;                          // EH_RED * 64
  double* part2 = dpart + (int64_t)EH_RED * 2 * EH_NB;    // EH_RED
  double* scal = part2 + EH_RED;                          // 8
  double* PR = scal + 8;                                  // ceil(n/128) x n
  double* PC = PR + ceil_div(n, EH_SB) * (int64_t)n;      // ceil(n/128) x n
  static bool attr = false;
  const size_t smem_up = (size_t)(64 * EH_TP + 64 * EH_TQ) * sizeof(double);
  if (!attr) {
    if (cudaFuncSetAttribute(eh_syr2k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_up) != cudaSuccess)
      return SCRNA_ERR_CUDA;
    attr = true;
  }
  for (int j = 0; j < n - 2; j += EH_NB) {
    const int nb = (n - 2 - j) < EH_NB ? (n - 2 - j) : EH_NB;
    for (int i = 0; i < nb; ++i) {
      const int c = j + i;
      const int m = n - c;
      int blocks = (m + 127) / 128;            // 128-thread CTAs: twice the CTAs of these latency-bound row kernels
      if (blocks > EH_RED) blocks = EH_RED;
      eh_col_kernel<<<blocks, 128, 0, st>>>(A, ld, n, c, i, Vp, Wp, x, part, d, scal);
      eh_house_kernel<<<1, 32, 0, st>>>(part, blocks, c, e, tau, scal);
      {
        const int64_t T = (n - 1) / EH_SB - (c + 1) / EH_SB + 1;
        eh_symv_upper_kernel<<<(unsigned)(T * (T + 1) / 2), 256, 0, st>>>(A, ld, n, c, x, scal, PR, PC);
      }
      int vb = (n + 255) / 256;
      if (vb > EH_RED) vb = EH_RED;
      eh_panel_dots_kernel<<<vb, 256, 0, st>>>(n, c, i, Vp, Wp, x, scal, A + (int64_t)c * ld, dpart);
      int wblocks = (m + 31) / 32;             // 32 rows x 8 threads per CTA
      if (wblocks > EH_RED) wblocks = EH_RED;
      eh_w_kernel<<<wblocks, 256, 0, st>>>(n, c, i, Vp, Wp, dpart, vb, y, PR, PC, w, part2);
      eh_wfin_kernel<<<vb, 256, 0, st>>>(n, c, i, Vp, Wp, w, part2, wblocks, scal);
    }
    const int lo = j + nb;
    if (lo < n) {
      dim3 grid((unsigned)ceil_div(n - lo, 128), (unsigned)ceil_div(n - lo, 64));
      eh_syr2k_kernel<<<grid, 256, smem_up, st>>>(A, ld, n, lo, nb, Vp, Wp, 1);
    }
  }
  // the last 2 x 2 block: d[n-2], d[n-1], e[n-2] are what is left in A; no reflector
  // (A[n-2][n-2], A[n-2][n-1], A[n-1][n-1] carry every update because the trailing update covers them)
  cudaMemcpyAsync(d + n - 2, A + (int64_t)(n - 2) * ld + (n - 2), sizeof(double), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(e + n - 2, A + (int64_t)(n - 2) * ld + (n - 1), sizeof(double), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(d + n - 1, A + (int64_t)(n - 1) * ld + (n - 1), sizeof(double), cudaMemcpyDeviceToDevice, st);
  cudaMemsetAsync(tau + n - 2, 0, 2 * sizeof(double), st);
  return last_launch_status();
}

// work of the blocked back-transformation: T of every block + W1 partials + W2
extern "C" int64_t scrna_ormtr_work_elems(int n, int c) {
  if (n < 3 || c < 1) return SCRNA_ERR_BADARG;
  const int64_t nblocks = ceil_div(n - 2, OB), cpad = ceil_div(c, 64) * 64;
  const int64_t chunks = ceil_div(n, OB_RC) + 1;
  return nblocks * OB * OB + chunks * cpad * OB + (int64_t)OB * cpad;
}

// ZT (c x n, one vector per row, leading dimension n) <- Q z for every row, in blocks of 32 reflectors (compact WY, two
// DMMA GEMMs per block).  Needs n even and ZT 16-byte aligned (else SCRNA_ERR_UNSUPPORTED: use scrna_ormtr_f64).
extern "C" int scrna_ormtr_blocked_f64(const double* A, int64_t ld, int n, const double* tau, double* ZT, int c,
                                       double* work, void* stream) {
  if (!A || !tau || !ZT || !work || n < 3 || ld < n || c <= 0) return SCRNA_ERR_BADARG;
  if ((n & 1) || (reinterpret_cast<uintptr_t>(ZT) & 15)) return SCRNA_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nref = n - 2, nblocks = (int)ceil_div(nref, OB), cpad = (int)(ceil_div(c, 64) * 64);
  double* Tg = work;
  double* W1p = Tg + (int64_t)nblocks * OB * OB;
  double* W2k = W1p + (ceil_div(n, OB_RC) + 1) * (int64_t)cpad * OB;
  const size_t smem_u = (size_t)OB * (EH_TP + EH_TQ) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(ob_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u) != cudaSuccess)
      return SCRNA_ERR_CUDA;
    attr = true;
  }
  ob_t_kernel<<<nblocks, 256, 0, st>>>(A, ld, n, tau, Tg);
  for (int bk = nblocks - 1; bk >= 0; --bk) {
    const int j = bk * OB, nb = nref - j < OB ? nref - j : OB;
    const int rbase = ((j + 1) / 16) * 16;                 // first row that can be non-zero, aligned down to a stage
    const int nchunks = (int)ceil_div(n - rbase, OB_RC);
    ob_w1_kernel<<<dim3((unsigned)ceil_div(c, 64), (unsigned)nchunks), 128, 0, st>>>(A, ld, n, j, nb, rbase, ZT, c, cpad,
                                                                                    W1p);
    ob_w2_kernel<<<(unsigned)(cpad / 8), 256, 0, st>>>(W1p, nchunks, c, cpad, Tg + (int64_t)bk * OB * OB, W2k);
    ob_update_kernel<<<dim3((unsigned)ceil_div(n - rbase, 128), (unsigned)ceil_div(c, 64)), 256, smem_u, st>>>(
        A, ld, n, j, nb, rbase, ZT, c, cpad, W2k);
  }
  return last_launch_status();
}

extern "C" int scrna_stebz_top_f64(const double* d, const double* e, int n, int c, double gl, double gu, double pivmin,
                                   double* lam, void* stream) {
  if (!d || !e || !lam || n < 2 || c < 1 || c > n || !(gu > gl)) return SCRNA_ERR_BADARG;
  eh_bisect_kernel<<<(unsigned)ceil_div(c, 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d, e, n, c, gl, gu,
                                                                                              pivmin, lam);
  return last_launch_status();
}

// work: 4 * n * c doubles (a, b, c, d2) ; iwork: n * c bytes ; X: n x c (row-major, [row][vector])
extern "C" int scrna_stein_f64(const double* d, const double* e, int n, int c, const double* lam, double tnorm,
                               double* work, int8_t* iwork, double* X, int iters, void* stream) {
  if (!d || !e || !lam || !work || !iwork || !X || n < 2 || c < 1 || iters < 1) return SCRNA_ERR_BADARG;
  const int64_t nc = (int64_t)n * c;
  eh_invit_kernel<<<(unsigned)ceil_div(c, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d, e, n, c, lam, tnorm, work, work + nc, work + 2 * nc, work + 3 * nc, iwork, X, iters);
  return last_launch_status();
}

extern "C" int scrna_ormtr_f64(const double* A, int64_t ld, int n, const double* tau, double* ZT, int c, void* stream) {
  if (!A || !tau || !ZT || n < 3 || c < 1 || ld < n) return SCRNA_ERR_BADARG;
  if (n <= EH_BT_MAXN) {
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(eh_backtransform_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               EH_BT_MAXN * 8) != cudaSuccess)
        return SCRNA_ERR_CUDA;
      attr = true;
    }
    const int grid = c < num_sms() ? c : num_sms();
    eh_backtransform_smem_kernel<<<grid, EH_BT, (size_t)n * sizeof(double), reinterpret_cast<cudaStream_t>(stream)>>>(
        A, ld, n, tau, ZT, c);
    return last_launch_status();
  }
  eh_backtransform_kernel<<<(unsigned)ceil_div(c, EH_BC), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, ld, n, tau,
                                                                                                          ZT, c);
  return last_launch_status();
}
