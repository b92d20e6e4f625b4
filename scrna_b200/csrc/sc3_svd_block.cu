// K13 at scale: right singular vectors of the n x n SC3 matrix by BLOCK one-sided Jacobi in fp64.
//
// The reference takes a full LAPACK SVD (scipy.linalg.svd, sc3_clustering_impl.py:178) and keeps the first
// `components` right singular vectors (sc3_clustering.py:95-98).  The spectrum below the few cluster
// directions is a flat noise bulk (neighbouring singular values 0.1 % apart at rank 0.07 n), which no
// subspace / Krylov iteration resolves; a dense O(n^3) method is needed, and it has to be GEMM-shaped to run
// at the FP64 pipe's rate instead of HBM's (the unblocked kernel in sc3_transform.cu streams the whole
// matrix once per rotation round: 32 n^3 bytes per sweep).
//
// Formulation.  One-sided Jacobi is applied to N = M^T: rotations R orthogonalise the columns of N, and at
// convergence N.R = V.Sigma where V are the LEFT singular vectors of N = the RIGHT singular vectors of M.
// Column j of N is row j of M, so the working array W (row j = column j of N.R) starts as M itself, no
// transposition and no accumulated rotation matrix: v_j = W[j] / |W[j]|, sigma_j = |W[j]|.
//
// Blocking.  Columns are grouped in blocks of 32; a round pairs the blocks (round-robin tournament) and for
// every pair (I, J), independently:
//   1. gram   : G = [W_I; W_J] . [W_I; W_J]^T                 (64 x 64, reduction over n, split partials; DMMA)
//   2. eigen  : two-sided cyclic Jacobi on G in shared memory  -> U (64 x 64 orthogonal), G.U = U.diag
//   3. apply  : [W_I; W_J] <- U^T . [W_I; W_J]                  (64 x 64 times 64 x n; DMMA)
// A sweep is nblk-1 rounds; traffic per sweep drops from 32 n^3 to n^3 bytes and the work is two GEMMs.
#include "common.cuh"

namespace scrna {

constexpr int BJ_B = 32;          // columns per block
constexpr int BJ_P = 2 * BJ_B;    // columns per pair
constexpr int BJ_LD = BJ_P + 1;   // shared-memory pitch (doubles)

__device__ __forceinline__ void bj_pair(int m, int round, int i, int& p, int& q) {
  const int mm = m - 1;
  if (i == 0) { p = round % mm; q = mm; }
  else { p = (round + i) % mm; q = (round - i + mm) % mm; }
  if (p > q) { const int t = p; p = q; q = t; }
}

// FP64 tensor-core MMA (DMMA): D(8x8) += A(8x4, row) . B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4] and
// D[l/4][2(l%4) .. +1].  Same peak as the DFMA pipe on B200 (18.5 vs 17 TFMA/s measured,
// profiles/r01_dmma_rate_microbench.txt) but 8 FMA per lane for two operand doubles: the 4x4 register-tile
// DFMA kernels were shared-memory bound (ncu: 70 % LSU wavefronts, fp64 pipe 28 %).
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ---- 1. pair Gram partials: part[pair][split][64*64] ------------------------------------------------
// 64 rows x 32 k staged row-major (pitch 36: fragment loads are 2-way = optimal for 64-bit, staging stores
// conflict-free); 8 warps, each a 16 x 32 slice of the 64 x 64 Gram as 2 x 4 DMMA tiles.
constexpr int BJ_KC = 32;          // reduction chunk of the Gram kernel
constexpr int BJ_GP = BJ_KC + 4;   // pitch of the staged chunk
__global__ void __launch_bounds__(256)
bj_gram_kernel(const double* __restrict__ W, int64_t ld, int n, int nblk, int m, int round, int nsplit,
               double* __restrict__ part) {
  int p, q;
  bj_pair(m, round, blockIdx.x, p, q);
  if (q >= nblk) return;  // bye
  __shared__ __align__(16) double tile[2 * BJ_P * BJ_GP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 1, wc = warp & 1;        // rows 16 wr .., columns 32 wc ..
  const int lr = lane >> 2, lk = lane & 3;
  int per = (int)(((int64_t)n + nsplit - 1) / nsplit);
  per += per & 1;   // even spans: 16-byte staging units never straddle two splits
  const int c_begin = blockIdx.y * per;
  const int c_end = min(n, c_begin + per);
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // double-buffered staging: 16-byte cp.async when the rows allow it (even n, ld, span start), else plain loads
  const bool vec = ((n | (int)(ld & 1) | c_begin) & 1) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  auto stage = [&](int c0, double* T) {
    if (vec) {
      const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(T);
      for (int e = tid; e < BJ_P * (BJ_KC / 2); e += 256) {
        const int r = e / (BJ_KC / 2), u = e % (BJ_KC / 2);
        const int c = c0 + 2 * u;
        const int64_t row = (r < BJ_B) ? (int64_t)p * BJ_B + r : (int64_t)q * BJ_B + (r - BJ_B);
        const bool ok = c < c_end;   // n and the spans are even: a unit is wholly inside or outside
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tbase + (uint32_t)(r * BJ_GP + 2 * u) * 8u),
                     "l"(W + row * ld + (ok ? c : 0)), "r"(ok ? 16u : 0u)
                     : "memory");
      }
    } else {
      for (int e = tid; e < BJ_P * BJ_KC; e += 256) {
        const int r = e / BJ_KC, c = e % BJ_KC;
        const int64_t row = (r < BJ_B) ? (int64_t)p * BJ_B + r : (int64_t)q * BJ_B + (r - BJ_B);
        T[r * BJ_GP + c] = (c0 + c < c_end) ? W[row * ld + c0 + c] : 0.0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if (c_begin < c_end) stage(c_begin, tile);
  for (int c0 = c_begin; c0 < c_end; c0 += BJ_KC) {
    const double* T = tile + buf * (BJ_P * BJ_GP);
    if (c0 + BJ_KC < c_end) {
      stage(c0 + BJ_KC, tile + (buf ^ 1) * (BJ_P * BJ_GP));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BJ_KC / 4; ++kk) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = T[(16 * wr + 8 * i + lr) * BJ_GP + 4 * kk + lk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = T[(32 * wc + 8 * j + lr) * BJ_GP + 4 * kk + lk];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    __syncthreads();  // done with this buffer before the next iteration refills it
    buf ^= 1;
  }
  double* out = part + ((int64_t)blockIdx.x * nsplit + blockIdx.y) * (BJ_P * BJ_P);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<double2*>(&out[(16 * wr + 8 * i + lr) * BJ_P + 32 * wc + 8 * j + 2 * lk]) =
          make_double2(acc[i][j][0], acc[i][j][1]);
}

// ---- 2. eigenvectors of the 64 x 64 pair Gram: U[pair][64*64] (row-major, G.U = U.diag) ---------------
constexpr int BJ_ET = 512;   // threads of the eigen kernel: one CTA per pair, three pairs resident per SM
__global__ void __launch_bounds__(BJ_ET)
bj_eigen_kernel(const double* __restrict__ part, int nsplit, int nblk, int m, int round, double tol,
                double null_sq, double* __restrict__ Uout, int* __restrict__ rotated) {
  int p, q;
  bj_pair(m, round, blockIdx.x, p, q);
  if (q >= nblk) return;
  extern __shared__ double bj_smem[];
  double* G = bj_smem;                  // 64 x 65
  double* U = G + BJ_P * BJ_LD;         // 64 x 65
  __shared__ double cs[BJ_B][2];
  __shared__ int pq[BJ_B][2];
  __shared__ int any_rot, big_rot;
  const int tid = threadIdx.x;
  const double* src = part + (int64_t)blockIdx.x * nsplit * (BJ_P * BJ_P);
  for (int e = tid; e < BJ_P * BJ_P; e += BJ_ET) {
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += src[(int64_t)sp * (BJ_P * BJ_P) + e];  // fixed order
    const int r = e / BJ_P, c = e % BJ_P;
    G[r * BJ_LD + c] = s;
    U[r * BJ_LD + c] = (r == c) ? 1.0 : 0.0;
  }
  if (tid == 0) big_rot = 0;
  __syncthreads();
  // symmetrise (the two triangles were summed in different orders)
  for (int e = tid; e < BJ_P * BJ_P; e += BJ_ET) {
    const int r = e / BJ_P, c = e % BJ_P;
    if (r < c) {
      const double v = 0.5 * (G[r * BJ_LD + c] + G[c * BJ_LD + r]);
      G[r * BJ_LD + c] = v;
      G[c * BJ_LD + r] = v;
    }
  }
  __syncthreads();
  for (int sweep = 0; sweep < 30; ++sweep) {
    if (tid == 0) any_rot = 0;
    __syncthreads();
    for (int step = 0; step < BJ_P - 1; ++step) {
      if (tid < BJ_B) {
        int a, b;
        bj_pair(BJ_P, step, tid, a, b);
        const double gaa = G[a * BJ_LD + a], gbb = G[b * BJ_LD + b], gab = G[a * BJ_LD + b];
        double c = 1.0, s = 0.0;
        // columns below the numerical-null threshold are left alone: they are noise and would rotate forever
        if (gaa > null_sq && gbb > null_sq && fabs(gab) > tol * sqrt(gaa * gbb)) {
          const double zeta = (gbb - gaa) / (2.0 * gab);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          c = 1.0 / sqrt(1.0 + t * t);
          s = c * t;
          any_rot = 1;
          // Only couplings above 1e-8 ask for another sweep: Jacobi converges quadratically, so once every
          // coupling of a sweep is below 1e-8 the rotations just applied leave O(1e-16 / gap) behind -- the
          // confirmation sweep that would merely observe this is skipped.
          if (fabs(gab) > 1e-8 * sqrt(gaa * gbb)) big_rot = 1;
        }
        cs[tid][0] = c; cs[tid][1] = s;
        pq[tid][0] = a; pq[tid][1] = b;
      }
      __syncthreads();
      // rows: G <- J^T G   (row a' = c.a - s.b ; row b' = s.a + c.b)
      for (int e = tid; e < BJ_B * BJ_P; e += BJ_ET) {
        const int pr = e / BJ_P, col = e % BJ_P;
        const double c = cs[pr][0], s = cs[pr][1];
        if (s != 0.0) {
          const int a = pq[pr][0], b = pq[pr][1];
          const double x = G[a * BJ_LD + col], y = G[b * BJ_LD + col];
          G[a * BJ_LD + col] = c * x - s * y;
          G[b * BJ_LD + col] = s * x + c * y;
        }
      }
      __syncthreads();
      // columns: G <- G J, U <- U J
      for (int e = tid; e < BJ_B * BJ_P; e += BJ_ET) {
        const int pr = e / BJ_P, row = e % BJ_P;
        const double c = cs[pr][0], s = cs[pr][1];
        if (s != 0.0) {
          const int a = pq[pr][0], b = pq[pr][1];
          double x = G[row * BJ_LD + a], y = G[row * BJ_LD + b];
          G[row * BJ_LD + a] = c * x - s * y;
          G[row * BJ_LD + b] = s * x + c * y;
          x = U[row * BJ_LD + a]; y = U[row * BJ_LD + b];
          U[row * BJ_LD + a] = c * x - s * y;
          U[row * BJ_LD + b] = s * x + c * y;
        }
      }
      __syncthreads();
    }
    if (!any_rot) break;
    __syncthreads();
  }
  double* dst = Uout + (int64_t)blockIdx.x * (BJ_P * BJ_P);
  for (int e = tid; e < BJ_P * BJ_P; e += BJ_ET) dst[e] = U[(e / BJ_P) * BJ_LD + (e % BJ_P)];
  if (tid == 0 && big_rot) atomicAdd(rotated, 1);
}

// ---- 3. apply: rows [I; J] of W  <-  U^T . rows -----------------------------------------------------
// One CTA = 128 columns of the pair's 64 rows; 8 warps, each a 32 x 32 output slice as 4 x 4 DMMA tiles
// (8 operand loads per 16 DMMA).  U is staged as [s][r] (pitch 72), the rows as [s][col] (pitch 136).  (A
// 64-column variant with three CTAs per SM measured slower: 63.6 vs 55.7 s at n = 20 000.)
constexpr int BJ_AC = 128;
constexpr int BJ_UP = BJ_P + 8;
constexpr int BJ_AP = BJ_AC + 8;
__global__ void __launch_bounds__(256)
bj_apply_kernel(double* __restrict__ W, int64_t ld, int n, int nblk, int m, int round,
                const double* __restrict__ Uall) {
  int p, q;
  bj_pair(m, round, blockIdx.x, p, q);
  if (q >= nblk) return;
  extern __shared__ __align__(16) double bj_smem[];
  double* Us = bj_smem;                 // [s][r]   64 x 72
  double* T = Us + BJ_P * BJ_UP;        // [s][col] 64 x 136
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 2, wc = warp & 3;        // rows 32 wr .., columns 32 wc ..
  const int lr = lane >> 2, lk = lane & 3;
  const int c0 = blockIdx.y * BJ_AC;
  const double* Ug = Uall + (int64_t)blockIdx.x * (BJ_P * BJ_P);
  for (int e = tid; e < BJ_P * BJ_P; e += 256) Us[(e / BJ_P) * BJ_UP + (e % BJ_P)] = Ug[e];
  for (int e = tid; e < BJ_P * BJ_AC; e += 256) {
    const int s = e / BJ_AC, c = e % BJ_AC;
    const int64_t row = (s < BJ_B) ? (int64_t)p * BJ_B + s : (int64_t)q * BJ_B + (s - BJ_B);
    T[s * BJ_AP + c] = (c0 + c < n) ? W[row * ld + c0 + c] : 0.0;
  }
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < BJ_P / 4; ++kk) {
    const int s = 4 * kk + lk;
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = Us[s * BJ_UP + 32 * wr + 8 * i + lr];   // A[r][s] = U[s][r]
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = T[s * BJ_AP + 32 * wc + 8 * j + lr];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 32 * wr + 8 * i + lr;
    const int64_t row = (r < BJ_B) ? (int64_t)p * BJ_B + r : (int64_t)q * BJ_B + (r - BJ_B);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 32 * wc + 8 * j + 2 * lk;
      double* wp = W + row * ld + c;
      if (c + 1 < n && ((reinterpret_cast<uintptr_t>(wp) & 15) == 0)) {
        *reinterpret_cast<double2*>(wp) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (c < n) wp[0] = acc[i][j][0];
        if (c + 1 < n) wp[1] = acc[i][j][1];
      }
    }
  }
}

// ---- 3b. apply, pipelined: one CTA walks several 128-column chunks of its pair ---------------------------
// U is staged once per CTA and the row chunks are double-buffered with cp.async (the next chunk lands while the
// DMMAs of the current one run); used when the rows are 16-byte aligned (even n and ld).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(256, 1)
bj_apply_pipe_kernel(double* __restrict__ W, int64_t ld, int n, int nblk, int m, int round,
                     const double* __restrict__ Uall, int chunks_per_cta) {
  int p, q;
  bj_pair(m, round, blockIdx.x, p, q);
  if (q >= nblk) return;
  extern __shared__ __align__(16) double bj_smem[];
  double* Us = bj_smem;                       // [s][r]   64 x 72
  double* Tb[2] = {Us + BJ_P * BJ_UP, Us + BJ_P * BJ_UP + BJ_P * BJ_AP};   // 2 x [s][col] 64 x 136
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 2, wc = warp & 3;
  const int lr = lane >> 2, lk = lane & 3;
  const int nchunks = (n + BJ_AC - 1) / BJ_AC;
  const int first = blockIdx.y * chunks_per_cta;
  const int last = min(nchunks, first + chunks_per_cta);
  if (first >= last) return;
  const double* Ug = Uall + (int64_t)blockIdx.x * (BJ_P * BJ_P);
  for (int e = tid; e < BJ_P * BJ_P; e += 256) Us[(e / BJ_P) * BJ_UP + (e % BJ_P)] = Ug[e];

  auto issue = [&](int chunk, double* T) {
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(T);
    for (int e = tid; e < BJ_P * (BJ_AC / 2); e += 256) {
      const int s = e / (BJ_AC / 2), u = e % (BJ_AC / 2);
      const int c = chunk * BJ_AC + 2 * u;
      const int64_t row = (s < BJ_B) ? (int64_t)p * BJ_B + s : (int64_t)q * BJ_B + (s - BJ_B);
      const bool ok = c < n;  // n is even: a 16-byte unit is either wholly inside or wholly outside
      cp_async16(tbase + (uint32_t)(s * BJ_AP + 2 * u) * 8u, W + row * ld + (ok ? c : 0), ok ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue(first, Tb[0]);
  for (int ch = first; ch < last; ++ch) {
    const double* T = Tb[(ch - first) & 1];
    if (ch + 1 < last) {
      issue(ch + 1, Tb[(ch - first + 1) & 1]);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < BJ_P / 4; ++kk) {
      const int s = 4 * kk + lk;
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Us[s * BJ_UP + 32 * wr + 8 * i + lr];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = T[s * BJ_AP + 32 * wc + 8 * j + lr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    const int c0 = ch * BJ_AC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 32 * wr + 8 * i + lr;
      const int64_t row = (r < BJ_B) ? (int64_t)p * BJ_B + r : (int64_t)q * BJ_B + (r - BJ_B);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + 32 * wc + 8 * j + 2 * lk;
        if (c < n) *reinterpret_cast<double2*>(W + row * ld + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
    __syncthreads();  // every warp is done with this buffer before the next iteration refills it
  }
}

// out[i][r] = W[order[r]][i] * scale[order[r]]  (the kept right singular vectors as unit columns)
__global__ void bj_gather_kernel(const double* __restrict__ W, int64_t ld, int n, const int* __restrict__ order,
                                 const double* __restrict__ scale, int c, double* __restrict__ out, int64_t ldo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i >= n || r >= c) return;
  const int src = order[r];
  out[(int64_t)i * ldo + r] = W[(int64_t)src * ld + i] * scale[src];
}

// out[r][i] = W[order[r]][i] * scale[r]  (c x n: the kept vectors as scaled ROWS, for V.diag(s).V^T as a Gram)
__global__ void bj_gather_rows_kernel(const double* __restrict__ W, int64_t ld, int n, const int* __restrict__ order,
                                      const double* __restrict__ scale, int c, double* __restrict__ out, int64_t ldo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (i >= n || r >= c) return;
  out[(int64_t)r * ldo + i] = W[(int64_t)order[r] * ld + i] * scale[r];
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_gather_rows_scaled_f64(const double* W, int64_t ld, int n, const int32_t* order,
                                            const double* scale, int c, double* out, int64_t ldo, void* stream) {
  if (!W || !order || !scale || !out || n <= 0 || c <= 0 || ldo < n) return SCRNA_ERR_BADARG;
  bj_gather_rows_kernel<<<dim3((unsigned)ceil_div(n, 256), c), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, ld, n, order, scale, c, out, ldo);
  return last_launch_status();
}

extern "C" int scrna_bjacobi_blocks(int n) { return (int)ceil_div(n, BJ_B); }

extern "C" int scrna_bjacobi_splits(int n) {
  // enough Gram CTAs (pairs x splits) for six per SM (the staging of one overlaps the DMMAs of the others),
  // spans of at least 256 columns
  const int pairs = (scrna_bjacobi_blocks(n) + 1) / 2;
  int s = (int)ceil_div(6 * (int64_t)num_sms(), pairs > 0 ? pairs : 1);
  if (s > n / 256) s = n / 256;
  if (s < 1) s = 1;
  if (s > 16) s = 16;
  return s;
}

// work: pairs * (splits + 1) * 4096 doubles, pairs = ceil(nblk / 2)
extern "C" int64_t scrna_bjacobi_work_elems(int n) {
  const int nblk = scrna_bjacobi_blocks(n);
  const int m = (nblk + 1) / 2 * 2;
  return (int64_t)(m / 2) * (scrna_bjacobi_splits(n) + 1) * (BJ_P * BJ_P);
}

extern "C" int scrna_bjacobi_sweep_f64(double* W, int64_t ld, int n, double tol, double null_sq, double* work,
                                       int32_t* rotated, void* stream) {
  if (!W || !work || !rotated || n <= 0 || ld < n) return SCRNA_ERR_BADARG;
  const int nblk = scrna_bjacobi_blocks(n);
  if (nblk < 2) return SCRNA_ERR_UNSUPPORTED;  // n <= 32: use the unblocked kernel (scrna_jacobi_sweep_f64)
  const int m = (nblk + 1) / 2 * 2;
  const int pairs = m / 2;
  const int nsplit = scrna_bjacobi_splits(n);
  double* part = work;
  double* U = work + (int64_t)pairs * nsplit * (BJ_P * BJ_P);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr = false;
  const size_t smem = 2 * (size_t)BJ_P * BJ_LD * sizeof(double);
  const size_t smem_apply = ((size_t)BJ_P * BJ_UP + (size_t)BJ_P * BJ_AP) * sizeof(double);
  const size_t smem_pipe = ((size_t)BJ_P * BJ_UP + 2 * (size_t)BJ_P * BJ_AP) * sizeof(double);
  const bool pipelined = (n % 2 == 0) && (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int nchunks = (int)ceil_div(n, BJ_AC);
  int cgroups = (int)ceil_div(6 * (int64_t)num_sms(), pairs);   // about six waves of one CTA per SM
  if (cgroups > nchunks) cgroups = nchunks;
  if (cgroups < 1) cgroups = 1;
  const int cpc = (int)ceil_div(nchunks, cgroups);
  cgroups = (int)ceil_div(nchunks, cpc);
  if (!attr) {
    cudaFuncSetAttribute(bj_apply_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pipe);
    cudaFuncSetAttribute(bj_eigen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bj_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply);
    attr = true;
  }
  for (int round = 0; round < m - 1; ++round) {
    bj_gram_kernel<<<dim3(pairs, nsplit), 256, 0, st>>>(W, ld, n, nblk, m, round, nsplit, part);
    bj_eigen_kernel<<<pairs, BJ_ET, smem, st>>>(part, nsplit, nblk, m, round, tol, null_sq, U, rotated);
    if (pipelined)
      bj_apply_pipe_kernel<<<dim3(pairs, cgroups), 256, smem_pipe, st>>>(W, ld, n, nblk, m, round, U, cpc);
    else
      bj_apply_kernel<<<dim3(pairs, (unsigned)ceil_div(n, BJ_AC)), 256, smem_apply, st>>>(W, ld, n, nblk, m, round, U);
  }
  return last_launch_status();
}

extern "C" int scrna_gather_scaled_f64(const double* W, int64_t ld, int n, const int32_t* order, const double* scale,
                                       int c, double* out, int64_t ldo, void* stream) {
  if (!W || !order || !scale || !out || n <= 0 || c <= 0 || ldo < c) return SCRNA_ERR_BADARG;
  bj_gather_kernel<<<dim3((unsigned)ceil_div(n, 256), c), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, ld, n, order, scale, c, out, ldo);
  return last_launch_status();
}
