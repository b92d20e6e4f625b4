// SC3 k-means ensemble for sm_100a: sklearn KMeans(n_init=10, init='k-means++', algorithm='lloyd') as the
// reference calls it (scRNA/sc3_clustering_impl.py:206-219 -> SP/sklearn/cluster/_kmeans.py:180-279,
// 630-758, 1436-1562; _k_means_lloyd.pyx:168-218; _k_means_common.pyx:167-300), fp64.
//
//   km_prepare      : X = V[:, :d] - mean, squared row norms, tol = mean(var) * 1e-4
//   K14 kpp_*       : greedy k-means++ with the uniforms pre-drawn on the host (the reference draws them
//                     from the global legacy RandomState)
//   K15 lloyd_assign: labels = argmin_j (|c_j|^2 - 2 x.c_j), ties -> lowest j (one warp per cell)
//   K16 lloyd_update: per-cluster sums in cell order (deterministic), average, shifts, change flag
//   K17 inertia
// The host (scrna_b200/sc3_clustering_impl.py) only sequences launches and reads three scalars per
// Lloyd iteration for the stop rule.
#include "common.cuh"

namespace scrna {

// state block of one k-means run (device, doubles): see KM_* offsets
// KM_DONE / KM_ITERS: device-side stop rule of the Lloyd loop (scrna_lloyd_iteration): 1 = labels unchanged (strict
// convergence), 2 = centre shift within tol, 3 = max_iter; every kernel of a gated iteration returns at once when it
// is set, so the host enqueues iterations in batches and polls instead of reading scalars after each one.
enum { KM_TOL = 0, KM_POT = 1, KM_SHIFT = 2, KM_INERTIA = 3, KM_CHANGED = 4, KM_EMPTY = 5, KM_DONE = 6, KM_ITERS = 7,
       KM_STATE = 8 };
__device__ __forceinline__ bool km_done(const double* gate) { return gate != nullptr && gate[KM_DONE] != 0.0; }

// Element strides between two restarts of a batched fit (all zero for a single restart); the restart is blockIdx.z.
struct KmBatch {
  long long centres, labels, counts, state, celld, psum, pcnt;
};

// ---- prepare ---------------------------------------------------------------------------------------
// column means and (biased) variances in two passes over V, each a segmented sum: CTA = 128 features x one of
// KM_CSEG row segments, partials combined in segment order (fixed order: run-to-run identical)
constexpr int KM_CSEG = 64;
template <bool DEV>
__global__ void __launch_bounds__(128)
km_colpart_kernel(const double* __restrict__ V, int64_t ldv, int n, int d, const double* __restrict__ mean,
                  double* __restrict__ part /* KM_CSEG x d */) {
  const int f = blockIdx.x * 128 + threadIdx.x;
  if (f >= d) return;
  const int per = (n + KM_CSEG - 1) / KM_CSEG;
  const int i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  const double m = DEV ? mean[f] : 0.0;
  const double* v = V + f;
  double s = 0.0;
  int i = i0;
  for (; i + 8 <= i1; i += 8) {
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = v[(int64_t)(i + u) * ldv];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (DEV) { const double t = x[u] - m; s += t * t; }
      else s += x[u];
    }
  }
  for (; i < i1; ++i) {
    if (DEV) { const double t = v[(int64_t)i * ldv] - m; s += t * t; }
    else s += v[(int64_t)i * ldv];
  }
  part[(int64_t)blockIdx.y * d + f] = s;
}
__global__ void km_colfinal_kernel(const double* __restrict__ part, int n, int d, double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= d) return;
  double s = 0.0;
  for (int g = 0; g < KM_CSEG; ++g) s += part[(int64_t)g * d + f];
  out[f] = s / (double)n;
}

__global__ void km_center_kernel(const double* __restrict__ V, int64_t ldv, int n, int d,
                                 const double* __restrict__ mean, double* __restrict__ X, int64_t ldx,
                                 double* __restrict__ xsq) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = 0.0;
  for (int f = lane; f < d; f += 32) {
    const double v = V[(int64_t)i * ldv + f] - mean[f];
    X[(int64_t)i * ldx + f] = v;
    s = fma(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) xsq[i] = s;
}

__global__ void km_tol_kernel(const double* __restrict__ var, int d, double* __restrict__ state) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int f = 0; f < d; ++f) s += var[f];
    state[KM_TOL] = (s / (double)d) * 1e-4;
  }
}

// ---- k-means++ -------------------------------------------------------------------------------------
// first centre: RandomState.choice(n, p=1/n) == searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right')
__global__ void kpp_first_kernel(int n, double u, int* __restrict__ idx) {
  if (threadIdx.x != 0) return;
  const double p = 1.0 / (double)n;
  // cdf[i] = fl(cumsum)(i) / cdf[n-1]; sequential like numpy
  double c = 0.0, last = 0.0;
  for (int i = 0; i < n; ++i) last += p;
  int pick = n;
  for (int i = 0; i < n; ++i) {
    c += p;
    if (c / last > u) { pick = i; break; }
  }
  if (pick > n - 1) pick = n - 1;
  idx[0] = pick;
}

// d2[t][i] = max(xsq[i] + xsq[c_t] - 2 x_i.x_{c_t}, 0); optionally min with closest[i]; one warp per cell
template <int MAXT>
__global__ void __launch_bounds__(256)
kpp_dist_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const double* __restrict__ xsq,
                const int* __restrict__ cand, int ntr, const double* __restrict__ closest,
                double* __restrict__ out /* ntr x n */, long long s_cand = 0, long long s_closest = 0,
                long long s_out = 0) {
  // restart of a batched initialisation = blockIdx.y (strides zero otherwise)
  cand += blockIdx.y * s_cand; out += blockIdx.y * s_out;
  if (closest) closest += blockIdx.y * s_closest;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  double dot[MAXT];
#pragma unroll
  for (int t = 0; t < MAXT; ++t) dot[t] = 0.0;
  const double* xi = X + (int64_t)i * ldx;
  for (int f = lane; f < d; f += 32) {
    const double v = xi[f];
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t < ntr) dot[t] = fma(v, X[(int64_t)cand[t] * ldx + f], dot[t]);
  }
#pragma unroll
  for (int t = 0; t < MAXT; ++t) dot[t] = warp_sum(dot[t]);
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
      if (t < ntr) {
        // _euclidean_distances: (-2*dot + XX) + YY, clipped at 0
        double v = (-2.0 * dot[t] + xsq[cand[t]]) + xsq[i];
        v = fmax(v, 0.0);
        if (closest) v = fmin(closest[i], v);
        out[(int64_t)t * n + i] = v;
      }
  }
}

// sums of each row of a (rows x n) matrix in a fixed order: one block per row
__global__ void __launch_bounds__(1024)
row_sums_kernel(const double* __restrict__ M, int n, double* __restrict__ out, long long s_m = 0, long long s_out = 0) {
  __shared__ double scratch[32];
  M += blockIdx.y * s_m; out += blockIdx.y * s_out;
  const double* row = M + (int64_t)blockIdx.x * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += row[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// candidates = searchsorted(cumsum(closest), u_t * pot), clipped to n-1.  Single block; the cumsum is a
// chunked scan (per-thread sequential chunks, then offsets) kept in global scratch.
__global__ void __launch_bounds__(1024)
kpp_candidates_kernel(const double* __restrict__ closest, int n, const double* __restrict__ uniforms, int ntr,
                      const double* __restrict__ state, double* __restrict__ cums, int* __restrict__ cand,
                      long long s_uniforms = 0) {
  __shared__ double chunk_sum[1024];
  closest += (int64_t)blockIdx.y * n; cums += (int64_t)blockIdx.y * n; uniforms += blockIdx.y * s_uniforms;
  state += blockIdx.y * KM_STATE; cand += blockIdx.y * 8;
  const int T = blockDim.x;
  const int per = (n + T - 1) / T;
  const int b = threadIdx.x * per, e = min(n, b + per);
  double s = 0.0;
  for (int i = b; i < e; ++i) s += closest[i];
  chunk_sum[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    for (int t = 0; t < T; ++t) { const double v = chunk_sum[t]; chunk_sum[t] = run; run += v; }
  }
  __syncthreads();
  double run = chunk_sum[threadIdx.x];
  for (int i = b; i < e; ++i) { run += closest[i]; cums[i] = run; }
  __syncthreads();
  if (threadIdx.x < ntr) {
    const double val = uniforms[threadIdx.x] * state[KM_POT];
    int lo = 0, hi = n;  // first index with cums[idx] >= val (side='left')
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cums[mid] < val) lo = mid + 1; else hi = mid;
    }
    if (lo > n - 1) lo = n - 1;
    cand[threadIdx.x] = lo;
  }
}

// pick the candidate with the smallest potential (first minimum), adopt its distances
__global__ void kpp_select_kernel(const double* __restrict__ pots, int ntr, const int* __restrict__ cand,
                                  const double* __restrict__ dmin /* ntr x n */, int n,
                                  double* __restrict__ closest, double* __restrict__ state,
                                  int* __restrict__ idx, int c, int k = 0, long long s_dmin = 0) {
  __shared__ int best;
  pots += blockIdx.y * 8; cand += blockIdx.y * 8; dmin += blockIdx.y * s_dmin; closest += (int64_t)blockIdx.y * n;
  state += blockIdx.y * KM_STATE; idx += blockIdx.y * k;
  if (threadIdx.x == 0) {
    int b = 0;
    for (int t = 1; t < ntr; ++t)
      if (pots[t] < pots[b]) b = t;
    best = b;
    state[KM_POT] = pots[b];
    idx[c] = cand[b];
  }
  __syncthreads();
  const double* src = dmin + (int64_t)best * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) closest[i] = src[i];
}

__global__ void kpp_first_pot_kernel(const double* __restrict__ pots, double* __restrict__ state) {
  if (threadIdx.x == 0) state[blockIdx.x * KM_STATE + KM_POT] = pots[blockIdx.x * 8];
}

__global__ void gather_centers_kernel(const double* __restrict__ X, int64_t ldx, int d, const int* __restrict__ idx,
                                      int k, double* __restrict__ C, int64_t ldc, long long s_c = 0) {
  idx += blockIdx.z * k; C += blockIdx.z * s_c;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (f < d && j < k) C[(int64_t)j * ldc + f] = X[(int64_t)idx[j] * ldx + f];
}

// ---- Lloyd ---------------------------------------------------------------------------------------
// One warp owns ASSIGN_CW cells and sweeps the features once per block of ASSIGN_KB centres: the cell entries stay
// in registers, every centre entry read from L1 serves all the warp's cells, and each (cell, centre) dot product is
// accumulated per lane in feature order and then warp-reduced -- the same arithmetic, in the same order, as one dot
// product at a time.
constexpr int ASSIGN_CW = 2, ASSIGN_KB = 8;
__global__ void __launch_bounds__(256)
lloyd_assign_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const double* __restrict__ C,
                    int64_t ldc, int k, int* __restrict__ labels, const double* __restrict__ gate) {
  extern __shared__ double csq[];  // k squared centre norms
  if (km_done(gate)) return;
  for (int j = threadIdx.x >> 5; j < k; j += (blockDim.x >> 5)) {
    double s = 0.0;
    for (int f = threadIdx.x & 31; f < d; f += 32) { const double v = C[(int64_t)j * ldc + f]; s = fma(v, v, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) csq[j] = s;
  }
  __syncthreads();
  const int i0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ASSIGN_CW;
  const int lane = threadIdx.x & 31;
  if (i0 >= n) return;
  const double* xr[ASSIGN_CW];
#pragma unroll
  for (int c = 0; c < ASSIGN_CW; ++c) xr[c] = X + (int64_t)min(i0 + c, n - 1) * ldx;
  double best[ASSIGN_CW];
  int bj[ASSIGN_CW];
#pragma unroll
  for (int c = 0; c < ASSIGN_CW; ++c) { best[c] = 0.0; bj[c] = 0; }
  for (int j0 = 0; j0 < k; j0 += ASSIGN_KB) {
    const int kb = min(ASSIGN_KB, k - j0);
    const double* cb = C + (int64_t)j0 * ldc;
    double acc[ASSIGN_CW][ASSIGN_KB];
#pragma unroll
    for (int c = 0; c < ASSIGN_CW; ++c)
#pragma unroll
      for (int q = 0; q < ASSIGN_KB; ++q) acc[c][q] = 0.0;
    int f = lane;
    for (; f + 96 < d; f += 128) {   // four feature steps of every cell in flight before the first use
      double x[4][ASSIGN_CW];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < ASSIGN_CW; ++c) x[u][c] = xr[c][f + 32 * u];
      asm volatile("" ::: "memory");   // keep the loads batched: the compiler otherwise interleaves them with their uses
#pragma unroll
      for (int q = 0; q < ASSIGN_KB; ++q) {
        if (q < kb) {
          const double* cq = cb + (int64_t)q * ldc + f;
          const double c0 = cq[0], c1 = cq[32], c2 = cq[64], c3 = cq[96];
#pragma unroll
          for (int c = 0; c < ASSIGN_CW; ++c)
            acc[c][q] = fma(x[3][c], c3, fma(x[2][c], c2, fma(x[1][c], c1, fma(x[0][c], c0, acc[c][q]))));
        }
      }
    }
    for (; f < d; f += 32) {
#pragma unroll
      for (int q = 0; q < ASSIGN_KB; ++q) {
        if (q < kb) {
          const double c0 = cb[(int64_t)q * ldc + f];
#pragma unroll
          for (int c = 0; c < ASSIGN_CW; ++c) acc[c][q] = fma(xr[c][f], c0, acc[c][q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < ASSIGN_KB; ++q) {
      if (q < kb) {
#pragma unroll
        for (int c = 0; c < ASSIGN_CW; ++c) {
          const double v = csq[j0 + q] - 2.0 * warp_sum(acc[c][q]);
          if (j0 + q == 0 || v < best[c]) { best[c] = v; bj[c] = j0 + q; }
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < ASSIGN_CW; ++c)
      if (i0 + c < n) labels[i0 + c] = bj[c];
  }
}

// sums[j][f] = sum over cells with label j; counts[j].  The cells are cut into KM_SEG fixed segments: a
// block owns 128 features of one segment and adds each cell (in cell order) into the shared-memory row of its
// label -- one pass over X for all k clusters; the segment partials are then added in segment order, so the
// result is deterministic (sklearn's own order depends on its OpenMP chunking, _k_means_lloyd.pyx:168-218).
constexpr int KM_SEG = 32;
__global__ void __launch_bounds__(128)
lloyd_sums_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const int* __restrict__ labels, int k,
                  double* __restrict__ psum, int* __restrict__ pcnt, const double* __restrict__ gate) {
  extern __shared__ double sacc[];  // k x 128
  if (km_done(gate)) return;
  const int seg = blockIdx.y, tid = threadIdx.x;
  const int f = blockIdx.x * 128 + tid;
  const int per = (n + KM_SEG - 1) / KM_SEG;
  const int i0 = seg * per, i1 = min(n, i0 + per);
  for (int j = 0; j < k; ++j) sacc[j * 128 + tid] = 0.0;
  if (f < d) {
    const double* xf = X + f;
    int i = i0;
    // Register double buffer: the sixteen (label, entry) pairs of the NEXT step are requested before the current
    // sixteen are added (in cell order) into shared memory, so sixteen loads per thread stay in flight.
    if (i + 16 <= i1) {
      int l[16];
      double x[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { l[u] = labels[i + u]; x[u] = xf[(int64_t)(i + u) * ldx]; }
      for (; i + 32 <= i1; i += 16) {
        int ln[16];
        double xn[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { ln[u] = labels[i + 16 + u]; xn[u] = xf[(int64_t)(i + 16 + u) * ldx]; }
#pragma unroll
        for (int u = 0; u < 16; ++u) sacc[l[u] * 128 + tid] += x[u];
#pragma unroll
        for (int u = 0; u < 16; ++u) { l[u] = ln[u]; x[u] = xn[u]; }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) sacc[l[u] * 128 + tid] += x[u];
      i += 16;
    }
    for (; i + 4 <= i1; i += 4) {
      const int l0 = labels[i], l1 = labels[i + 1], l2 = labels[i + 2], l3 = labels[i + 3];
      const double x0 = xf[(int64_t)i * ldx], x1 = xf[(int64_t)(i + 1) * ldx], x2 = xf[(int64_t)(i + 2) * ldx],
                   x3 = xf[(int64_t)(i + 3) * ldx];
      sacc[l0 * 128 + tid] += x0;
      sacc[l1 * 128 + tid] += x1;
      sacc[l2 * 128 + tid] += x2;
      sacc[l3 * 128 + tid] += x3;
    }
    for (; i < i1; ++i) sacc[labels[i] * 128 + tid] += xf[(int64_t)i * ldx];
    for (int j = 0; j < k; ++j) psum[((int64_t)seg * k + j) * d + f] = sacc[j * 128 + tid];
  }
  if (blockIdx.x == 0 && tid < k) {
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += (labels[i] == tid);
    pcnt[seg * k + tid] = cnt;
  }
}

__global__ void lloyd_sums_final_kernel(const double* __restrict__ psum, const int* __restrict__ pcnt, int k, int d,
                                        double* __restrict__ sums, int64_t lds, int* __restrict__ counts,
                                        const double* __restrict__ gate, KmBatch bt) {
  const int rr = blockIdx.z;
  psum += rr * bt.psum; pcnt += rr * bt.pcnt; sums += rr * bt.centres; counts += rr * bt.counts;
  if (gate) gate += rr * bt.state;
  if (km_done(gate)) return;
  const int j = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < d) {
    double s = 0.0;
    for (int seg = 0; seg < KM_SEG; ++seg) s += psum[((int64_t)seg * k + j) * d + f];
    sums[(int64_t)j * lds + f] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int c = 0;
    for (int seg = 0; seg < KM_SEG; ++seg) c += pcnt[seg * k + j];
    counts[j] = c;
  }
}

// averages, empty-cluster handling (copy from the heaviest cluster, _k_means_common.pyx:262-280), shifts,
// label-change flag; single block
__global__ void __launch_bounds__(1024)
lloyd_finish_kernel(double* __restrict__ Cnew, const double* __restrict__ Cold, int64_t ldc, int k, int d,
                    const int* __restrict__ counts, const int* __restrict__ labels, int* __restrict__ labels_old,
                    int n, double* __restrict__ state, int max_iter, KmBatch bt) {
  __shared__ double scratch[32];
  const int rr = blockIdx.z;
  Cnew += rr * bt.centres; Cold += rr * bt.centres; counts += rr * bt.counts; labels += rr * bt.labels;
  labels_old += rr * bt.labels; state += rr * bt.state;
  if (max_iter > 0 && state[KM_DONE] != 0.0) return;
  __shared__ int s_changed, s_amax, s_empty;
  if (threadIdx.x == 0) {
    int am = 0, ne = 0;
    for (int j = 0; j < k; ++j) { if (counts[j] > counts[am]) am = j; ne += (counts[j] == 0); }
    s_amax = am; s_empty = ne; s_changed = 0;
  }
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (counts[j] > 0) {
      const double alpha = 1.0 / (double)counts[j];
      for (int f = threadIdx.x; f < d; f += blockDim.x) Cnew[(int64_t)j * ldc + f] *= alpha;
    }
  }
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (counts[j] == 0)
      for (int f = threadIdx.x; f < d; f += blockDim.x) Cnew[(int64_t)j * ldc + f] = Cnew[(int64_t)s_amax * ldc + f];
  }
  __syncthreads();
  double tot = 0.0;  // sum_j shift_j^2 with shift_j = sqrt(sum_f diff^2)
  for (int j = 0; j < k; ++j) {
    double s = 0.0;
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
      const double t = Cnew[(int64_t)j * ldc + f] - Cold[(int64_t)j * ldc + f];
      s = fma(t, t, s);
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) { const double sh = sqrt(s); tot += sh * sh; }
    __syncthreads();
  }
  int ch = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int l = labels[i];
    ch |= (l != labels_old[i]);
    labels_old[i] = l;
  }
  if (ch) s_changed = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    state[KM_SHIFT] = tot;
    state[KM_CHANGED] = (double)s_changed;
    state[KM_EMPTY] = (double)s_empty;
    if (max_iter > 0) {   // gated loop: the stop rule of sklearn's _kmeans_single_lloyd (_kmeans.py:700-722) on the device
      const double it = state[KM_ITERS] + 1.0;
      state[KM_ITERS] = it;
      if (s_changed == 0) state[KM_DONE] = 1.0;
      else if (tot <= state[KM_TOL]) state[KM_DONE] = 2.0;
      else if (it >= (double)max_iter) state[KM_DONE] = 3.0;
    }
  }
}

// per-cell squared distance to its assigned centre
__global__ void __launch_bounds__(256)
km_cell_dist_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const double* __restrict__ C,
                    int64_t ldc, const int* __restrict__ labels, double* __restrict__ out) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const double* xi = X + (int64_t)i * ldx;
  const double* cj = C + (int64_t)labels[i] * ldc;
  double s = 0.0;
  for (int f = lane; f < d; f += 32) { const double t = xi[f] - cj[f]; s = fma(t, t, s); }
  s = warp_sum(s);
  if (lane == 0) out[i] = s;
}

__global__ void km_store_sum_kernel(const double* __restrict__ src, double* __restrict__ state, int slot) {
  if (threadIdx.x == 0) state[slot] = src[0];
}

// Relocate empty clusters (_k_means_common.pyx:167-212): every empty cluster (ascending id) takes the cell
// that is farthest from its (old) centre, in descending distance order, whose contribution leaves its old
// cluster.  Single block, launched every iteration: returns at once when no cluster is empty (the normal
// case), otherwise computes the cell distances itself.
__global__ void __launch_bounds__(1024)
km_relocate_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const double* __restrict__ Cold,
                   int64_t ldc, double* __restrict__ sums, int64_t lds, int* __restrict__ counts,
                   const int* __restrict__ labels, double* __restrict__ celld, int k, const double* __restrict__ gate,
                   KmBatch bt) {
  __shared__ double bv[32];
  const int rr = blockIdx.z;
  Cold += rr * bt.centres; sums += rr * bt.centres; counts += rr * bt.counts; labels += rr * bt.labels;
  celld += rr * bt.celld;
  if (gate) gate += rr * bt.state;
  if (km_done(gate)) return;
  __shared__ int bi[32];
  __shared__ int s_far, s_any;
  __shared__ double s_max;
  if (threadIdx.x == 0) {
    int any = 0;
    for (int j = 0; j < k; ++j) any |= (counts[j] == 0);
    s_any = any;
  }
  __syncthreads();
  if (!s_any) return;
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < n; i += nw) {
      const double* xi = X + (int64_t)i * ldx;
      const double* cj = Cold + (int64_t)labels[i] * ldc;
      double s = 0.0;
      for (int f = lane; f < d; f += 32) { const double t = xi[f] - cj[f]; s = fma(t, t, s); }
      s = warp_sum(s);
      if (lane == 0) celld[i] = s;
    }
  }
  __syncthreads();
  for (int e = 0; e < k; ++e) {
    if (counts[e] != 0) continue;  // uniform across the block
    double mv = -1.0; int mi = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (celld[i] > mv) { mv = celld[i]; mi = i; }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, mv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (ov > mv || (ov == mv && oi >= 0 && (mi < 0 || oi < mi))) { mv = ov; mi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = mv; bi[threadIdx.x >> 5] = mi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = -1.0; int ix = -1;
      for (int w = 0; w < (blockDim.x >> 5); ++w)
        if (bv[w] > v || (bv[w] == v && bi[w] >= 0 && (ix < 0 || bi[w] < ix))) { v = bv[w]; ix = bi[w]; }
      s_far = ix; s_max = v;
    }
    __syncthreads();
    if (s_max <= 0.0 || s_far < 0) return;  // np.max(distances) == 0: nothing to relocate
    const int far = s_far, old = labels[far];
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
      const double v = X[(int64_t)far * ldx + f];
      sums[(int64_t)old * lds + f] -= v;
      sums[(int64_t)e * lds + f] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) { counts[e] = 1; counts[old] -= 1; celld[far] = -1.0; }
    __syncthreads();
  }
}

// ---- all restarts of a fit in one launch set -----------------------------------------------------------------------
// The n_init restarts of a fit are independent given their k-means++ centres, so one Lloyd iteration of ALL of them
// is one pass over X: the assignment is an fp64 GEMM X (n x d) . C^T (d x R*k) on the FP64 tensor cores with the
// argmin as its epilogue, the centre sums read each entry of X once for all restarts.  A restart whose stop rule has
// fired (state[r][KM_DONE]) is skipped tile by tile.
constexpr int KB_KC = 16;            // features per pipeline stage
constexpr int KB_S = KB_KC + 4;      // shared-memory pitch (doubles): conflict-free DMMA fragment reads
constexpr int KB_M = 64;             // cells per CTA (4 warps x 16)
constexpr int KB_NCOL = 48;          // centre columns per CTA (6 DMMA n-tiles)
constexpr int KB_STAGES = 2;           // 36 KB of shared memory per CTA: six CTAs per SM, one wave at 20 000 cells
constexpr int KB_RMAX = 16;          // restarts per fit (64 features x 16 restarts = one CTA of the sums kernel)

__device__ __forceinline__ void km_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void km_cp16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// csq[r][j] = |centre j of restart r|^2 (j < k), one warp each
__global__ void __launch_bounds__(256)
km_csq_batched_kernel(const double* __restrict__ C, int64_t ldc, int64_t strideC, int k, int kp8, int R, int d,
                      double* __restrict__ csq, const double* __restrict__ state) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= R * kp8) return;
  const int r = w / kp8, j = w % kp8;
  if (state[r * KM_STATE + KM_DONE] != 0.0) return;
  double s = 0.0;
  if (j < k) {
    const double* c = C + r * strideC + (int64_t)j * ldc;
    for (int f = threadIdx.x & 31; f < d; f += 32) s = fma(c[f], c[f], s);
    s = warp_sum(s);
  }
  if ((threadIdx.x & 31) == 0) csq[w] = s;
}

// grid (ceil(n / 64), restart groups); 128 threads.  Restart group g owns restarts [g*rb, min(R, (g+1)*rb)), their
// centres are the columns of the B operand (kp8 = k rounded up to 8 columns per restart, rb * kp8 <= 48).
__global__ void __launch_bounds__(128)
lloyd_assign_batched_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const double* __restrict__ C,
                            int64_t ldc, int64_t strideC, int k, int kp8, int R, int rb,
                            const double* __restrict__ csq, int* __restrict__ labels,
                            const double* __restrict__ state) {
  extern __shared__ __align__(16) double kb_smem[];
  __shared__ int s_live[KB_RMAX];
  __shared__ int s_any;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  const int r_lo = blockIdx.y * rb, r_hi = min(R, r_lo + rb), nr = r_hi - r_lo;
  if (tid == 0) {
    int any = 0;
    for (int q = 0; q < nr; ++q) { s_live[q] = state[(r_lo + q) * KM_STATE + KM_DONE] == 0.0; any |= s_live[q]; }
    s_any = any;
  }
  __syncthreads();
  if (!s_any) return;
  const int i0 = blockIdx.x * KB_M;
  const int ncol = nr * kp8;                 // <= KB_NCOL
  const int t8 = kp8 >> 3;                   // DMMA n-tiles per restart
  double* xs = kb_smem;                                   // [stage][KB_M][KB_S]
  double* cs = kb_smem + KB_STAGES * KB_M * KB_S;         // [stage][KB_NCOL][KB_S]
  const uint32_t xs_a = (uint32_t)__cvta_generic_to_shared(xs), cs_a = (uint32_t)__cvta_generic_to_shared(cs);
  const int nk = (d + KB_KC - 1) / KB_KC;
  auto issue = [&](int kc) {
    const int st = kc % KB_STAGES, f0 = kc * KB_KC;
    for (int e = tid; e < KB_M * (KB_KC / 2); e += 128) {
      const int row = e >> 3, ch = e & 7, f = f0 + 2 * ch, i = i0 + row;
      const int valid = (i < n) ? max(0, min(2, d - f)) : 0;
      km_cp16(xs_a + (uint32_t)((st * KB_M + row) * KB_S + 2 * ch) * 8u,
              X + (int64_t)(valid ? i : 0) * ldx + (valid ? f : 0), (uint32_t)valid * 8u);
    }
    for (int e = tid; e < ncol * (KB_KC / 2); e += 128) {
      const int col = e >> 3, ch = e & 7, f = f0 + 2 * ch, q = col / kp8, j = col % kp8;
      const int valid = (j < k && s_live[q]) ? max(0, min(2, d - f)) : 0;
      km_cp16(cs_a + (uint32_t)((st * KB_NCOL + col) * KB_S + 2 * ch) * 8u,
              C + (r_lo + q) * strideC + (int64_t)(valid ? j : 0) * ldc + (valid ? f : 0), (uint32_t)valid * 8u);
    }
  };
  unsigned tmask = 0;                       // DMMA n-tiles of live restarts
  for (int t = 0; t < KB_NCOL / 8; ++t)
    if (8 * t < ncol && s_live[t / t8]) tmask |= 1u << t;
  double acc[2][KB_NCOL / 8][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < KB_NCOL / 8; ++t) acc[u][t][0] = acc[u][t][1] = 0.0;
  for (int sgl = 0; sgl < KB_STAGES - 1; ++sgl) {
    if (sgl < nk) issue(sgl);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group %0;" ::"n"(KB_STAGES - 2) : "memory");
    __syncthreads();
    if (kc + KB_STAGES - 1 < nk) issue(kc + KB_STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* xt = xs + (kc % KB_STAGES) * KB_M * KB_S + (warp * 16 + lr) * KB_S + lk;
    const double* ct = cs + (kc % KB_STAGES) * KB_NCOL * KB_S + lr * KB_S + lk;
#pragma unroll
    for (int ks = 0; ks < KB_KC / 4; ++ks) {
      const double a0 = xt[4 * ks], a1 = xt[8 * KB_S + 4 * ks];
#pragma unroll
      for (int t = 0; t < KB_NCOL / 8; ++t) {
        if (tmask >> t & 1u) {
          const double b = ct[t * 8 * KB_S + 4 * ks];
          km_dmma(acc[0][t], a0, b);
          km_dmma(acc[1][t], a1, b);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // epilogue: per cell and restart the lowest-index minimum of csq[j] - 2 x.c_j
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = i0 + warp * 16 + u * 8 + lr;
    for (int q = 0; q < nr; ++q) {
      if (!s_live[q]) continue;
      const int r = r_lo + q;
      double bv = 0.0;
      int bj = -1;
#pragma unroll
      for (int t = 0; t < KB_NCOL / 8; ++t) {
        if (t >= q * t8 && t < (q + 1) * t8) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = (t - q * t8) * 8 + 2 * lk + h;
            if (j < k) {
              const double v = csq[r * kp8 + j] - 2.0 * acc[u][t][h];
              if (bj < 0 || v < bv) { bv = v; bj = j; }
            }
          }
        }
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (oj >= 0 && (bj < 0 || ov < bv || (ov == bv && oj < bj))) { bv = ov; bj = oj; }
      }
      if (lk == 0 && i < n) {
        if (bj < 0) bj = 0;                  // every score NaN: the unbatched kernel keeps label 0 too
        labels[(int64_t)r * n + i] = bj;
      }
    }
  }
}

// k-means++ candidate distances of all restarts as the same GEMM: the B columns are the candidate ROWS of X (eight
// column slots per restart, ntr <= 8 of them used), the epilogue is sklearn's _euclidean_distances
// ((-2 x.c + |c|^2) + |x|^2, clipped at 0) and the running minimum with closest[].  grid (ceil(n / 64), restart groups
// of six).
__global__ void __launch_bounds__(128)
kpp_dist_batched_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const double* __restrict__ xsq,
                        const int* __restrict__ cand, long long s_cand, int ntr, int R,
                        const double* __restrict__ closest /* R x n or null */, double* __restrict__ out,
                        long long s_out /* per restart; candidate t at + t * n */) {
  extern __shared__ __align__(16) double kb_smem[];
  __shared__ int s_cand_row[KB_NCOL];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  constexpr int RB = KB_NCOL / 8;
  const int r_lo = blockIdx.y * RB, nr = min(R, r_lo + RB) - r_lo;
  const int ncol = nr * 8;
  if (tid < ncol) {
    const int q = tid >> 3, j = tid & 7;
    s_cand_row[tid] = j < ntr ? cand[(r_lo + q) * s_cand + j] : -1;
  }
  __syncthreads();
  const int i0 = blockIdx.x * KB_M;
  double* xs = kb_smem;
  double* cs = kb_smem + KB_STAGES * KB_M * KB_S;
  const uint32_t xs_a = (uint32_t)__cvta_generic_to_shared(xs), cs_a = (uint32_t)__cvta_generic_to_shared(cs);
  const int nk = (d + KB_KC - 1) / KB_KC;
  auto issue = [&](int kc) {
    const int st = kc % KB_STAGES, f0 = kc * KB_KC;
    for (int e = tid; e < KB_M * (KB_KC / 2); e += 128) {
      const int row = e >> 3, ch = e & 7, f = f0 + 2 * ch, i = i0 + row;
      const int valid = (i < n) ? max(0, min(2, d - f)) : 0;
      km_cp16(xs_a + (uint32_t)((st * KB_M + row) * KB_S + 2 * ch) * 8u,
              X + (int64_t)(valid ? i : 0) * ldx + (valid ? f : 0), (uint32_t)valid * 8u);
    }
    for (int e = tid; e < ncol * (KB_KC / 2); e += 128) {
      const int col = e >> 3, ch = e & 7, f = f0 + 2 * ch, src = s_cand_row[col];
      const int valid = src >= 0 ? max(0, min(2, d - f)) : 0;
      km_cp16(cs_a + (uint32_t)((st * KB_NCOL + col) * KB_S + 2 * ch) * 8u,
              X + (int64_t)(valid ? src : 0) * ldx + (valid ? f : 0), (uint32_t)valid * 8u);
    }
  };
  double acc[2][RB][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int t = 0; t < RB; ++t) acc[u][t][0] = acc[u][t][1] = 0.0;
  for (int sgl = 0; sgl < KB_STAGES - 1; ++sgl) {
    if (sgl < nk) issue(sgl);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group %0;" ::"n"(KB_STAGES - 2) : "memory");
    __syncthreads();
    if (kc + KB_STAGES - 1 < nk) issue(kc + KB_STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* xt = xs + (kc % KB_STAGES) * KB_M * KB_S + (warp * 16 + lr) * KB_S + lk;
    const double* ct = cs + (kc % KB_STAGES) * KB_NCOL * KB_S + lr * KB_S + lk;
#pragma unroll
    for (int ks = 0; ks < KB_KC / 4; ++ks) {
      const double a0 = xt[4 * ks], a1 = xt[8 * KB_S + 4 * ks];
#pragma unroll
      for (int t = 0; t < RB; ++t) {
        if (t < nr) {
          const double b = ct[t * 8 * KB_S + 4 * ks];
          km_dmma(acc[0][t], a0, b);
          km_dmma(acc[1][t], a1, b);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = i0 + warp * 16 + u * 8 + lr;
    if (i >= n) continue;
    const double xi = xsq[i];
#pragma unroll
    for (int t = 0; t < RB; ++t) {
      if (t < nr) {
        const int r = r_lo + t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 2 * lk + h;
          if (j < ntr) {
            double v = (-2.0 * acc[u][t][h] + xsq[s_cand_row[t * 8 + j]]) + xi;
            v = fmax(v, 0.0);
            if (closest) v = fmin(closest[(int64_t)r * n + i], v);
            out[r * s_out + (int64_t)j * n + i] = v;
          }
        }
      }
    }
  }
}

// Centre sums of all restarts in one launch: thread = (feature of a 64-feature block, restart), CTA = 64 features x
// R restarts x one cell segment.  Every thread runs exactly the loop of lloyd_sums_kernel for its restart (cells of the
// segment in order, the adds into its own shared-memory rows, sixteen loads in flight), so the sums are bit-identical
// to the unbatched kernel's; the R warps that need the same entries of X share them through L1, and the warps of a
// restart whose stop rule has fired leave at once.
constexpr int KB_SF = 64;
template <int THREADS, int DEPTH>   // up to THREADS = 64 x R threads per CTA; DEPTH loads in flight per thread
__global__ void __launch_bounds__(THREADS)
lloyd_sums_batched_kernel(const double* __restrict__ X, int64_t ldx, int n, int d, const int* __restrict__ labels_all,
                          int k, int R, double* __restrict__ psum, int* __restrict__ pcnt,
                          const double* __restrict__ state) {
  extern __shared__ double sacc[];  // R x k x 64
  const int seg = blockIdx.y, tx = threadIdx.x, r = threadIdx.y;
  if (state[r * KM_STATE + KM_DONE] != 0.0) return;        // uniform over the warp (blockDim.x = 64)
  const int f = blockIdx.x * KB_SF + tx;
  const int per = (n + KM_SEG - 1) / KM_SEG;
  const int i0 = seg * per, i1 = min(n, i0 + per);
  double* acc = sacc + (int64_t)r * k * KB_SF + tx;
  const int* labels = labels_all + (int64_t)r * n;
  for (int j = 0; j < k; ++j) acc[j * KB_SF] = 0.0;
  if (f < d) {
    const double* xf = X + f;
    int i = i0;
    if (i + DEPTH <= i1) {
      int l[DEPTH];
      double x[DEPTH];
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) { l[u] = labels[i + u]; x[u] = xf[(int64_t)(i + u) * ldx]; }
      for (; i + 2 * DEPTH <= i1; i += DEPTH) {
        int ln[DEPTH];
        double xn[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) { ln[u] = labels[i + DEPTH + u]; xn[u] = xf[(int64_t)(i + DEPTH + u) * ldx]; }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) acc[l[u] * KB_SF] += x[u];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) { l[u] = ln[u]; x[u] = xn[u]; }
      }
#pragma unroll
      for (int u = 0; u < DEPTH; ++u) acc[l[u] * KB_SF] += x[u];
      i += DEPTH;
    }
    for (; i + 4 <= i1; i += 4) {
      const int l0 = labels[i], l1 = labels[i + 1], l2 = labels[i + 2], l3 = labels[i + 3];
      const double x0 = xf[(int64_t)i * ldx], x1 = xf[(int64_t)(i + 1) * ldx], x2 = xf[(int64_t)(i + 2) * ldx],
                   x3 = xf[(int64_t)(i + 3) * ldx];
      acc[l0 * KB_SF] += x0;
      acc[l1 * KB_SF] += x1;
      acc[l2 * KB_SF] += x2;
      acc[l3 * KB_SF] += x3;
    }
    for (; i < i1; ++i) acc[labels[i] * KB_SF] += xf[(int64_t)i * ldx];
    for (int j = 0; j < k; ++j) psum[(((int64_t)r * KM_SEG + seg) * k + j) * d + f] = acc[j * KB_SF];
  }
  if (blockIdx.x == 0 && tx < k) {
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += (labels[i] == tx);
    pcnt[((int64_t)r * KM_SEG + seg) * k + tx] = cnt;
  }
}

}  // namespace scrna

using namespace scrna;

extern "C" int64_t scrna_km_prepare_work_elems(int d) { return d > 0 ? (int64_t)(2 + KM_CSEG) * d : SCRNA_ERR_BADARG; }

extern "C" int scrna_km_prepare(const double* V, int64_t ldv, int n, int d, double* X, int64_t ldx, double* xsq,
                                double* work, double* state, void* stream) {
  // work: scrna_km_prepare_work_elems(d) doubles (column means, variances, segment partials)
  if (!V || !X || !xsq || !work || !state || n <= 0 || d <= 0 || ldv < d || ldx < d) return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  double* part = work + 2 * (int64_t)d;
  const dim3 gp((unsigned)ceil_div(d, 128), KM_CSEG);
  km_colpart_kernel<false><<<gp, 128, 0, st>>>(V, ldv, n, d, nullptr, part);
  km_colfinal_kernel<<<(unsigned)ceil_div(d, 128), 128, 0, st>>>(part, n, d, work);
  km_colpart_kernel<true><<<gp, 128, 0, st>>>(V, ldv, n, d, work, part);
  km_colfinal_kernel<<<(unsigned)ceil_div(d, 128), 128, 0, st>>>(part, n, d, work + d);
  km_center_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(V, ldv, n, d, work, X, ldx, xsq);
  km_tol_kernel<<<1, 32, 0, st>>>(work + d, d, state);
  return last_launch_status();
}

__global__ void kpp_set_first_kernel(int first, int* __restrict__ idx) {
  if (threadIdx.x == 0) idx[0] = first;
}
__global__ void kpp_set_firsts_kernel(const int* __restrict__ firsts, int R, int k, int n, int* __restrict__ idx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) idx[r * k] = min(max(firsts[r], 0), n - 1);
}

static int kmeans_pp_launch(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, double u0, int first,
                            const double* uniforms, int ntrials, double* closest, double* cand_dist,
                            double* cums, double* pots, int32_t* cand, int32_t* idx, double* state, double* C,
                            int64_t ldc, void* stream) {
  // u0: the uniform of RandomState.choice; uniforms: (k-1)*ntrials doubles on the device (the
  // uniform(size=ntrials) draws of centres 1..k-1); cand_dist: ntrials*n; cums: n; pots: ntrials
  if (!X || !xsq || (!uniforms && k > 1) || !closest || !cand_dist || !cums || !pots || !cand || !idx || !state ||
      !C || n <= 0 || d <= 0 || k <= 0 || k > n || ntrials <= 0 || ntrials > 8 || ldx < d || ldc < d)
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (first >= 0) kpp_set_first_kernel<<<1, 32, 0, st>>>(first, idx);
  else kpp_first_kernel<<<1, 32, 0, st>>>(n, u0, idx);
  const unsigned gb = (unsigned)ceil_div(n, 8);
  kpp_dist_kernel<8><<<gb, 256, 0, st>>>(X, ldx, n, d, xsq, idx, 1, nullptr, closest);
  row_sums_kernel<<<1, 1024, 0, st>>>(closest, n, pots);
  kpp_first_pot_kernel<<<1, 32, 0, st>>>(pots, state);
  for (int c = 1; c < k; ++c) {
    kpp_candidates_kernel<<<1, 1024, 0, st>>>(closest, n, uniforms + (int64_t)(c - 1) * ntrials, ntrials, state,
                                             cums, cand);
    kpp_dist_kernel<8><<<gb, 256, 0, st>>>(X, ldx, n, d, xsq, cand, ntrials, closest, cand_dist);
    row_sums_kernel<<<ntrials, 1024, 0, st>>>(cand_dist, n, pots);
    kpp_select_kernel<<<64, 256, 0, st>>>(pots, ntrials, cand, cand_dist, n, closest, state, idx, c);
  }
  dim3 grid((unsigned)ceil_div(d, 128), (unsigned)k);
  gather_centers_kernel<<<grid, 128, 0, st>>>(X, ldx, d, idx, k, C, ldc);
  return last_launch_status();
}

extern "C" int scrna_kmeans_pp(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, double u0,
                               const double* uniforms, int ntrials, double* closest, double* cand_dist,
                               double* cums, double* pots, int32_t* cand, int32_t* idx, double* state, double* C,
                               int64_t ldc, void* stream) {
  return kmeans_pp_launch(X, ldx, n, d, xsq, k, u0, -1, uniforms, ntrials, closest, cand_dist, cums, pots, cand, idx,
                          state, C, ldc, stream);
}

// Same, with the first centre already chosen by the caller (RandomState.choice over a uniform p is a sequential
// n-term cumulative sum; the host evaluates it once per n with the reference's own NumPy arithmetic instead of a
// single device thread walking it for every restart).
extern "C" int scrna_kmeans_pp_from(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, int first,
                                    const double* uniforms, int ntrials, double* closest, double* cand_dist,
                                    double* cums, double* pots, int32_t* cand, int32_t* idx, double* state, double* C,
                                    int64_t ldc, void* stream) {
  if (first < 0 || first >= n) return SCRNA_ERR_BADARG;
  return kmeans_pp_launch(X, ldx, n, d, xsq, k, 0.0, first, uniforms, ntrials, closest, cand_dist, cums, pots, cand, idx,
                          state, C, ldc, stream);
}

// k-means++ of all restarts of a fit at once: every kernel of the sequence carries the restart as a grid dimension, so
// an initialisation costs the same 4k launches for n_init restarts as for one and the candidate distances of all
// restarts share the passes over X.  firsts: n_init first centres (device), uniforms: n_init * (1 + (k-1)*ntrials)
// doubles in the reference's draw order (the first of every restart is skipped: the host used it for `firsts`),
// closest / cums: n_init x n, cand_dist: n_init x ntrials x n, pots / cand: n_init x 8, idx: n_init x k,
// state: n_init x 8, C: n_init x k x ldc.
extern "C" int scrna_kmeans_pp_batched(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, int n_init,
                                       const int32_t* firsts, const double* uniforms, int ntrials, double* closest,
                                       double* cand_dist, double* cums, double* pots, int32_t* cand, int32_t* idx,
                                       double* state, double* C, int64_t ldc, void* stream) {
  if (!X || !xsq || !firsts || (!uniforms && k > 1) || !closest || !cand_dist || !cums || !pots || !cand || !idx ||
      !state || !C || n <= 0 || d <= 0 || k <= 0 || k > n || n_init <= 0 || n_init > 65535 || ntrials <= 0 ||
      ntrials > 8 || ldx < d || ldc < d)
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned R = (unsigned)n_init;
  const long long per = 1 + (long long)(k - 1) * ntrials;
  kpp_set_firsts_kernel<<<(R + 63) / 64, 64, 0, st>>>(firsts, n_init, k, n, idx);
  const unsigned gb = (unsigned)ceil_div(n, 8);
  // candidate distances as one fp64 tensor-core GEMM for all restarts when the rows allow 16-byte staging
  const bool gemm = !(ldx & 1) && !((uintptr_t)X & 15);
  const size_t smem_g = (size_t)KB_STAGES * (KB_M + KB_NCOL) * KB_S * sizeof(double);
  const dim3 gg((unsigned)ceil_div(n, KB_M), (unsigned)ceil_div(n_init, KB_NCOL / 8));
  if (gemm) {
    static bool attr = false;
    if (!attr) {
      if (cudaFuncSetAttribute(kpp_dist_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g) !=
          cudaSuccess)
        return SCRNA_ERR_CUDA;
      attr = true;
    }
    kpp_dist_batched_kernel<<<gg, 128, smem_g, st>>>(X, ldx, n, d, xsq, idx, k, 1, n_init, nullptr, closest, n);
  } else {
    kpp_dist_kernel<8><<<dim3(gb, R), 256, 0, st>>>(X, ldx, n, d, xsq, idx, 1, nullptr, closest, k, 0, n);
  }
  row_sums_kernel<<<dim3(1, R), 1024, 0, st>>>(closest, n, pots, n, 8);
  kpp_first_pot_kernel<<<R, 32, 0, st>>>(pots, state);
  for (int c = 1; c < k; ++c) {
    kpp_candidates_kernel<<<dim3(1, R), 1024, 0, st>>>(closest, n, uniforms + 1 + (int64_t)(c - 1) * ntrials, ntrials,
                                                      state, cums, cand, per);
    if (gemm)
      kpp_dist_batched_kernel<<<gg, 128, smem_g, st>>>(X, ldx, n, d, xsq, cand, 8, ntrials, n_init, closest, cand_dist,
                                                       (long long)ntrials * n);
    else
      kpp_dist_kernel<8><<<dim3(gb, R), 256, 0, st>>>(X, ldx, n, d, xsq, cand, ntrials, closest, cand_dist, 8, n,
                                                      (long long)ntrials * n);
    row_sums_kernel<<<dim3((unsigned)ntrials, R), 1024, 0, st>>>(cand_dist, n, pots, (long long)ntrials * n, 8);
    kpp_select_kernel<<<dim3(64, R), 256, 0, st>>>(pots, ntrials, cand, cand_dist, n, closest, state, idx, c, k,
                                                   (long long)ntrials * n);
  }
  gather_centers_kernel<<<dim3((unsigned)ceil_div(d, 128), (unsigned)k, R), 128, 0, st>>>(X, ldx, d, idx, k, C, ldc,
                                                                                        (long long)k * ldc);
  return last_launch_status();
}

extern "C" int scrna_lloyd_assign(const double* X, int64_t ldx, int n, int d, const double* C, int64_t ldc, int k,
                                  int32_t* labels, void* stream) {
  if (!X || !C || !labels || n <= 0 || d <= 0 || k <= 0 || ldx < d || ldc < d) return SCRNA_ERR_BADARG;
  lloyd_assign_kernel<<<(unsigned)ceil_div(n, 8 * ASSIGN_CW), 256, (size_t)k * sizeof(double), (cudaStream_t)stream>>>(
      X, ldx, n, d, C, ldc, k, labels, nullptr);
  return last_launch_status();
}

extern "C" int64_t scrna_lloyd_sums_work_elems(int k, int d) { return (int64_t)KM_SEG * k * (d + 1); }

extern "C" int scrna_lloyd_sums(const double* X, int64_t ldx, int n, int d, const int32_t* labels, int k,
                                double* sums, int64_t lds, int32_t* counts, double* work, void* stream) {
  if (!X || !labels || !sums || !counts || !work || n <= 0 || d <= 0 || k <= 0 || k > 128 || ldx < d || lds < d)
    return SCRNA_ERR_BADARG;
  double* psum = work;                                                     // KM_SEG x k x d
  int* pcnt = reinterpret_cast<int*>(work + (int64_t)KM_SEG * k * d);      // KM_SEG x k ints
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)k * 128 * sizeof(double);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(lloyd_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8);
    attr = true;
  }
  lloyd_sums_kernel<<<dim3((unsigned)ceil_div(d, 128), KM_SEG), 128, smem, st>>>(X, ldx, n, d, labels, k, psum, pcnt,
                                                                                 nullptr);
  lloyd_sums_final_kernel<<<dim3((unsigned)ceil_div(d, 128), (unsigned)k), 128, 0, st>>>(psum, pcnt, k, d, sums, lds,
                                                                                       counts, nullptr, KmBatch{});
  return last_launch_status();
}

// One gated Lloyd iteration (assign -> sums -> relocate -> finish with the stop rule on the device).  The host swaps
// Ccur / Cnxt every enqueued iteration; once state[KM_DONE] is set the remaining enqueued iterations are no-ops, and
// the final centres are in the buffer that was `Cnxt` of iteration state[KM_ITERS].
extern "C" int scrna_lloyd_iteration(const double* X, int64_t ldx, int n, int d, const double* Ccur, double* Cnxt,
                                     int64_t ldc, int k, int32_t* labels, int32_t* labels_old, int32_t* counts,
                                     double* celld, double* work, double* state, int max_iter, void* stream) {
  if (!X || !Ccur || !Cnxt || !labels || !labels_old || !counts || !celld || !work || !state || n <= 0 || d <= 0 ||
      k <= 0 || k > 128 || ldx < d || ldc < d || max_iter <= 0)
    return SCRNA_ERR_BADARG;
  double* psum = work;
  int* pcnt = reinterpret_cast<int*>(work + (int64_t)KM_SEG * k * d);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(lloyd_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8);
    attr = true;
  }
  lloyd_assign_kernel<<<(unsigned)ceil_div(n, 8 * ASSIGN_CW), 256, (size_t)k * sizeof(double), st>>>(X, ldx, n, d, Ccur, ldc, k, labels,
                                                                                         state);
  lloyd_sums_kernel<<<dim3((unsigned)ceil_div(d, 128), KM_SEG), 128, (size_t)k * 128 * sizeof(double), st>>>(
      X, ldx, n, d, labels, k, psum, pcnt, state);
  lloyd_sums_final_kernel<<<dim3((unsigned)ceil_div(d, 128), (unsigned)k), 128, 0, st>>>(psum, pcnt, k, d, Cnxt, ldc,
                                                                                       counts, state, KmBatch{});
  km_relocate_kernel<<<1, 1024, 0, st>>>(X, ldx, n, d, Ccur, ldc, Cnxt, ldc, counts, labels, celld, k, state, KmBatch{});
  lloyd_finish_kernel<<<1, 1024, 0, st>>>(Cnxt, Ccur, ldc, k, d, counts, labels, labels_old, n, state, max_iter,
                                          KmBatch{});
  return last_launch_status();
}

// Which fits the batched kernels take: <= 16 restarts, k <= 48 (one restart's centres fit one CTA's columns) and the
// per-restart shared-memory rows of the sums kernel within 200 KB.
extern "C" int scrna_lloyd_batched_supported(int k, int n_init) {
  return k >= 1 && k <= KB_NCOL && n_init >= 1 && n_init <= KB_RMAX && (int64_t)n_init * k * KB_SF * 8 <= 200 * 1024;
}
extern "C" int64_t scrna_lloyd_batched_work_elems(int k, int d, int n_init) {
  return (int64_t)n_init * KM_SEG * k * (d + 1) + (int64_t)n_init * (((k + 7) / 8) * 8);
}

// One gated Lloyd iteration of ALL restarts of a fit.  Ccur / Cnxt: n_init x k x ldc (the host swaps them every
// call), labels / labels_old: n_init x n, counts: n_init x k, celld: n_init x n, state: n_init x 8
// (zero [6..7] and set [0] = tol per restart before the first call), work: scrna_lloyd_batched_work_elems doubles.
// X, Ccur, Cnxt need 16-byte aligned rows (ldx, ldc even).
extern "C" int scrna_lloyd_iteration_batched(const double* X, int64_t ldx, int n, int d, const double* Ccur,
                                             double* Cnxt, int64_t ldc, int k, int n_init, int32_t* labels,
                                             int32_t* labels_old, int32_t* counts, double* celld, double* work,
                                             double* state, int max_iter, void* stream) {
  if (!X || !Ccur || !Cnxt || !labels || !labels_old || !counts || !celld || !work || !state || n <= 0 ||
      d <= 0 || ldx < d || ldc < d || (ldx & 1) || (ldc & 1) || max_iter <= 0 ||
      !scrna_lloyd_batched_supported(k, n_init) || ((uintptr_t)X & 15) || ((uintptr_t)Ccur & 15))
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = n_init, kp8 = ((k + 7) / 8) * 8;
  const int64_t strideC = (int64_t)k * ldc;
  double* psum = work;
  int* pcnt = reinterpret_cast<int*>(work + (int64_t)R * KM_SEG * k * d);
  double* csq = work + (int64_t)R * KM_SEG * k * (d + 1);
  const size_t smem_a = (size_t)KB_STAGES * (KB_M + KB_NCOL) * KB_S * sizeof(double);
  const size_t smem_s = (size_t)R * k * KB_SF * sizeof(double);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(lloyd_assign_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a) !=
            cudaSuccess ||
        cudaFuncSetAttribute(lloyd_sums_batched_kernel<640, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(lloyd_sums_batched_kernel<1024, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             200 * 1024) != cudaSuccess)
      return SCRNA_ERR_CUDA;
    attr = true;
  }
  const int groups = (int)ceil_div((int64_t)R * kp8, KB_NCOL);
  const int rb = (int)ceil_div(R, groups);
  km_csq_batched_kernel<<<(unsigned)ceil_div((int64_t)R * kp8, 8), 256, 0, st>>>(Ccur, ldc, strideC, k, kp8, R, d, csq,
                                                                               state);
  lloyd_assign_batched_kernel<<<dim3((unsigned)ceil_div(n, KB_M), (unsigned)ceil_div(R, rb)), 128, smem_a, st>>>(
      X, ldx, n, d, Ccur, ldc, strideC, k, kp8, R, rb, csq, labels, state);
  const dim3 gs((unsigned)ceil_div(d, KB_SF), KM_SEG), bs(KB_SF, (unsigned)R);
  if (R <= 10)
    lloyd_sums_batched_kernel<640, 16><<<gs, bs, smem_s, st>>>(X, ldx, n, d, labels, k, R, psum, pcnt, state);
  else
    lloyd_sums_batched_kernel<1024, 8><<<gs, bs, smem_s, st>>>(X, ldx, n, d, labels, k, R, psum, pcnt, state);
  KmBatch bt;
  bt.centres = strideC; bt.labels = n; bt.counts = k; bt.state = KM_STATE; bt.celld = n;
  bt.psum = (long long)KM_SEG * k * d; bt.pcnt = (long long)KM_SEG * k;
  lloyd_sums_final_kernel<<<dim3((unsigned)ceil_div(d, 128), (unsigned)k, (unsigned)R), 128, 0, st>>>(
      psum, pcnt, k, d, Cnxt, ldc, counts, state, bt);
  km_relocate_kernel<<<dim3(1, 1, (unsigned)R), 1024, 0, st>>>(X, ldx, n, d, Ccur, ldc, Cnxt, ldc, counts, labels, celld, k,
                                                              state, bt);
  lloyd_finish_kernel<<<dim3(1, 1, (unsigned)R), 1024, 0, st>>>(Cnxt, Ccur, ldc, k, d, counts, labels, labels_old, n, state,
                                                               max_iter, bt);
  return last_launch_status();
}

extern "C" int scrna_lloyd_finish(double* Cnew, const double* Cold, int64_t ldc, int k, int d, const int32_t* counts,
                                  const int32_t* labels, int32_t* labels_old, int n, double* state, void* stream) {
  if (!Cnew || !Cold || !counts || !labels || !labels_old || !state || k <= 0 || d <= 0 || n <= 0 || ldc < d)
    return SCRNA_ERR_BADARG;
  lloyd_finish_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(Cnew, Cold, ldc, k, d, counts, labels, labels_old, n, state,
                                                            0, KmBatch{});
  return last_launch_status();
}

extern "C" int scrna_km_cell_dist(const double* X, int64_t ldx, int n, int d, const double* C, int64_t ldc,
                                  const int32_t* labels, double* out, void* stream) {
  if (!X || !C || !labels || !out || n <= 0 || d <= 0 || ldx < d || ldc < d) return SCRNA_ERR_BADARG;
  km_cell_dist_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(X, ldx, n, d, C, ldc, labels, out);
  return last_launch_status();
}

extern "C" int scrna_km_inertia(const double* X, int64_t ldx, int n, int d, const double* C, int64_t ldc,
                                const int32_t* labels, double* celld, double* state, void* stream) {
  if (!X || !C || !labels || !celld || !state || n <= 0 || d <= 0 || ldx < d || ldc < d) return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  km_cell_dist_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(X, ldx, n, d, C, ldc, labels, celld);
  row_sums_kernel<<<1, 1024, 0, st>>>(celld, n, state + KM_INERTIA);
  return last_launch_status();
}

extern "C" int scrna_km_relocate(const double* X, int64_t ldx, int n, int d, const double* Cold, int64_t ldc,
                                 double* sums, int64_t lds, int32_t* counts, const int32_t* labels, double* celld,
                                 int k, void* stream) {
  if (!X || !Cold || !sums || !counts || !labels || !celld || n <= 0 || d <= 0 || k <= 0 || ldx < d || lds < d ||
      ldc < d)
    return SCRNA_ERR_BADARG;
  km_relocate_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(X, ldx, n, d, Cold, ldc, sums, lds, counts, labels, celld,
                                                          k, nullptr, KmBatch{});
  return last_launch_status();
}
