// NMF epilogues for sm_100a: partial reduction, k x k Grams, the multiplicative-update and
// coordinate-descent factor updates, and the device-side stop rule.
//
// These operate on the small rank-k factors (g x k and k x n) in fp64 -- the reference computes
// everything in float64 (abstract_clustering.py:28) and only the X products are accumulated in
// fp32 by the sweeps -- and write the fp32 shadow the next sweep reads.
#include "common.cuh"

namespace scrna {

constexpr int UPD_THREADS = 128;
constexpr int UPD_LD = UPD_THREADS + 1;  // smem row pitch (doubles): column reads of the Gram epilogue stay conflict-free

// ---- fixed-order reduction of the sweep partials ---------------------------------------------
__global__ void reduce_partials_kernel(const float* __restrict__ part, int nsplit, int64_t count,
                                       int64_t split_stride, double* __restrict__ out,
                                       const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < nsplit; ++p) s += (double)part[(int64_t)p * split_stride + i];
    out[i] = s;
  }
}

// K1 (row-major partials nsplit x g x kp) -> component-major fp64 out[t*ld + i]
__global__ void reduce_rows_t_kernel(const float* __restrict__ part, int nsplit, int g, int kp,
                                     double* __restrict__ out, int64_t ld, const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  const int64_t count = (int64_t)g * kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < nsplit; ++p) s += (double)part[(int64_t)p * count + i];
    const int64_t row = i / kp;
    const int t = (int)(i - row * kp);
    out[(int64_t)t * ld + row] = s;
  }
}

// ---- Gram of a factor: out[a][b] = sum_r F[r*rs + a*cs] * F[r*rs + b*cs] -----------------------
constexpr int GRAM_TILE = 64;     // rows per smem tile
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_BLOCKS = 128;

__global__ void __launch_bounds__(GRAM_THREADS)
gram_partial_kernel(const double* __restrict__ F, int64_t rs, int64_t cs, int64_t rows, int k, int kp,
                    double* __restrict__ partials, const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  extern __shared__ double tile[];  // GRAM_TILE x (k+1)
  const int kk = k * k;
  const int ldt = k + 1;
  // each thread owns outputs o = threadIdx.x, +GRAM_THREADS, ... (k*k <= 4096 -> <= 16 each)
  double acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = 0.0;

  for (int64_t r0 = (int64_t)blockIdx.x * GRAM_TILE; r0 < rows; r0 += (int64_t)gridDim.x * GRAM_TILE) {
    const int nr = (int)min((int64_t)GRAM_TILE, rows - r0);
    __syncthreads();
    if (cs == 1) {  // row-major factor: consecutive threads walk along the components
      for (int i = threadIdx.x; i < nr * k; i += GRAM_THREADS) {
        const int r = i / k, a = i - r * k;
        tile[r * ldt + a] = F[(r0 + r) * rs + a];
      }
    } else {  // component-major factor (k x ld): consecutive threads walk along the rows
      for (int i = threadIdx.x; i < nr * k; i += GRAM_THREADS) {
        const int a = i / nr, r = i - a * nr;
        tile[r * ldt + a] = F[(r0 + r) * rs + (int64_t)a * cs];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int o = threadIdx.x + q * GRAM_THREADS;
      if (o < kk) {
        const int a = o / k, b = o - a * k;
        double s = acc[q];
        for (int r = 0; r < nr; ++r) s = fma(tile[r * ldt + a], tile[r * ldt + b], s);
        acc[q] = s;
      }
    }
  }
  double* out = partials + (int64_t)blockIdx.x * kp * kp;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int o = threadIdx.x + q * GRAM_THREADS;
    if (o < kk) {
      const int a = o / k, b = o - a * k;
      out[a * kp + b] = acc[q];
    }
  }
}

// one warp per output element: lanes stride over the block partials (fixed order), butterfly sum
__global__ void __launch_bounds__(256)
gram_final_kernel(const double* __restrict__ partials, int nblk, int k, int kp, double* __restrict__ out,
                  const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= kp * kp) return;
  const int a = o / kp, b = o - a * kp;
  double s = 0.0;
  if (a < k && b < k)
    for (int p = lane; p < nblk; p += 32) s += partials[(int64_t)p * kp * kp + o];
  s = warp_sum(s);
  if (lane == 0) out[o] = s;
}

// ---- factor update ---------------------------------------------------------------------------
// One thread per factor row r (a gene of G, or a cell of C).  Master F64 and the products `num` are
// component-major (kp x ld, element [t*ld + r]) so consecutive threads read consecutive addresses.
// The row's k entries live in shared memory as sw[t * UPD_THREADS + tid] so that the CD sweep can
// index them by the (uniform) permuted component without spilling; the k x k Gram is broadcast from
// shared memory.  Outputs: master, fp32 row-major shadow S32 (rows x kp, what the streaming sweep
// broadcasts from smem) and optionally the fp32 component-major shadow T32 (kp x ld).
template <int SOLVER>
__global__ void __launch_bounds__(UPD_THREADS)
update_kernel(double* __restrict__ F64, int64_t ld, int64_t rows, int k, int kp,
              const double* __restrict__ num, const float* __restrict__ num_parts, int nsplit,
              const double* __restrict__ gram, double l1, double l2,
              const int32_t* __restrict__ perms, int perm_stride, int perm_offset,
              float* __restrict__ S32, float* __restrict__ T32, float* __restrict__ TC32,
              double* __restrict__ viol_part, double* __restrict__ gram_part_out, NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  extern __shared__ double smem[];
  double* sg = smem;                         // kp*kp Gram
  double* sw = sg + kp * kp;             // k x UPD_LD factor entries
  double* sn = sw + (size_t)k * UPD_LD;  // k x UPD_LD numerators
  __shared__ double scratch[32];
  __shared__ int sperm[SCRNA_MAX_K];
  __shared__ int snan;

  const int tid = threadIdx.x;
  for (int o = tid; o < kp * kp; o += UPD_THREADS) sg[o] = gram[o];
  if (SOLVER == SCRNA_SOLVER_CD && tid < k) {
    const int it = ctrl ? ctrl->iter : 0;
    sperm[tid] = perms ? perms[((int64_t)it * perm_stride + perm_offset) * kp + tid] : tid;
  }
  if (tid == 0) snan = 0;
  __syncthreads();

  const int64_t r = (int64_t)blockIdx.x * UPD_THREADS + tid;
  const bool live = r < rows;
  double viol = 0.0;
  bool nan_seen = false;
  if (live) {
    for (int t = 0; t < k; ++t) sw[t * UPD_LD + tid] = F64[(int64_t)t * ld + r];
    if (num_parts) {
      // fixed-order fp64 sum of the sweep's split partials (kp x ld fp32 each); eight components at a
      // time so that 8 x (unrolled splits) independent loads are in flight per thread
      const int64_t stride = (int64_t)kp * ld;
      for (int t0 = 0; t0 < k; t0 += 8) {
        double a8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a8[u] = 0.0;
        const float* base = num_parts + (int64_t)t0 * ld + r;
#pragma unroll 4
        for (int p = 0; p < nsplit; ++p) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (t0 + u < k) a8[u] += (double)base[(int64_t)p * stride + (int64_t)u * ld];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (t0 + u < k) sn[(t0 + u) * UPD_LD + tid] = a8[u];
      }
    } else {
      for (int t = 0; t < k; ++t) sn[t * UPD_LD + tid] = num[(int64_t)t * ld + r];
    }
    if (SOLVER == SCRNA_SOLVER_MU) {
      // W *= XHt / (W.HHt + l1 + l2*W); the denominator uses the OLD row throughout.  Eight components at
      // a time: eight independent fp64 chains per thread, the Gram row broadcast from shared memory.
      for (int t0 = 0; t0 < k; t0 += 8) {
        double den[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) den[u] = 0.0;
        for (int q = 0; q < k; ++q) {
          const double wq = sw[q * UPD_LD + tid];
          const double* gq = sg + q * kp + t0;  // kp is a multiple of 4 and the Gram is zero-padded to kp x kp
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (t0 + u < kp) den[u] += wq * gq[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + u;
          if (t < k) {
            const double w = sw[t * UPD_LD + tid];
            double d = den[u];
            if (l1 > 0.0) d += l1;
            if (l2 > 0.0) d = d + l2 * w;
            if (d == 0.0) d = (double)kEpsMU;
            sn[t * UPD_LD + tid] = w * (sn[t * UPD_LD + tid] / d);  // numerator slot := new value
          }
        }
      }
      for (int t = 0; t < k; ++t) sw[t * UPD_LD + tid] = sn[t * UPD_LD + tid];
    } else {
      // _update_cdnmf_fast with HHt.diag += l2 and XHt -= l1 (_nmf.py:383-388)
      for (int s = 0; s < k; ++s) {
        const int t = sperm[s];
        double grad = -(sn[t * UPD_LD + tid] - l1);
        for (int q = 0; q < k; ++q) {
          double h = sg[t * kp + q];
          if (q == t) h += l2;
          grad += h * sw[q * UPD_LD + tid];
        }
        const double w = sw[t * UPD_LD + tid];
        const double pg = (w == 0.0) ? fmin(0.0, grad) : grad;
        viol += fabs(pg);
        const double hess = sg[t * kp + t] + l2;
        if (hess != 0.0) sw[t * UPD_LD + tid] = fmax(w - grad / hess, 0.0);
      }
    }
    for (int t = 0; t < k; ++t) {
      const double v = sw[t * UPD_LD + tid];
      nan_seen |= (v != v);
      const int64_t e = (int64_t)t * ld + r;
      F64[e] = v;
      if (T32) T32[e] = (float)v;
    }
    if (TC32) {
      // [hi | lo] TF32 operand shadow of the tensor-core sweep (nmf_sweep_tc.cu): [(r/4)][n < 2kp][r%4]
      float* base = TC32 + ((r >> 2) * 2 * kp) * 4 + (r & 3);
      for (int t = 0; t < k; ++t) {
        const double v = sw[t * UPD_LD + tid];
        const float h = __uint_as_float(__float_as_uint((float)v) & 0xffffe000u);
        base[t * 4] = h;
        base[(kp + t) * 4] = __uint_as_float(__float_as_uint((float)(v - (double)h)) & 0xffffe000u);
      }
    }
    if (S32) {
      float* srow = S32 + r * kp;
      for (int t = 0; t < kp; t += 4) {
        float4 o;
        o.x = (t + 0 < k) ? (float)sw[(t + 0) * UPD_LD + tid] : 0.f;
        o.y = (t + 1 < k) ? (float)sw[(t + 1) * UPD_LD + tid] : 0.f;
        o.z = (t + 2 < k) ? (float)sw[(t + 2) * UPD_LD + tid] : 0.f;
        o.w = (t + 3 < k) ? (float)sw[(t + 3) * UPD_LD + tid] : 0.f;
        *reinterpret_cast<float4*>(srow + t) = o;
      }
    }
  }
  if (nan_seen) snan = 1;
  if (gram_part_out && !live)
    for (int t = 0; t < k; ++t) sw[t * UPD_LD + tid] = 0.0;
  const double bsum = block_sum(viol, scratch);  // contains __syncthreads
  if (tid == 0) {
    if (viol_part) viol_part[blockIdx.x] = bsum;
    if (snan && ctrl) atomicOr(&ctrl->flags, SCRNA_FLAG_NAN);
  }
  if (gram_part_out) {
    // Gram of this block's NEW rows: the next half step needs F^T F (K4 fused into the epilogue)
    // upper triangle only (mirrored), four independent chains over the block's rows
    double* gp = gram_part_out + (int64_t)blockIdx.x * kp * kp;
    for (int o = tid; o < kp * kp; o += UPD_THREADS) {
      const int a = o / kp, b = o - a * kp;
      if (a >= k || b >= k) {
        gp[o] = 0.0;
      } else if (a <= b) {
        const double* ra = sw + a * UPD_LD;
        const double* rb = sw + b * UPD_LD;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int q = 0; q < UPD_THREADS; q += 4) {
          a0 = fma(ra[q], rb[q], a0);
          a1 = fma(ra[q + 1], rb[q + 1], a1);
          a2 = fma(ra[q + 2], rb[q + 2], a2);
          a3 = fma(ra[q + 3], rb[q + 3], a3);
        }
        const double acc = (a0 + a1) + (a2 + a3);
        gp[o] = acc;
        gp[b * kp + a] = acc;
      }
    }
  }
}

// ---- MU update, tiled: 32 rows x 8 component groups per CTA -----------------------------------------------
// The row-per-thread kernel above keeps the CD sweep's sequential order; for the multiplicative update nothing
// is sequential, and at rank 32-64 (136 KB of shared memory, four warps per SM) or on a small shard (fewer
// CTAs than SMs) it was latency bound: 0.4 ms per half step at k=50.  Here a CTA owns tiles of 32 rows, thread
// (r, grp) owns CPT = ceil(k/8) components of row r: the split-partial sums, the k x k denominators (the old
// row broadcast from shared memory) and the writes are spread over 8x more threads, ~60 KB of shared memory
// leaves three CTAs per SM, and the Gram of the new rows accumulates in registers across the CTA's tiles.
constexpr int UT_ROWS = 32;
constexpr int UT_GROUPS = 8;
constexpr int UT_THREADS = UT_ROWS * UT_GROUPS;
constexpr int UT_LD = UT_ROWS + 1;
constexpr int UT_MAXCPT = SCRNA_MAX_K / UT_GROUPS;           // 8
constexpr int UT_MAXG = SCRNA_MAX_K * SCRNA_MAX_K / UT_THREADS;  // 16 Gram entries per thread

__global__ void __launch_bounds__(UT_THREADS)
update_mu_tiled_kernel(double* __restrict__ F64, int64_t ld, int64_t rows, int k, int kp,
                       const double* __restrict__ num, const float* __restrict__ num_parts, int nsplit,
                       const double* __restrict__ gram, double l1, double l2, float* __restrict__ S32,
                       float* __restrict__ T32, float* __restrict__ TC32, double* __restrict__ viol_part,
                       double* __restrict__ gram_part_out, NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  extern __shared__ double smem[];
  double* sg = smem;                      // kp x kp Gram of the other factor
  double* sw = sg + kp * kp;              // k x UT_LD old row entries
  double* sv = sw + (size_t)k * UT_LD;    // k x UT_LD new row entries
  __shared__ int snan;
  const int tid = threadIdx.x, r = tid & (UT_ROWS - 1), grp = tid >> 5;
  const int cpt = (k + UT_GROUPS - 1) / UT_GROUPS;
  const int t_begin = grp * cpt, t_end = min(k, t_begin + cpt);
  for (int o = tid; o < kp * kp; o += UT_THREADS) sg[o] = gram[o];
  if (tid == 0) snan = 0;
  double gacc[UT_MAXG];
#pragma unroll
  for (int i = 0; i < UT_MAXG; ++i) gacc[i] = 0.0;
  bool nan_seen = false;
  const int64_t ntiles = (rows + UT_ROWS - 1) / UT_ROWS;
  const int64_t stride = (int64_t)kp * ld;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = tile * UT_ROWS + r;
    const bool live = row < rows;
    __syncthreads();  // previous tile's Gram accumulation has read sv
    double nv[UT_MAXCPT];
#pragma unroll
    for (int u = 0; u < UT_MAXCPT; ++u) nv[u] = 0.0;
    for (int u = 0; u < UT_MAXCPT; ++u) {
      const int t = t_begin + u;
      if (u < cpt && t < t_end) {
        const int64_t e = (int64_t)t * ld + row;
        sw[t * UT_LD + r] = live ? F64[e] : 0.0;
        if (live) {
          if (num_parts) {  // fixed-order fp64 sum of the sweep's split partials
            double a = 0.0;
#pragma unroll 4
            for (int p = 0; p < nsplit; ++p) a += (double)num_parts[(int64_t)p * stride + e];
            nv[u] = a;
          } else {
            nv[u] = num[e];
          }
        }
      }
    }
    __syncthreads();
    // F *= num / (F.gram + l1 + l2*F): denominators of this thread's components, q in order
    double den[UT_MAXCPT];
#pragma unroll
    for (int u = 0; u < UT_MAXCPT; ++u) den[u] = 0.0;
    for (int q = 0; q < k; ++q) {
      const double wq = sw[q * UT_LD + r];
      const double* gq = sg + q * kp + t_begin;
#pragma unroll
      for (int u = 0; u < UT_MAXCPT; ++u)
        if (u < cpt && t_begin + u < t_end) den[u] += wq * gq[u];
    }
#pragma unroll
    for (int u = 0; u < UT_MAXCPT; ++u) {
      const int t = t_begin + u;
      if (u < cpt && t < t_end) {
        const double w = sw[t * UT_LD + r];
        double d = den[u];
        if (l1 > 0.0) d += l1;
        if (l2 > 0.0) d = d + l2 * w;
        if (d == 0.0) d = (double)kEpsMU;
        const double v = live ? w * (nv[u] / d) : 0.0;
        sv[t * UT_LD + r] = v;
        if (live) {
          nan_seen |= (v != v);
          const int64_t e = (int64_t)t * ld + row;
          F64[e] = v;
          if (T32) T32[e] = (float)v;
          if (S32) S32[row * kp + t] = (float)v;
          if (TC32) {
            float* base = TC32 + ((row >> 2) * 2 * kp) * 4 + (row & 3);
            const float h = __uint_as_float(__float_as_uint((float)v) & 0xffffe000u);
            base[t * 4] = h;
            base[(kp + t) * 4] = __uint_as_float(__float_as_uint((float)(v - (double)h)) & 0xffffe000u);
          }
        }
      }
    }
    if (live && S32)  // padding components of the row-major shadow stay zero
      for (int t = k + grp; t < kp; t += UT_GROUPS) S32[row * kp + t] = 0.f;
    __syncthreads();
    if (gram_part_out) {
      // Gram of the tile's new rows, entries o = tid, tid + 256, ... (rows of sv are conflict-free by pitch)
#pragma unroll
      for (int i = 0; i < UT_MAXG; ++i) {
        const int o = tid + i * UT_THREADS;
        if (o < kp * kp) {
          const int a = o / kp, b = o - a * kp;
          if (a < k && b < k) {
            const double* ra = sv + a * UT_LD;
            const double* rb = sv + b * UT_LD;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
            for (int q = 0; q < UT_ROWS; q += 2) {
              a0 = fma(ra[q], rb[q], a0);
              a1 = fma(ra[q + 1], rb[q + 1], a1);
            }
            gacc[i] += a0 + a1;
          }
        }
      }
    }
  }
  if (nan_seen) snan = 1;
  __syncthreads();
  if (tid == 0) {
    if (viol_part) viol_part[blockIdx.x] = 0.0;
    if (snan && ctrl) atomicOr(&ctrl->flags, SCRNA_FLAG_NAN);
  }
  if (gram_part_out) {
    double* gp = gram_part_out + (int64_t)blockIdx.x * kp * kp;
#pragma unroll
    for (int i = 0; i < UT_MAXG; ++i) {
      const int o = tid + i * UT_THREADS;
      if (o < kp * kp) gp[o] = gacc[i];
    }
  }
}

// fp64 component-major master (kp x ld) -> fp32 row-major shadow (rows x kp) [+ component-major shadow]
__global__ void shadows_kernel(const double* __restrict__ F64, int64_t ld, int64_t rows, int k, int kp,
                               float* __restrict__ S32, float* __restrict__ T32) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  for (int t = 0; t < kp; ++t) {
    const float v = t < k ? (float)F64[(int64_t)t * ld + r] : 0.f;
    if (S32) S32[r * kp + t] = v;
    if (T32 && t < k) T32[(int64_t)t * ld + r] = v;
  }
}

// dst[c][r] = src[r][c] (fp32), dst padding columns [rows, ld_dst) zeroed by the caller
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int64_t cols,
                 float* __restrict__ dst, int64_t ld_dst) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;
    if (c < cols && r < rows) dst[c * ld_dst + r] = tile[tx][ty + i];
  }
}

__global__ void iter_end_kernel(NmfCtrl* ctrl, const double* __restrict__ va, int na,
                                const double* __restrict__ vb, int nb, double tol, int max_iter) {
  if (ctrl->done) return;
  // fixed-order sums, one warp
  double s = 0.0;
  const int lane = threadIdx.x;
  for (int i = lane; i < na; i += 32) s += va[i];
  double a = warp_sum(s);
  s = 0.0;
  for (int i = lane; i < nb; i += 32) s += vb[i];
  double b = warp_sum(s);
  if (lane == 0) {
    const double v = a + b;
    const int it = ctrl->iter + 1;
    ctrl->iter = it;
    if (tol >= 0.0) {
      ctrl->viol_last = v;
      if (it == 1) ctrl->viol_init = v;
      const double vi = ctrl->viol_init;
      if (vi == 0.0 || v / vi <= tol) {
        ctrl->done = 1;
        ctrl->n_iter = it;
      }
    }
    if (!ctrl->done && it >= max_iter) {
      ctrl->done = 1;
      ctrl->n_iter = it;
    }
  }
}

__global__ void ctrl_reset_kernel(NmfCtrl* ctrl) {
  if (threadIdx.x == 0) {
    ctrl->iter = 0; ctrl->done = 0; ctrl->n_iter = 0; ctrl->flags = 0;
    ctrl->viol_init = 0.0; ctrl->viol_last = 0.0; ctrl->err_last = 0.0;
    ctrl->reserved[0] = ctrl->reserved[1] = ctrl->reserved[2] = 0.0;
  }
}

__global__ void cast_kernel(const double* __restrict__ src, float* __restrict__ dst, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

__global__ void convert_pad_kernel(const double* __restrict__ src, int64_t ld_src, float* __restrict__ dst,
                                   int64_t ld_dst, int64_t rows, int64_t cols) {
  const int64_t r = blockIdx.y;
  if (r >= rows) return;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ld_dst;
       c += (int64_t)gridDim.x * blockDim.x)
    dst[r * ld_dst + c] = c < cols ? (float)src[r * ld_src + c] : 0.f;
}

static int grid_for(int64_t count, int threads) {
  int64_t b = ceil_div(count, threads);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_abi_version(void) { return 2; }

extern "C" int scrna_device_info(int* sm_count, int* max_smem_per_block, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SCRNA_ERR_CUDA;
  int v = 0;
  if (sm_count) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sm_count = v; }
  if (max_smem_per_block) { cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); *max_smem_per_block = v; }
  if (cc_major) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); *cc_major = v; }
  if (cc_minor) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); *cc_minor = v; }
  return last_launch_status();
}

extern "C" int scrna_nmf_reduce_rows(const float* part, int nsplit, int g, int kp, double* out, int64_t ld,
                                     const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!part || !out || nsplit <= 0 || g <= 0 || kp <= 0 || ld < g) return SCRNA_ERR_BADARG;
  const int64_t count = (int64_t)g * kp;
  reduce_rows_t_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(
      part, nsplit, g, kp, out, ld, (const NmfCtrl*)ctrl);
  return last_launch_status();
}

extern "C" int scrna_nmf_reduce_cols(const float* part, int nsplit, int kp, int64_t ldc, int64_t ncols,
                                     double* out, const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!part || !out || nsplit <= 0 || kp <= 0 || ldc <= 0 || ncols > ldc) return SCRNA_ERR_BADARG;
  const int64_t count = (int64_t)kp * ldc;
  reduce_partials_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(
      part, nsplit, count, count, out, (const NmfCtrl*)ctrl);
  return last_launch_status();
}

extern "C" int scrna_nmf_gram_blocks(void) { return GRAM_BLOCKS; }

extern "C" int scrna_nmf_gram(const double* F, int64_t rs, int64_t cs, int64_t rows, int k, int kp,
                              double* partials, double* out, const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!F || !partials || !out || rows <= 0 || k <= 0 || k > kp || kp > SCRNA_MAX_K) return SCRNA_ERR_BADARG;
  int nblk = (int)ceil_div(rows, GRAM_TILE);
  if (nblk > GRAM_BLOCKS) nblk = GRAM_BLOCKS;
  const size_t smem = (size_t)GRAM_TILE * (k + 1) * sizeof(double);
  cudaStream_t st = (cudaStream_t)stream;
  gram_partial_kernel<<<nblk, GRAM_THREADS, smem, st>>>(F, rs, cs, rows, k, kp, partials, (const NmfCtrl*)ctrl);
  gram_final_kernel<<<(unsigned)ceil_div(kp * kp, 8), 256, 0, st>>>(partials, nblk, k, kp, out, (const NmfCtrl*)ctrl);
  return last_launch_status();
}

static int tiled_blocks(int64_t rows) {
  int64_t b = ceil_div(rows, UT_ROWS);
  const int64_t cap = 3 * (int64_t)num_sms();
  return (int)(b < cap ? b : cap);
}

// number of per-block slots (violation sums, Gram partials) an update launch may write
extern "C" int scrna_nmf_update_blocks(int64_t rows) {
  const int a = (int)ceil_div(rows, UPD_THREADS), b = tiled_blocks(rows);
  return a > b ? a : b;
}

extern "C" int scrna_nmf_update(int solver, double* F64, int64_t ld, int64_t rows, int k, int kp,
                                const double* num, const float* num_parts, int nsplit, const double* gram,
                                double l1, double l2, const int32_t* perms, int perm_stride, int perm_offset,
                                float* S32, float* T32, float* TC32, void* Fh, float* inv_scale, int kq, int parts,
                                double* viol_part, double* gram_partials, double* gram_out, scrna_nmf_ctrl_t* ctrl,
                                void* stream) {
  if (!F64 || (!num && !num_parts) || (num_parts && nsplit <= 0) || !gram || rows <= 0 || ld < rows || k <= 0 ||
      k > kp || kp > SCRNA_MAX_K || (kp & 3) || (gram_out && !gram_partials) || (Fh && (!gram_partials || !inv_scale)))
    return SCRNA_ERR_BADARG;
  const size_t smem = ((size_t)kp * kp + 2 * (size_t)k * UPD_LD) * sizeof(double);
  int nblk = (int)ceil_div(rows, UPD_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(update_mu_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(update_kernel<SCRNA_SOLVER_MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(update_kernel<SCRNA_SOLVER_CD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  if (solver == SCRNA_SOLVER_MU && kp < 48 && rows >= 2 * (int64_t)num_sms() * UPD_THREADS) {
    // ranks up to 32 on long factors: the row-per-thread kernel has more loads in flight per thread and measured
    // faster; short factors (the replicated gene factor, a cell shard on 4-8 GPUs) leave it with fewer CTAs than
    // SMs and go to the tiled kernel below
    update_kernel<SCRNA_SOLVER_MU><<<nblk, UPD_THREADS, smem, st>>>(F64, ld, rows, k, kp, num, num_parts, nsplit, gram,
                                                                  l1, l2, perms, perm_stride, perm_offset, S32, T32,
                                                                  TC32, viol_part, gram_partials, (NmfCtrl*)ctrl);
  } else if (solver == SCRNA_SOLVER_MU) {
    nblk = tiled_blocks(rows);
    const size_t smem_t = ((size_t)kp * kp + 2 * (size_t)k * UT_LD) * sizeof(double);
    update_mu_tiled_kernel<<<nblk, UT_THREADS, smem_t, st>>>(F64, ld, rows, k, kp, num, num_parts, nsplit, gram, l1, l2,
                                                            S32, T32, TC32, viol_part, gram_partials, (NmfCtrl*)ctrl);
  } else if (solver == SCRNA_SOLVER_CD)
    update_kernel<SCRNA_SOLVER_CD><<<nblk, UPD_THREADS, smem, st>>>(F64, ld, rows, k, kp, num, num_parts, nsplit, gram,
                                                                  l1, l2, perms, perm_stride, perm_offset, S32, T32,
                                                                  TC32, viol_part, gram_partials, (NmfCtrl*)ctrl);
  else
    return SCRNA_ERR_BADARG;
  if (Fh) {
    // fp16 operand shadow of the 16-bit sweep: needs the per-component scale, i.e. the Gram diagonal of the rows
    // just written -- the Gram finish is folded into the same launch (nmf_sweep_h.cu)
    if (int rc = last_launch_status()) return rc;
    return scrna_nmf_h_shadow(F64, ld, rows, k, kq, parts, gram_partials, nblk, kp, gram_out, Fh, inv_scale, ctrl,
                              stream);
  }
  if (gram_out) {
    const int warps = kp * kp;
    gram_final_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(gram_partials, nblk, k, kp, gram_out,
                                                                   (const NmfCtrl*)ctrl);
  }
  return last_launch_status();
}

extern "C" int scrna_nmf_shadows(const double* F64, int64_t ld, int64_t rows, int k, int kp, float* S32,
                                 float* T32, void* stream) {
  if (!F64 || rows <= 0 || ld < rows || k <= 0 || k > kp) return SCRNA_ERR_BADARG;
  shadows_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(F64, ld, rows, k, kp, S32, T32);
  return last_launch_status();
}

extern "C" int scrna_transpose_f32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* dst,
                                   int64_t ld_dst, void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || cols > ld_src || rows > ld_dst) return SCRNA_ERR_BADARG;
  const int64_t by = ceil_div(rows, 32);
  if (by > 65535) return SCRNA_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)by);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, dst, ld_dst);
  return last_launch_status();
}

extern "C" int scrna_nmf_iter_end(scrna_nmf_ctrl_t* ctrl, const double* viol_a, int na, const double* viol_b,
                                  int nb, double tol, int max_iter, void* stream) {
  if (!ctrl) return SCRNA_ERR_BADARG;
  iter_end_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((NmfCtrl*)ctrl, viol_a, viol_a ? na : 0, viol_b,
                                                      viol_b ? nb : 0, tol, max_iter);
  return last_launch_status();
}

extern "C" int scrna_nmf_ctrl_reset(scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!ctrl) return SCRNA_ERR_BADARG;
  ctrl_reset_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((NmfCtrl*)ctrl);
  return last_launch_status();
}

extern "C" int scrna_cast_f64_f32(const double* src, float* dst, int64_t count, void* stream) {
  if (!src || !dst || count < 0) return SCRNA_ERR_BADARG;
  if (count == 0) return SCRNA_OK;
  cast_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, count);
  return last_launch_status();
}

extern "C" int scrna_convert_pad_f64_f32(const double* src, int64_t ld_src, float* dst, int64_t ld_dst,
                                         int64_t rows, int64_t cols, void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || cols > ld_src || cols > ld_dst || rows > 65535 * 64LL)
    return SCRNA_ERR_BADARG;
  // blockIdx.y is limited to 65535: process row bands
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    int bx = (int)ceil_div(ld_dst, 256);
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)nr);
    convert_pad_kernel<<<grid, 256, 0, st>>>(src + r0 * ld_src, ld_src, dst + r0 * ld_dst, ld_dst, nr, cols);
  }
  return last_launch_status();
}
