// NMF X-sweeps for sm_100a: the two HBM-bound passes of one NMF iteration.
//
//   K1  xht_partial : part[s][i][t] = sum_{j in span s} X[i][j] * C32[t][j]     (X . C^T)
//   K3  wtx_partial : part[s][t][j] = sum_{i in span s} G32[i][t] * X[i][j]     (G^T . X)
//
// Both stream X (genes x cells, row-major fp32) exactly once with coalesced 128-bit loads that
// bypass L1 allocation; the rank-k factor is the only re-used operand and stays on chip
// (registers / shared memory / L1).  Algorithmic bytes per launch = g * ldx * 4 (+ factor + partials).
// Replaces the two BLAS products of sklearn's _update_coordinate_descent (_nmf.py:380-381) and of
// _multiplicative_update_w/_h (_nmf.py:537-538, 644) that the reference reaches through
// decomp.NMF(...).fit_transform (scRNA/nmf_clustering.py:25-29).
#include "common.cuh"

namespace scrna {

// -------------------------------------------------------------------------------------------
// K1: lanes run along cells (contiguous axis), each warp owns R consecutive gene rows and keeps
// R x KP accumulators per lane; the cross-lane reduction happens once, at the end of the span.
// -------------------------------------------------------------------------------------------
template <int KP, int C, int R>
__global__ void __launch_bounds__(128, (KP <= 16) ? 3 : 2)
xht_partial_kernel(const float* __restrict__ X, int64_t ldx, int g, int64_t ncols,
                   const float* __restrict__ C32, int64_t ldc, float* __restrict__ part,
                   int64_t chunks_per_split, int64_t total_chunks, const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  constexpr int CHUNK = 32 * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 4 + warp) * R;
  if (row0 >= g) return;  // warp-uniform, no block barriers in this kernel

  const int64_t ch_begin = (int64_t)blockIdx.y * chunks_per_split;
  int64_t ch_end = ch_begin + chunks_per_split;
  if (ch_end > total_chunks) ch_end = total_chunks;

  const float* xrow[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int row = row0 + r;
    if (row > g - 1) row = g - 1;  // clamped rows are computed but never written
    xrow[r] = X + (int64_t)row * ldx;
  }

  float acc[R][KP];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < KP; ++t) acc[r][t] = 0.f;

  Vec<C> xa[R];
  int64_t j = ch_begin * CHUNK + (int64_t)lane * C;
  bool valid = (ch_begin < ch_end) && (j < ncols);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (valid) xa[r] = Vec<C>::stream(xrow[r] + j);
    else
#pragma unroll
      for (int c = 0; c < C; ++c) xa[r].v[c] = 0.f;
  }

  for (int64_t ch = ch_begin; ch < ch_end; ++ch) {
    // prefetch the next chunk of X while this one is consumed
    Vec<C> xn[R];
    const int64_t jn = j + CHUNK;
    const bool valid_n = (ch + 1 < ch_end) && (jn < ncols);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (valid_n) xn[r] = Vec<C>::stream(xrow[r] + jn);
      else
#pragma unroll
        for (int c = 0; c < C; ++c) xn[r].v[c] = 0.f;
    }
    if (valid) {
#pragma unroll
      for (int t = 0; t < KP; ++t) {
        const Vec<C> ct = Vec<C>::cached(C32 + (int64_t)t * ldc + j);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < C; ++c) acc[r][t] = fmaf(xa[r].v[c], ct.v[c], acc[r][t]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) xa[r] = xn[r];
    j = jn;
    valid = valid_n;
  }

  // one butterfly per accumulator, once per warp and span
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < KP; ++t) acc[r][t] = warp_sum(acc[r][t]);

  if (lane == 0) {
    float* out = part + ((int64_t)blockIdx.y * g + row0) * KP;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < g) {
#pragma unroll
        for (int t = 0; t < KP; t += 4)
          *reinterpret_cast<float4*>(out + r * KP + t) =
              make_float4(acc[r][t], acc[r][t + 1], acc[r][t + 2], acc[r][t + 3]);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// K3: each lane owns C cells and KP x C accumulators; the gene-factor rows of the CTA's gene span
// are staged in shared memory and broadcast to all lanes.
// -------------------------------------------------------------------------------------------
template <int KP, int C, int U, bool F2>
__global__ void __launch_bounds__(256, 2)
wtx_partial_kernel(const float* __restrict__ X, int64_t ldx, int g, int64_t ncols,
                   const float* __restrict__ G32, float* __restrict__ part, int64_t ldc,
                   int rows_per_split, int tile_rows, const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  extern __shared__ float4 smem4[];
  const float* sW = reinterpret_cast<const float*>(smem4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j0 = ((int64_t)blockIdx.x * 8 + warp) * (32 * C) + (int64_t)lane * C;
  const bool active = j0 < ncols;
  const int r_begin = blockIdx.y * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > g) r_end = g;

  // accumulators as component pairs (t, t+1): with F2 one packed fma.rn.f32x2 updates both
  float2 acc[KP / 2][C];
#pragma unroll
  for (int t = 0; t < KP / 2; ++t)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[t][c] = make_float2(0.f, 0.f);

  for (int t0 = r_begin; t0 < r_end; t0 += tile_rows) {
    const int nr = min(tile_rows, r_end - t0);
    __syncthreads();  // previous tile fully consumed
    {
      const float4* src = reinterpret_cast<const float4*>(G32 + (int64_t)t0 * KP);
      const int n4 = nr * (KP / 4);
      for (int i = threadIdx.x; i < n4; i += 256) smem4[i] = __ldg(src + i);
    }
    __syncthreads();
    if (!active) continue;
    const float* xr = X + (int64_t)t0 * ldx + j0;

    Vec<C> xa[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u < nr) xa[u] = Vec<C>::stream(xr + (int64_t)u * ldx);
      else
#pragma unroll
        for (int c = 0; c < C; ++c) xa[u].v[c] = 0.f;
    }
    for (int r = 0; r < nr; r += U) {
      Vec<C> xn[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rn = r + U + u;
        if (rn < nr) xn[u] = Vec<C>::stream(xr + (int64_t)rn * ldx);
        else
#pragma unroll
          for (int c = 0; c < C; ++c) xn[u].v[c] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // rows past the tile read zero X (above); clamp the smem row so the index stays in range
        const int rr = (r + u < nr) ? (r + u) : (nr - 1);
        const float4* w4 = reinterpret_cast<const float4*>(sW + rr * KP);
#pragma unroll
        for (int t = 0; t < KP; t += 4) {
          const float4 w = w4[t / 4];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float x = xa[u].v[c];
            if constexpr (F2) {
              const float2 xx = make_float2(x, x);
              acc[t / 2][c] = __ffma2_rn(make_float2(w.x, w.y), xx, acc[t / 2][c]);
              acc[t / 2 + 1][c] = __ffma2_rn(make_float2(w.z, w.w), xx, acc[t / 2 + 1][c]);
            } else {
              acc[t / 2][c].x = fmaf(w.x, x, acc[t / 2][c].x);
              acc[t / 2][c].y = fmaf(w.y, x, acc[t / 2][c].y);
              acc[t / 2 + 1][c].x = fmaf(w.z, x, acc[t / 2 + 1][c].x);
              acc[t / 2 + 1][c].y = fmaf(w.w, x, acc[t / 2 + 1][c].y);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) xa[u] = xn[u];
    }
  }

  if (active) {
    float* out = part + ((int64_t)blockIdx.y * KP) * ldc + j0;
#pragma unroll
    for (int t = 0; t < KP; ++t) {
      Vec<C> o;
#pragma unroll
      for (int c = 0; c < C; ++c) o.v[c] = (t & 1) ? acc[t / 2][c].y : acc[t / 2][c].x;
      o.store(out + (int64_t)t * ldc);
    }
  }
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
template <int KP>
struct SweepCfg {
  // cells per lane so that KP * C <= 64 accumulators (K3) and the K1 register tile stays <= 64
  static constexpr int C = KP <= 16 ? 4 : (KP <= 32 ? 2 : 1);
  static constexpr int R = KP <= 8 ? 8 : (KP <= 16 ? 4 : (KP <= 32 ? 2 : 1));  // R*KP <= 64
  static constexpr int CB = KP <= 16 ? 4 : (KP <= 32 ? 2 : 1);                 // K3: KP*CB <= 64
  static constexpr int U = KP <= 8 ? 8 : 4;                                    // rows in flight per lane
};

static int pick_splits(int64_t tiles, int64_t slots, int max_split) {
  // smallest split count whose last wave is >= 90% full, preferring >= 2 waves' worth of CTAs
  if (tiles <= 0) return 1;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_split; ++s) {
    const int64_t ctas = tiles * s;
    const int64_t waves = ceil_div(ctas, slots);
    const double eff = (double)ctas / (double)(waves * slots);
    const double score = eff - (ctas < 2 * slots ? 0.05 : 0.0) - 0.002 * s;
    if (score > best_eff) {
      best_eff = score;
      best = s;
    }
  }
  return best;
}

template <int KP>
static int xht_launch(const float* X, int64_t ldx, int g, int64_t ncols, const float* C32, int64_t ldc,
                      float* part, int nsplit, const NmfCtrl* ctrl, cudaStream_t st) {
  using Cfg = SweepCfg<KP>;
  constexpr int CHUNK = 32 * Cfg::C;
  const int64_t total_chunks = ceil_div(ncols, CHUNK);
  const int64_t cps = ceil_div(total_chunks, nsplit);
  dim3 grid((unsigned)ceil_div(g, 4 * Cfg::R), (unsigned)nsplit);
  xht_partial_kernel<KP, Cfg::C, Cfg::R><<<grid, 128, 0, st>>>(X, ldx, g, ncols, C32, ldc, part, cps,
                                                               total_chunks, ctrl);
  return last_launch_status();
}

template <int KP>
static int wtx_tile_rows(int rows_per_split) {
  int tr = (48 * 1024) / (KP * 4);  // <= 48 KB static-limit-free dynamic smem
  if (tr > rows_per_split) tr = rows_per_split;
  tr = (tr / 8) * 8;
  if (tr < 8) tr = 8;
  return tr;
}

template <int KP>
static int wtx_launch(const float* X, int64_t ldx, int g, int64_t ncols, const float* G32, float* part,
                      int64_t ldc, int nsplit, int variant, const NmfCtrl* ctrl, cudaStream_t st) {
  using Cfg = SweepCfg<KP>;
  const int rps = (int)ceil_div(g, nsplit);
  const int tr = wtx_tile_rows<KP>(rps);
  dim3 grid((unsigned)ceil_div(ncols, 8 * 32 * Cfg::CB), (unsigned)nsplit);
  const size_t smem = (size_t)tr * KP * sizeof(float);
  // variant 0: scalar FFMA; variant >= 1 (default): packed fma.rn.f32x2 (FFMA2), bit-identical results
  if (variant >= 1)
    wtx_partial_kernel<KP, Cfg::CB, Cfg::U, true><<<grid, 256, smem, st>>>(X, ldx, g, ncols, G32, part, ldc, rps,
                                                                          tr, ctrl);
  else
    wtx_partial_kernel<KP, Cfg::CB, Cfg::U, false><<<grid, 256, smem, st>>>(X, ldx, g, ncols, G32, part, ldc, rps,
                                                                           tr, ctrl);
  return last_launch_status();
}

#define SCRNA_DISPATCH_KP(kp, CALL)   \
  switch (kp) {                       \
    case 4: { constexpr int KP = 4; CALL; } break;    \
    case 8: { constexpr int KP = 8; CALL; } break;    \
    case 12: { constexpr int KP = 12; CALL; } break;  \
    case 16: { constexpr int KP = 16; CALL; } break;  \
    case 20: { constexpr int KP = 20; CALL; } break;  \
    case 24: { constexpr int KP = 24; CALL; } break;  \
    case 28: { constexpr int KP = 28; CALL; } break;  \
    case 32: { constexpr int KP = 32; CALL; } break;  \
    case 40: { constexpr int KP = 40; CALL; } break;  \
    case 48: { constexpr int KP = 48; CALL; } break;  \
    case 56: { constexpr int KP = 56; CALL; } break;  \
    case 64: { constexpr int KP = 64; CALL; } break;  \
    default: return SCRNA_ERR_UNSUPPORTED;            \
  }

template <int KP>
static int xht_tiles(int g) {
  return (int)ceil_div(g, 4 * SweepCfg<KP>::R);
}
template <int KP>
static int64_t wtx_tiles(int64_t ncols) {
  return ceil_div(ncols, 8 * 32 * SweepCfg<KP>::CB);
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_nmf_xht_splits(int g, int64_t ncols, int kp) {
  if (g <= 0 || ncols <= 0) return SCRNA_ERR_BADARG;
  int64_t tiles = 0;
  int64_t chunks = 0;
  SCRNA_DISPATCH_KP(kp, { tiles = xht_tiles<KP>(g); chunks = ceil_div(ncols, 32 * SweepCfg<KP>::C); });
  int max_split = (int)(chunks < 64 ? chunks : 64);
  if (max_split < 1) max_split = 1;
  return pick_splits(tiles, (int64_t)num_sms() * 3, max_split);
}

extern "C" int scrna_nmf_wtx_splits(int g, int64_t ncols, int kp) {
  if (g <= 0 || ncols <= 0) return SCRNA_ERR_BADARG;
  int64_t tiles = 0;
  SCRNA_DISPATCH_KP(kp, { tiles = wtx_tiles<KP>(ncols); });
  int max_split = g / 64;
  if (max_split > 64) max_split = 64;
  if (max_split < 1) max_split = 1;
  return pick_splits(tiles, (int64_t)num_sms() * 2, max_split);
}

extern "C" int scrna_nmf_xht_partial(const float* X, int64_t ldx, int g, int64_t ncols, const float* C32,
                                     int64_t ldc, int kp, float* part, int nsplit, int variant,
                                     const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!X || !C32 || !part || g <= 0 || ncols <= 0 || nsplit <= 0 || (ldx & 3) || (ldc & 3) ||
      (ncols & 3) || ncols > ldx || ncols > ldc)
    return SCRNA_ERR_BADARG;
  (void)variant;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const NmfCtrl* c = reinterpret_cast<const NmfCtrl*>(ctrl);
  SCRNA_DISPATCH_KP(kp, return xht_launch<KP>(X, ldx, g, ncols, C32, ldc, part, nsplit, c, st));
  return SCRNA_ERR_UNSUPPORTED;
}

extern "C" int scrna_nmf_wtx_partial(const float* X, int64_t ldx, int g, int64_t ncols, const float* G32,
                                     int kp, float* part, int64_t ldc, int nsplit, int variant,
                                     const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!X || !G32 || !part || g <= 0 || ncols <= 0 || nsplit <= 0 || (ldx & 3) || (ldc & 3) ||
      (ncols & 3) || ncols > ldx || ncols > ldc)
    return SCRNA_ERR_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const NmfCtrl* c = reinterpret_cast<const NmfCtrl*>(ctrl);
  SCRNA_DISPATCH_KP(kp, return wtx_launch<KP>(X, ldx, g, ncols, G32, part, ldc, nsplit, variant, c, st));
  return SCRNA_ERR_UNSUPPORTED;
}
