// Target transfer (DaNmfClustering.get_mixed_data, scRNA/nmf_clustering.py:147-205 and
// utils.get_transferred_data_matrix, scRNA/utils.py:187-230) for sm_100a.
//
//   K5  W^T.X              -> scrna_nmf_wtx_partial + scrna_nmf_reduce_cols (hoisted: constant over
//                             the iterations, the reference recomputes it every iteration, utils.py:201)
//   K6  mu_h_step          -> H *= WtX / ((W^T W).H)   (utils.py:199-201; zero denominators flagged)
//   K7  residual           -> sum |X - W.H|^p in ONE sweep of X without materialising W.H (utils.py:202)
//   K8  mix_gather         -> mixed = mix * W[:, argmax H] + (1-mix) * X   (nmf_clustering.py:185,200;
//                             W.H2 with one-hot H2 is a column gather of W, utils.py:208-210)
//       reject_stats       -> per-cell column sums and L1/L2 residuals vs W.H and W.H2
//                             (nmf_clustering.py:236-271) in one sweep instead of five dense products
#include "common.cuh"

namespace scrna {

// ---- K6 ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
transfer_mu_step_kernel(double* __restrict__ H64, float* __restrict__ H32, const double* __restrict__ WtX,
                        const double* __restrict__ WtW, int k, int kp, int64_t ncols, int64_t ldc,
                        NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  extern __shared__ double sg[];  // kp*kp
  for (int o = threadIdx.x; o < kp * kp; o += blockDim.x) sg[o] = WtW[o];
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  double h[SCRNA_MAX_K];
  bool zero = false;
  for (int t = 0; t < k; ++t) h[t] = H64[(int64_t)t * ldc + j];
  for (int t = 0; t < k; ++t) {
    double den = 0.0;
    for (int q = 0; q < k; ++q) den += sg[t * kp + q] * h[q];
    zero |= (den == 0.0);
    const double v = h[t] * (WtX[(int64_t)t * ldc + j] / den);
    H64[(int64_t)t * ldc + j] = v;
    H32[(int64_t)t * ldc + j] = (float)v;
  }
  if (zero && ctrl) atomicOr(&ctrl->flags, SCRNA_FLAG_ZERO_DEN);
}

// ---- K7 ---------------------------------------------------------------------------------------
// Same streaming structure as K3: each lane owns C cells, keeps their KP x C factor entries in
// registers, and the gene-factor rows are broadcast from shared memory.
template <typename XT, int KP, int C, int POWER>
__global__ void __launch_bounds__(256, 2)
residual_kernel(const XT* __restrict__ X, int64_t ldx, int64_t panel_stride, int g, int64_t ncols,
                const float* __restrict__ G32, const float* __restrict__ C32, int64_t ldc,
                int rows_per_split, int tile_rows, double* __restrict__ partials,
                const NmfCtrl* __restrict__ gate) {
  extern __shared__ float4 smem4[];
  __shared__ double scratch[32];
  if (ctrl_done(gate)) return;
  const float* sW = reinterpret_cast<const float*>(smem4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j0 = ((int64_t)blockIdx.x * 8 + warp) * (32 * C) + (int64_t)lane * C;
  const bool active = j0 < ncols;
  const int r_begin = blockIdx.y * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > g) r_end = g;

  float h[KP][C];
#pragma unroll
  for (int t = 0; t < KP; ++t) {
    if (active) {
      const Vec<C> v = Vec<C>::cached(C32 + (int64_t)t * ldc + j0);
#pragma unroll
      for (int c = 0; c < C; ++c) h[t][c] = v.v[c];
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) h[t][c] = 0.f;
    }
  }

  double total = 0.0;
  constexpr int U = 4;
  for (int t0 = r_begin; t0 < r_end; t0 += tile_rows) {
    const int nr = min(tile_rows, r_end - t0);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(G32 + (int64_t)t0 * KP);
      const int n4 = nr * (KP / 4);
      for (int i = threadIdx.x; i < n4; i += 256) smem4[i] = __ldg(src + i);
    }
    __syncthreads();
    if (!active) continue;
    // row-major: element (i, j) at i*ldx + j; panel-major (panel_stride != 0, ldx = 256): panels of 256 columns
    const XT* xr = X + (int64_t)t0 * ldx + (panel_stride ? (j0 >> 8) * panel_stride + (j0 & 255) : j0);
    float part = 0.f;  // fp32 within a tile (<= a few thousand terms), fp64 across tiles
    for (int r = 0; r < nr; r += U) {
      Vec<C> xa[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u < nr) xa[u] = load_x<C>(xr + (int64_t)(r + u) * ldx);
        else
#pragma unroll
          for (int c = 0; c < C; ++c) xa[u].v[c] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u < nr) {
          const float4* w4 = reinterpret_cast<const float4*>(sW + (r + u) * KP);
          float pred[C];
#pragma unroll
          for (int c = 0; c < C; ++c) pred[c] = 0.f;
#pragma unroll
          for (int t = 0; t < KP; t += 4) {
            const float4 w = w4[t / 4];
#pragma unroll
            for (int c = 0; c < C; ++c) {
              pred[c] = fmaf(w.x, h[t + 0][c], pred[c]);
              pred[c] = fmaf(w.y, h[t + 1][c], pred[c]);
              pred[c] = fmaf(w.z, h[t + 2][c], pred[c]);
              pred[c] = fmaf(w.w, h[t + 3][c], pred[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float d = xa[u].v[c] - pred[c];
            part += (POWER == 1) ? fabsf(d) : d * d;
          }
        }
      }
      if (((r / U) & 63) == 63) { total += (double)part; part = 0.f; }
    }
    total += (double)part;
  }
  const double bsum = block_sum(total, scratch);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = bsum;
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out,
                                    double scale, const NmfCtrl* __restrict__ gate = nullptr) {
  __shared__ double scratch[32];
  if (ctrl_done(gate)) return;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// ---- argmax over the k components of each cell (ties -> lowest index, np.argmax) ---------------
__global__ void argmax_cols_kernel(const double* __restrict__ H64, int k, int64_t ncols, int64_t ldc,
                                   int32_t* __restrict__ labels) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  double best = H64[j];
  int bi = 0;
  for (int t = 1; t < k; ++t) {
    const double v = H64[(int64_t)t * ldc + j];
    if (v > best) { best = v; bi = t; }
  }
  labels[j] = bi;
}

// ---- K8 ---------------------------------------------------------------------------------------
// One thread per (gene, 4 cells): mixed = mix * W[i][label_j] + (1 - mix) * X[i][j].
__global__ void __launch_bounds__(256)
mix_gather_kernel(const float* __restrict__ X, int64_t ldx, int g, int64_t ncols,
                  const double* __restrict__ G64, int kp, const int32_t* __restrict__ labels, double mix,
                  double* __restrict__ mixed64, int64_t ld64, float* __restrict__ mixed32, int64_t ld32,
                  double* __restrict__ newtrg64) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= ncols) return;
  int lb[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) lb[c] = (j + c < ncols) ? labels[j + c] : 0;
  for (int i = blockIdx.y; i < g; i += gridDim.y) {
    const float4 x = ld_stream(reinterpret_cast<const float4*>(X + (int64_t)i * ldx + j));
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const double* wrow = G64 + (int64_t)i * kp;
    float m32[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double w = wrow[lb[c]];
      const double m = mix * w + (1.0 - mix) * (double)xv[c];
      m32[c] = (float)m;
      if (j + c < ncols) {
        if (mixed64) mixed64[(int64_t)i * ld64 + j + c] = m;
        if (newtrg64) newtrg64[(int64_t)i * ld64 + j + c] = w;
      }
    }
    if (mixed32) {
      // padding columns [ncols, ld32) stay zero: X is zero there and label 0 would leak W otherwise
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (j + c >= ncols) m32[c] = 0.f;
      *reinterpret_cast<float4*>(mixed32 + (int64_t)i * ld32 + j) = make_float4(m32[0], m32[1], m32[2], m32[3]);
    }
  }
}

// ---- rejection statistics (nmf_clustering.py:236-271), per cell j:
//   out[0][j] = sum_i X[i][j]
//   out[1][j] = sum_i |X - W.H |      out[2][j] = sum_i (X - W.H )^2
//   out[3][j] = sum_i |X - W.H2|      out[4][j] = sum_i (X - W.H2)^2
// Thread owns one cell; gene-factor rows broadcast from shared memory; fp64 accumulation.
__global__ void __launch_bounds__(128)
reject_stats_kernel(const float* __restrict__ X, int64_t ldx, int g, int64_t ncols,
                    const double* __restrict__ G64, int k, int kp, const double* __restrict__ H64,
                    int64_t ldc, const int32_t* __restrict__ labels, double* __restrict__ out,
                    int64_t ldo) {
  extern __shared__ double sWd[];  // tile_rows * kp
  const int tile_rows = 64;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < ncols;
  double h[SCRNA_MAX_K];
  for (int t = 0; t < k; ++t) h[t] = live ? H64[(int64_t)t * ldc + j] : 0.0;
  const int lb = live ? labels[j] : 0;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int t0 = 0; t0 < g; t0 += tile_rows) {
    const int nr = min(tile_rows, g - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * kp; i += blockDim.x) sWd[i] = G64[(int64_t)t0 * kp + i];
    __syncthreads();
    if (!live) continue;
    for (int r = 0; r < nr; ++r) {
      const double x = (double)X[(int64_t)(t0 + r) * ldx + j];
      const double* w = sWd + r * kp;
      double pred = 0.0;
      for (int t = 0; t < k; ++t) pred += w[t] * h[t];
      const double d1 = x - pred;
      const double d2 = x - w[lb];
      s0 += x;
      s1 += fabs(d1);
      s2 += d1 * d1;
      s3 += fabs(d2);
      s4 += d2 * d2;
    }
  }
  if (live) {
    out[j] = s0;
    out[ldo + j] = s1;
    out[2 * ldo + j] = s2;
    out[3 * ldo + j] = s3;
    out[4 * ldo + j] = s4;
  }
}

template <int KP>
struct ResCfg {
  static constexpr int C = KP <= 8 ? 4 : (KP <= 16 ? 2 : 1);
};

template <typename XT, int KP>
static int residual_launch(const XT* X, int64_t ldx, int64_t panel_stride, int g, int64_t ncols, const float* G32,
                           const float* C32, int64_t ldc, int power, double* partials, double* out,
                           int nsplit, cudaStream_t st, const NmfCtrl* gate = nullptr) {
  constexpr int C = ResCfg<KP>::C;
  const int rps = (int)ceil_div(g, nsplit);
  int tr = (32 * 1024) / (KP * 4);
  if (tr > rps) tr = rps;
  tr = (tr / 4) * 4;
  if (tr < 4) tr = 4;
  dim3 grid((unsigned)ceil_div(ncols, 8 * 32 * C), (unsigned)nsplit);
  const size_t smem = (size_t)tr * KP * sizeof(float);
  if (power == 1)
    residual_kernel<XT, KP, C, 1><<<grid, 256, smem, st>>>(X, ldx, panel_stride, g, ncols, G32, C32, ldc, rps, tr,
                                                           partials, gate);
  else
    residual_kernel<XT, KP, C, 2><<<grid, 256, smem, st>>>(X, ldx, panel_stride, g, ncols, G32, C32, ldc, rps, tr,
                                                           partials, gate);
  sum_partials_kernel<<<1, 256, 0, st>>>(partials, (int)(grid.x * grid.y), out, 1.0, gate);
  return last_launch_status();
}

static int residual_splits(int g, int64_t tiles) {
  int s = (int)ceil_div((int64_t)num_sms() * 2, tiles);
  if (s > g / 64) s = g / 64;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  return s;
}

}  // namespace scrna

using namespace scrna;

#define SCRNA_DISPATCH_KP_T(kp, ...)    \
  switch (kp) {                       \
    case 4: { constexpr int KP = 4; __VA_ARGS__; } break;    \
    case 8: { constexpr int KP = 8; __VA_ARGS__; } break;    \
    case 12: { constexpr int KP = 12; __VA_ARGS__; } break;  \
    case 16: { constexpr int KP = 16; __VA_ARGS__; } break;  \
    case 20: { constexpr int KP = 20; __VA_ARGS__; } break;  \
    case 24: { constexpr int KP = 24; __VA_ARGS__; } break;  \
    case 28: { constexpr int KP = 28; __VA_ARGS__; } break;  \
    case 32: { constexpr int KP = 32; __VA_ARGS__; } break;  \
    case 40: { constexpr int KP = 40; __VA_ARGS__; } break;  \
    case 48: { constexpr int KP = 48; __VA_ARGS__; } break;  \
    case 56: { constexpr int KP = 56; __VA_ARGS__; } break;  \
    case 64: { constexpr int KP = 64; __VA_ARGS__; } break;  \
    default: return SCRNA_ERR_UNSUPPORTED;            \
  }

extern "C" int scrna_residual_blocks(int64_t ncols, int kp) {
  int64_t tiles = 0;
  SCRNA_DISPATCH_KP_T(kp, { tiles = ceil_div(ncols, 8 * 32 * ResCfg<KP>::C); });
  return (int)(tiles * 32);
}

extern "C" int scrna_residual(const float* X, int64_t ldx, int g, int64_t ncols, const float* G32,
                              const float* C32, int64_t ldc, int kp, int power, double* partials,
                              double* out, void* stream) {
  if (!X || !G32 || !C32 || !partials || !out || g <= 0 || ncols <= 0 || (ldx & 3) || (ldc & 3) ||
      (ncols & 3) || ncols > ldx || ncols > ldc || (power != 1 && power != 2))
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  SCRNA_DISPATCH_KP_T(kp, {
    const int64_t tiles = ceil_div(ncols, 8 * 32 * ResCfg<KP>::C);
    return residual_launch<float, KP>(X, ldx, 0, g, ncols, G32, C32, ldc, power, partials, out,
                                      residual_splits(g, tiles), st);
  });
  return SCRNA_ERR_UNSUPPORTED;
}

extern "C" int scrna_residual_h(const void* X16, int64_t rows_pad, int g, int64_t ncols, const float* G32,
                                const float* C32, int64_t ldc, int kp, int power, double* partials,
                                double* out, void* stream) {
  if (!X16 || !G32 || !C32 || !partials || !out || g <= 0 || ncols <= 0 || rows_pad < g || (ldc & 3) ||
      (ncols & 3) || ncols > ldc || (power != 1 && power != 2))
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  SCRNA_DISPATCH_KP_T(kp, {
    const int64_t tiles = ceil_div(ncols, 8 * 32 * ResCfg<KP>::C);
    return residual_launch<__half, KP>(reinterpret_cast<const __half*>(X16), 256, rows_pad * 256, g, ncols, G32, C32,
                                       ldc, power, partials, out, residual_splits(g, tiles), st);
  });
  return SCRNA_ERR_UNSUPPORTED;
}

extern "C" int scrna_sum_f64(const double* v, int n, double* out, double scale, void* stream) {
  if (!v || !out || n <= 0) return SCRNA_ERR_BADARG;
  sum_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(v, n, out, scale);
  return last_launch_status();
}

extern "C" int scrna_transfer_mu_step(double* H64, float* H32, const double* WtX, const double* WtW, int k,
                                      int kp, int64_t ncols, int64_t ldc, scrna_nmf_ctrl_t* ctrl,
                                      void* stream) {
  if (!H64 || !H32 || !WtX || !WtW || k <= 0 || k > kp || kp > SCRNA_MAX_K || ncols <= 0 || ncols > ldc)
    return SCRNA_ERR_BADARG;
  const size_t smem = (size_t)kp * kp * sizeof(double);
  transfer_mu_step_kernel<<<(unsigned)ceil_div(ncols, 128), 128, smem, (cudaStream_t)stream>>>(
      H64, H32, WtX, WtW, k, kp, ncols, ldc, (NmfCtrl*)ctrl);
  return last_launch_status();
}

// The stop rule of utils.get_transferred_data_matrix (utils.py:202-205) on the device: one thread turns the reduced
// residual into the mean absolute error, compares it with the previous one and sets ctrl->done; every kernel of a later
// enqueued iteration (K6, the gated residual, this one) then returns at once, so the host polls once per batch.
__global__ void transfer_stop_kernel(const double* __restrict__ res, double denom, double rel_err, int max_iter,
                                     NmfCtrl* __restrict__ ctrl) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || ctrl->done) return;
  const double new_err = res[0] / denom;
  const int it = ++ctrl->iter;
  const double err = it == 1 ? 1e10 : ctrl->reserved[0];
  ctrl->err_last = new_err;
  bool stop = (ctrl->flags & SCRNA_FLAG_ZERO_DEN) != 0;
  if (!stop && fabs((err - new_err) / err) <= rel_err && err >= new_err) stop = true;
  if (!stop) {
    ctrl->reserved[0] = new_err;
    stop = it >= max_iter;
  }
  if (stop) {
    ctrl->n_iter = it;
    ctrl->done = 1;
  }
}

extern "C" int scrna_transfer_iteration(const void* X, int x_is_f16, int64_t ldx, int g, int64_t n, int64_t ncols,
                                        const float* G32, double* H64, float* H32, const double* WtX,
                                        const double* WtW, int k, int kp, int64_t ldc, double* partials, double* out,
                                        scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!X || !G32 || !H64 || !H32 || !WtX || !WtW || !partials || !out || !ctrl || k <= 0 || k > kp ||
      kp > SCRNA_MAX_K || g <= 0 || n <= 0 || n > ncols || ncols > ldc || (ldc & 3) || (ncols & 3) ||
      (x_is_f16 ? ldx < g : ((ldx & 3) || ncols > ldx)))
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const NmfCtrl* gate = (const NmfCtrl*)ctrl;
  transfer_mu_step_kernel<<<(unsigned)ceil_div(n, 128), 128, (size_t)kp * kp * sizeof(double), st>>>(
      H64, H32, WtX, WtW, k, kp, n, ldc, (NmfCtrl*)ctrl);
  SCRNA_DISPATCH_KP_T(kp, {
    const int64_t tiles = ceil_div(ncols, 8 * 32 * ResCfg<KP>::C);
    if (x_is_f16)
      return residual_launch<__half, KP>(reinterpret_cast<const __half*>(X), 256, ldx * 256, g, ncols, G32, H32, ldc, 1,
                                         partials, out, residual_splits(g, tiles), st, gate);
    return residual_launch<float, KP>(reinterpret_cast<const float*>(X), ldx, 0, g, ncols, G32, H32, ldc, 1, partials,
                                      out, residual_splits(g, tiles), st, gate);
  });
  return SCRNA_ERR_UNSUPPORTED;
}

extern "C" int scrna_transfer_stop(const double* res, double denom, double rel_err, int max_iter,
                                   scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!res || !ctrl || !(denom > 0.0) || max_iter <= 0) return SCRNA_ERR_BADARG;
  transfer_stop_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(res, denom, rel_err, max_iter, (NmfCtrl*)ctrl);
  return last_launch_status();
}

extern "C" int scrna_argmax_cols(const double* H64, int k, int64_t ncols, int64_t ldc, int32_t* labels,
                                 void* stream) {
  if (!H64 || !labels || k <= 0 || ncols <= 0 || ncols > ldc) return SCRNA_ERR_BADARG;
  argmax_cols_kernel<<<(unsigned)ceil_div(ncols, 256), 256, 0, (cudaStream_t)stream>>>(H64, k, ncols, ldc,
                                                                                     labels);
  return last_launch_status();
}

// mix . (W.H) + (1 - mix) . X with the dense H (get_mixed_data(use_H2=False), nmf_clustering.py:180-184): four cells per
// thread, the k-term dot product of each (gene, cell) in component order, W.H never kept beyond the optional output.
__global__ void __launch_bounds__(256)
mix_dense_kernel(const float* __restrict__ X, int64_t ldx, int g, int64_t ncols, const double* __restrict__ G64, int k,
                 int kp, const double* __restrict__ H64, int64_t ldc, double mix, double* __restrict__ mixed64,
                 int64_t ld64, double* __restrict__ newtrg64) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= ncols) return;
  for (int i = blockIdx.y; i < g; i += gridDim.y) {
    const float4 x = ld_stream(reinterpret_cast<const float4*>(X + (int64_t)i * ldx + j));
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const double* wrow = G64 + (int64_t)i * kp;
    double w[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < k; ++t) {
      const double wt = wrow[t];
      const double* h = H64 + (int64_t)t * ldc + j;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (j + c < ncols) w[c] = fma(wt, h[c], w[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (j + c < ncols) {
        if (mixed64) mixed64[(int64_t)i * ld64 + j + c] = mix * w[c] + (1.0 - mix) * (double)xv[c];
        if (newtrg64) newtrg64[(int64_t)i * ld64 + j + c] = w[c];
      }
    }
  }
}

extern "C" int scrna_mix_dense(const float* X, int64_t ldx, int g, int64_t ncols, const double* G64, int k, int kp,
                               const double* H64, int64_t ldc, double mix, double* mixed64, int64_t ld64,
                               double* newtrg64, void* stream) {
  if (!X || !G64 || !H64 || g <= 0 || ncols <= 0 || k <= 0 || k > kp || (ldx & 3) || ncols > ldx || ncols > ldc ||
      (!mixed64 && !newtrg64) || ld64 < ncols)
    return SCRNA_ERR_BADARG;
  const int64_t quads = ceil_div(ncols, 4);
  int gy = (int)((int64_t)num_sms() * 8 / (ceil_div(quads, 256)));
  if (gy < 1) gy = 1;
  if (gy > g) gy = g;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)ceil_div(quads, 256), (unsigned)gy);
  mix_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, g, ncols, G64, k, kp, H64, ldc, mix, mixed64, ld64,
                                                          newtrg64);
  return last_launch_status();
}

extern "C" int scrna_mix_gather(const float* X, int64_t ldx, int g, int64_t ncols, const double* G64, int kp,
                                const int32_t* labels, double mix, double* mixed64, int64_t ld64,
                                float* mixed32, int64_t ld32, double* newtrg64, void* stream) {
  if (!X || !G64 || !labels || g <= 0 || ncols <= 0 || (ldx & 3) || ncols > ldx ||
      (mixed32 && ((ld32 & 3) || ld32 < ((ncols + 3) & ~3LL))) || ((mixed64 || newtrg64) && ld64 < ncols))
    return SCRNA_ERR_BADARG;
  const int64_t quads = ceil_div(ncols, 4);
  int gy = (int)((int64_t)num_sms() * 8 / (ceil_div(quads, 256)));
  if (gy < 1) gy = 1;
  if (gy > g) gy = g;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)ceil_div(quads, 256), (unsigned)gy);
  mix_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, g, ncols, G64, kp, labels, mix, mixed64,
                                                           ld64, mixed32, ld32, newtrg64);
  return last_launch_status();
}

extern "C" int scrna_reject_stats(const float* X, int64_t ldx, int g, int64_t ncols, const double* G64, int k,
                                  int kp, const double* H64, int64_t ldc, const int32_t* labels, double* out,
                                  int64_t ldo, void* stream) {
  if (!X || !G64 || !H64 || !labels || !out || g <= 0 || ncols <= 0 || k <= 0 || k > kp ||
      kp > SCRNA_MAX_K || ncols > ldx || ncols > ldc || ncols > ldo)
    return SCRNA_ERR_BADARG;
  const size_t smem = (size_t)64 * kp * sizeof(double);
  reject_stats_kernel<<<(unsigned)ceil_div(ncols, 128), 128, smem, (cudaStream_t)stream>>>(
      X, ldx, g, ncols, G64, k, kp, H64, ldc, labels, out, ldo);
  return last_launch_status();
}

// ---- pre-processing filters on the device ------------------------------------------
// One pass over a band of the fp64 host-layout matrix (rows x cols): per-column counts of entries >= thr
// (cell_filter, sc3_clustering_impl.py:32-42), per-row counts of entries >= thr and of entries > 0 (gene_filter,
// :45-58).  NaN counts as 0, as after the reference's in-place `data[np.isnan(data)] = 0`; nan_flag is raised so
// that the caller can apply that in-place zeroing to the host matrix only when there is something to zero.
// One CTA per row: row counts by block reduction, column counts by integer atomics (order-free, exact).
namespace scrna {
__global__ void __launch_bounds__(256)
filter_counts_kernel(const double* __restrict__ X, int64_t ld, int64_t rows, int64_t cols, double thr,
                     int32_t* __restrict__ col_ge, int32_t* __restrict__ row_ge, int32_t* __restrict__ row_gt0,
                     int32_t* __restrict__ nan_flag) {
  __shared__ int sge[8], sgt[8];
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const double* x = X + r * ld;
    int ge = 0, gt = 0;
    bool nan_seen = false;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
      double v = x[c];
      if (v != v) { nan_seen = true; v = 0.0; }
      if (v >= thr) { ++ge; atomicAdd(col_ge + c, 1); }
      gt += (v > 0.0);
    }
    if (nan_seen) atomicOr(nan_flag, 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ge += __shfl_xor_sync(0xffffffffu, ge, o);
      gt += __shfl_xor_sync(0xffffffffu, gt, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sge[threadIdx.x >> 5] = ge; sgt[threadIdx.x >> 5] = gt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int a = 0, b = 0;
      for (int w = 0; w < 8; ++w) { a += sge[w]; b += sgt[w]; }
      row_ge[r] = a;
      row_gt0[r] = b;
    }
  }
}
}  // namespace scrna

extern "C" int scrna_filter_counts_f64(const double* X, int64_t ld, int64_t rows, int64_t cols, double thr,
                                       int32_t* col_ge, int32_t* row_ge, int32_t* row_gt0, int32_t* nan_flag,
                                       void* stream) {
  if (!X || !col_ge || !row_ge || !row_gt0 || !nan_flag || rows <= 0 || cols <= 0 || ld < cols) return SCRNA_ERR_BADARG;
  int64_t grid = rows < (int64_t)num_sms() * 8 ? rows : (int64_t)num_sms() * 8;
  filter_counts_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(X, ld, rows, cols, thr, col_ge, row_ge, row_gt0,
                                                                       nan_flag);
  return last_launch_status();
}
