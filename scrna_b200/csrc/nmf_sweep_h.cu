// 16-bit X sweep for sm_100a: tcgen05.mma.kind::f16 fed DIRECTLY from the TMA-staged tile, no SIMT work
// on the stream at all.
//
//   part[s][t][j] = sum_{i in row span s} F[i][t] * X[i][j]          (F^T . X, same contract as K3)
//
// When every element of X is exactly representable in fp16 (raw counts <= 2048: probed once per matrix by
// scrna_f16_inexact) X and its transposed copy are STORED as fp16, which halves the bytes the HBM-bound
// iteration has to move (algorithmic bytes 2*g*n*s_X with s_X = 2).  The products stay at fp32
// accuracy by splitting the small factor instead of X:
//
//      s_t * F[:, t] = F1 + F2 (+ F3)    (fp16 parts, s_t a power of two per component)
//      F^T.X = X^T.[F1 | F2 (| F3)] summed over the column groups, unscaled by 1/s_t
//
// x (11 significant bits) times an fp16 part (11 bits) is exact in the fp32 accumulator, so the only rounding
// is the accumulation itself -- identical in kind to an fp32 FMA chain.  Two parts carry 22 significant bits of
// the factor (what the 3xTF32 sweep's [hi | lo] factor split carries; the default), three carry 33.  fp16's
// narrow exponent range is what the scale is for: per component and per GROUP of 128 factor rows,
// s = 2^e with max|F| * s in [2^13, 2^14) over the group, so no part can overflow and an entry v is
// represented to max(2^-22 |v| (two parts; 2^-33 with three), 2^-25 / s <= 2^-38 of the group's largest entry):
// full split precision down to 2^-16 of the group maximum, an absolute floor below.  The scale is
// local to the 128 rows an epilogue CTA has just written (no reduction over the whole factor, the shadow is
// emitted by the update kernel itself) and coincides with the accumulator drain period below, where it is
// undone.  tcgen05.mma.kind::f16 does not accept mixed f16 x bf16 operands (measured: illegal instruction,
// tools/microbench/f16_probe.cu), hence fp16 parts and not a bf16 split.
//
// Storage ("panel-major").  X16 is kept as panels of 256 columns: element (i, j) at
//      ((j / 256) * rows_pad + i) * 256 + j % 256,      rows_pad = rows rounded up to 64, padding zero,
// so that the rows a CTA streams are ONE contiguous region of HBM per panel (512 B per row, rows back to back)
// instead of 1 KB pieces a row pitch apart: measured +3 % of TMA throughput over the row-major layout, and cell
// shards on several GPUs stream exactly like the full matrix.
//
// Mapping.  A CTA owns 128*CM consecutive output columns j (CM = 4 up to rank 40, 2 above) and one span of
// reduction rows i.  Per 16 rows and per 128-column block m one MMA
//      D_m[128 lanes][NB] (+)= A_m[128][16] . B[16][NB],          NB = PARTS*KQ rounded up to 16
//   A (shared memory, MN-major, SWIZZLE_128B): exactly the bytes TMA wrote -- the X tile is staged as boxes of
//      64 columns x 64 rows (128-byte rows, the swizzle atom), two boxes side by side form the M = 128 operand
//      (LBO = box size), 8-row groups are the K atoms (SBO = 1024 B);
//   B (shared memory, K-major, no swizzle): the factor rows in the canonical layout [(r/8)][n < NB][r%8];
//   D (tensor memory): lane l of D_m = output column j0 + 128 m + l; column n = part n / KQ of component n % KQ.
// The tensor core truncates its fp32 accumulator at every accumulation (see nmf_sweep_tc.cu; measured here:
// a relative bias of ~1.6e-7 per MMA of the chain), so the accumulators are drained every 128 rows (one scale
// group: the drain multiplies by the group's 2^-e, exact) into fp32 running totals held in REGISTERS (one output
// column per thread and m); two accumulator sets alternate in tensor memory, so the MMA stream never waits for
// a drain.
//
// Pipeline (192 threads, one CTA per SM, a stage = 64 rows: 64 KB of X at CM = 4):
//      warp 4 lane 0 : TMA producer (NBOX tensor boxes + one bulk copy of the factor block per stage);
//      warp 5 lane 0 : MMA issuer (CM * 4 MMAs per stage, tcgen05.commit to "slot free" / "accumulators full");
//      warps 0-3     : drain tensor memory -> register totals -> (end of span) unscale and store the partial.
// Measured on the way (tools/microbench/f16_probe.cu, profiles/r02_f16_sweep_experiments.md): 64-row stages beat
// 32-row stages by 4 %; an L2 prefetcher running 4-14 stages ahead of the ring LOSES 3-30 % (the stream is
// bandwidth- not latency-bound); two co-resident CTAs per SM on 256-column tiles are neutral.
//
// Algorithmic bytes per launch: rows * ncols * 2 (X once) + rows * NB * 2 per column tile (factor, from L2).
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace scrna {

constexpr int H_THREADS = 192;  // 4 drain warps + producer + MMA issuer
constexpr int H_ALIGN = 128;    // row spans, panels and operand shadows are padded to this many rows
constexpr int H_GROUP = 128;    // rows sharing one fp16 scale per component = rows between two accumulator drains
constexpr int H_SR = 64;        // rows per pipeline stage = TMA box height = four K = 16 MMA steps
constexpr int H_PANEL = 256;    // columns per storage panel

template <int KQ, int PARTS>
struct HCfg {
  static_assert(KQ % 8 == 0 && KQ >= 8 && KQ <= 64, "component groups are padded to multiples of 8");
  static_assert(PARTS == 2 || PARTS == 3, "two- or three-part fp16 split of the factor");
  static constexpr int NB = ((PARTS * KQ + 15) / 16) * 16;   // MMA N: [F1 | F2 (| F3) | zero pad]
  // 128-column MMA blocks per CTA: 4 while the accumulators fit tensor memory, the register totals (CM * KQ
  // floats per drain thread) fit the register file and three 64-row stages fit shared memory, else 2
  static constexpr int CM = (4 * NB <= 512 && KQ <= 40 && 3 * (8 * H_SR * 128 + NB * 2 * H_SR) <= 226 * 1024 - 256) ? 4 : 2;
  static constexpr int COLS = 128 * CM;
  static constexpr int KSTEPS = H_SR / 16;
  static constexpr int BOX_BYTES = H_SR * 128;           // 64 fp16 columns x 64 rows
  static constexpr int NBOX = COLS / 64;
  static constexpr int XBYTES = NBOX * BOX_BYTES;
  static constexpr int BSTEP = NB * 32;                  // factor rows of one K-step: 2 chunks x NB x 16 B
  static constexpr int BBYTES = KSTEPS * BSTEP;
  static constexpr int STAGE_BYTES = XBYTES + BBYTES;    // multiple of 1024 (NB % 16 == 0)
  static constexpr int NACC = (2 * CM * NB <= 512) ? 2 : 1;
  static constexpr int NS_FIT = (226 * 1024 - 256) / STAGE_BYTES;
  static constexpr int NS = NS_FIT > 6 ? 6 : NS_FIT;
  static constexpr int SMEM_BYTES = NS * STAGE_BYTES + 1024;
  static_assert(NS >= 3, "need a pipelined shared-memory ring");
  static_assert(STAGE_BYTES % 1024 == 0, "swizzled boxes need 1024-byte aligned stage bases");
};

__device__ __forceinline__ void tc_mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout); layout 0 = no swizzle, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t h_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

// xmap: 2-D tensor map over the panel-major storage viewed as (npanels * rows_pad) rows of 256 columns.
// flags (bring-up probes only): 1 = no MMAs, 2 = no factor copies, 4 = no X copies
template <int KQ, int PARTS>
__global__ void __launch_bounds__(H_THREADS, 1)
wtx_h_kernel(const __grid_constant__ CUtensorMap xmap, int g, int rows_pad, int64_t ncols,
             const __half* __restrict__ Fh, const float* __restrict__ inv_scale, float* __restrict__ part, int64_t ldc,
             int rows_per_split, int flags, const NmfCtrl* __restrict__ ctrl) {
  using Cfg = HCfg<KQ, PARTS>;
  constexpr int NB = Cfg::NB, CM = Cfg::CM, NS = Cfg::NS, NACC = Cfg::NACC;
  if (ctrl_done(ctrl)) return;
  extern __shared__ uint8_t h_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NS + 2 * NACC];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(h_smem_raw) + 1023u) & ~1023u;
  const uint32_t full_x = smem_u32(&bars[0]);
  const uint32_t empty_x = smem_u32(&bars[NS]);
  const uint32_t acc_full = smem_u32(&bars[2 * NS]);
  const uint32_t acc_empty = smem_u32(&bars[2 * NS + NACC]);

  const int64_t col0 = (int64_t)blockIdx.x * Cfg::COLS;
  const int r_begin = blockIdx.y * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > g) r_end = g;
  const int nstages = r_end > r_begin ? (r_end - r_begin + H_SR - 1) / H_SR : 0;
  constexpr int drain = H_GROUP / H_SR;   // stages per drain period
  const int nperiods = (nstages + drain - 1) / drain;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(full_x + 8 * i, 1);
      mbar_init(empty_x + 8 * i, 1);
    }
    for (int i = 0; i < NACC; ++i) {
      mbar_init(acc_full + 8 * i, 1);
      mbar_init(acc_empty + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 4) {
    // ===== TMA producer.  Spans are multiples of the stage height and the storage is zero-padded to whole
    // stages and panels, so a stage never straddles two spans; a panel past the matrix reads as zeros.
    if (lane == 0) {
      const int panel0 = (int)(col0 / H_PANEL);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nstages; ++i) {
        const int r0 = r_begin + i * H_SR;
        const uint32_t dst = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t bar = full_x + 8 * s;
        mbar_wait(empty_x + 8 * s, ph ^ 1u);
        mbar_expect_tx(bar, ((flags & 4) ? 0u : (uint32_t)Cfg::XBYTES) + ((flags & 2) ? 0u : (uint32_t)Cfg::BBYTES));
        if (!(flags & 4)) {
#pragma unroll
          for (int b = 0; b < Cfg::NBOX; ++b)
            tma_load_2d(dst + b * Cfg::BOX_BYTES, &xmap, (b & 3) * 64, (panel0 + (b >> 2)) * rows_pad + r0, bar);
        }
        if (!(flags & 2)) bulk_g2s(dst + Cfg::XBYTES, Fh + (int64_t)r0 * NB, Cfg::BBYTES, bar);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // D = f32, A = B = f16, A MN-major (bit 15), B K-major, M = 128, N = NB
      constexpr uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
      int period = 0, pos = 0, s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nstages; ++i) {
        const bool first = pos == 0;
        const bool last = pos + 1 == drain || i + 1 == nstages;
        const int buf = period % NACC;
        // the drain of this accumulator set's previous period must have read it
        if (first && period >= NACC) mbar_wait(acc_empty + 8 * buf, (uint32_t)((period / NACC - 1) & 1));
        mbar_wait(full_x + 8 * s, ph);
        tc_fence_after();
        const uint32_t xsm = smem_base + s * Cfg::STAGE_BYTES, bsm = xsm + Cfg::XBYTES;
#pragma unroll
        for (int u = 0; u < ((flags & 1) ? 0 : Cfg::KSTEPS); ++u) {
          const uint64_t bdesc = h_smem_desc(bsm + u * Cfg::BSTEP, NB * 16u, 128u, 0u);
          const uint32_t acc = (first && u == 0) ? 0u : 1u;
#pragma unroll
          for (int m = 0; m < CM; ++m) {
            const uint64_t adesc = h_smem_desc(xsm + m * 2 * Cfg::BOX_BYTES + u * 2048, Cfg::BOX_BYTES, 1024u, 2u);
            tc_mma_f16_ss(tmem_base + (buf * CM + m) * NB, adesc, bdesc, idesc, acc);
          }
        }
        tc_commit(empty_x + 8 * s);
        if (last) {
          tc_commit(acc_full + 8 * buf);
          ++period;
          pos = 0;
        } else {
          ++pos;
        }
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===== drain warps (0-3): warp w reads tensor-memory lanes 32w .. 32w+31; thread tid owns output columns
    // col0 + 128 m + tid and keeps their running totals in registers =====
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    float tot[CM][KQ];
#pragma unroll
    for (int m = 0; m < CM; ++m)
#pragma unroll
      for (int t = 0; t < KQ; ++t) tot[m][t] = 0.f;
    for (int p = 0; p < nperiods; ++p) {
      const int buf = p % NACC;
      // the period's rows are one scale group of the factor shadow: 1 / scale per component
      const float4* inv = reinterpret_cast<const float4*>(inv_scale + ((int64_t)(r_begin / H_GROUP) + p) * KQ);
      mbar_wait(acc_full + 8 * buf, (uint32_t)((p / NACC) & 1));
      tc_fence_after();
#pragma unroll
      for (int m = 0; m < CM; ++m) {
        const uint32_t acc = tmem_base + lane_field + (buf * CM + m) * NB;
#pragma unroll
        for (int t0 = 0; t0 < KQ; t0 += 8) {
          float a[8], b[8], c[8];
          tc_ld8(acc + t0, a);
          tc_ld8(acc + KQ + t0, b);
          if (PARTS == 3) tc_ld8(acc + 2 * KQ + t0, c);
          tc_wait_ld();
          if (m + 1 == CM && t0 + 8 == KQ) {  // every column of this accumulator set has been read
            tc_fence_before();
            mbar_arrive(acc_empty + 8 * buf);
          }
          const float4 s0 = inv_scale ? __ldg(inv + t0 / 4) : make_float4(1.f, 1.f, 1.f, 1.f);
          const float4 s1 = inv_scale ? __ldg(inv + t0 / 4 + 1) : make_float4(1.f, 1.f, 1.f, 1.f);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int t = 0; t < 8; ++t)  // smallest parts first; unscale (a power of two: exact) and accumulate
            tot[m][t0 + t] = fmaf(PARTS == 3 ? (c[t] + b[t]) + a[t] : b[t] + a[t], sc[t], tot[m][t0 + t]);
        }
      }
    }
    float* const out = part + ((int64_t)blockIdx.y * KQ) * ldc + col0 + tid;
#pragma unroll
    for (int m = 0; m < CM; ++m) {
      if (col0 + m * 128 + tid < ncols) {
#pragma unroll
        for (int t = 0; t < KQ; ++t)
          out[(int64_t)t * ldc + m * 128] = tot[m][t];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- operand shadow ---------------------------------------------------------------------------------------
// CTAs [0, groups): one 128-row scale group each.  Per component the group's largest |entry| fixes the scale
// 2^e with max * 2^e in [2^13, 2^14) -- local to the group, so no reduction over the whole factor is needed and
// small-magnitude stretches of a factor keep their relative accuracy -- then the rows are split into fp16 parts,
// K-major canonical layout out[((r/8) * NB + n) * 8 + r%8], n = part * KQ + t; inv_scale[group * KQ + t] = 2^-e.
// A thread owns (8-row group, component) items: 64 contiguous bytes of the master in, 16-byte vectors out,
// consecutive threads writing consecutive vectors.
// CTAs [groups, ...): optionally the k x k Gram finish (one warp per entry; gram_final_kernel folded into the same
// launch so that an iteration does not grow by a kernel).
constexpr int HS_THREADS = 256;
__device__ __forceinline__ float h_group_scale(double mx, float* inv) {
  // mx * scale in [2^13, 2^14): every entry of the group is below 2^14, far inside fp16's range
  double sc = 1.0;
  if (mx > 0.0 && mx < 1e300) {
    int e;
    frexp(mx, &e);  // mx = m * 2^e, m in [0.5, 1)
    if (e < -100) e = -100;  // keep 2^e and 2^-e inside fp32 (entries below 2^-100 lose relative accuracy only)
    sc = ldexp(1.0, 14 - e);
  }
  *inv = (float)(1.0 / sc);
  return (float)sc;
}

__global__ void __launch_bounds__(HS_THREADS)
h_shadow_kernel(const double* __restrict__ F64, int64_t ld, int64_t rows, int k, int kq, int nb, int parts,
                const double* __restrict__ gpart, int nblk, int kp, double* __restrict__ gram_out,
                __half* __restrict__ out, float* __restrict__ inv_scale, int groups, const NmfCtrl* __restrict__ ctrl) {
  if (ctrl_done(ctrl)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x >= groups) {
    const int o = ((int)blockIdx.x - groups) * (HS_THREADS >> 5) + warp;
    if (o >= kp * kp || gram_out == nullptr) return;
    const int a = o / kp, b = o - a * kp;
    double s = 0.0;
    if (a < k && b < k)
      for (int p = lane; p < nblk; p += 32) s += gpart[(int64_t)p * kp * kp + o];
    s = warp_sum(s);
    if (lane == 0) gram_out[o] = s;
    return;
  }
  __shared__ double smax[16][SCRNA_MAX_K];
  __shared__ float sscale[SCRNA_MAX_K];
  const int64_t r0 = (int64_t)blockIdx.x * H_GROUP;
  // pass 1: per (8-row group, component) maxima
  for (int item = tid; item < 16 * kq; item += HS_THREADS) {
    const int rg = item / kq, t = item - rg * kq;
    double mx = 0.0;
    if (t < k)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t r = r0 + rg * 8 + e;
        if (r < rows) mx = fmax(mx, fabs(F64[(int64_t)t * ld + r]));
      }
    smax[rg][t] = mx;
  }
  __syncthreads();
  for (int t = tid; t < kq; t += HS_THREADS) {
    double mx = 0.0;
#pragma unroll
    for (int rg = 0; rg < 16; ++rg) mx = fmax(mx, smax[rg][t]);
    float inv;
    sscale[t] = h_group_scale(mx, &inv);
    inv_scale[(int64_t)blockIdx.x * kq + t] = inv;
  }
  __syncthreads();
  // pass 2: split and store
  for (int item = tid; item < 16 * kq; item += HS_THREADS) {
    const int rg = item / kq, t = item - rg * kq;
    union Pack { uint4 v; __half h[8]; } p1, p2, p3;
    const double sc = t < k ? (double)sscale[t] : 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t r = r0 + rg * 8 + e;
      double v = (r < rows && t < k) ? F64[(int64_t)t * ld + r] * sc : 0.0;
      p1.h[e] = __double2half(v);
      v -= (double)__half2float(p1.h[e]);
      p2.h[e] = __double2half(v);
      v -= (double)__half2float(p2.h[e]);
      p3.h[e] = __double2half(v);
    }
    uint4* base = reinterpret_cast<uint4*>(out) + ((r0 >> 3) + rg) * nb;
    base[t] = p1.v;
    base[kq + t] = p2.v;
    if (parts == 3) base[2 * kq + t] = p3.v;
    if (parts * kq + t < nb) base[parts * kq + t] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// inexact[0] |= 1 if any element of X is not exactly representable in fp16 (value or range)
__global__ void f16_exact_kernel(const float4* __restrict__ X, int64_t n4, int* __restrict__ inexact) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = X[i];
    bad |= __half2float(__float2half_rn(v.x)) != v.x;
    bad |= __half2float(__float2half_rn(v.y)) != v.y;
    bad |= __half2float(__float2half_rn(v.z)) != v.z;
    bad |= __half2float(__float2half_rn(v.w)) != v.w;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(inexact, 1);
}

// panel-major fp16 storage of a row-major fp32 matrix: dst[((c / 256) * rows_pad + r) * 256 + c % 256];
// padding (rows >= rows, columns >= cols) is zeroed by the caller
__global__ void narrow_f16_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int64_t cols,
                                  __half* __restrict__ dst, int64_t rows_pad) {
  const int64_t r = blockIdx.y;
  const float* s = src + r * ld_src;
  for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; c < cols;
       c += (int64_t)gridDim.x * blockDim.x * 2) {
    const float a = s[c], b = c + 1 < cols ? s[c + 1] : 0.f;
    *reinterpret_cast<__half2*>(dst + ((c >> 8) * rows_pad + r) * H_PANEL + (c & 255)) = __floats2half2_rn(a, b);
  }
}

// the transposed copy in the same storage: element (c, r) of X^T, i.e. dst[((r / 256) * cols_pad + c) * 256 + r % 256]
__global__ void __launch_bounds__(256)
transpose_f16_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int64_t cols,
                     __half* __restrict__ dst, int64_t cols_pad) {
  __shared__ float tile[64][65];
  const int64_t c0 = (int64_t)blockIdx.x * 64, r0 = (int64_t)blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
  const int px = threadIdx.x & 31, py = threadIdx.x >> 5;  // 32 row pairs x 8
#pragma unroll
  for (int i = 0; i < 64; i += 8) {
    const int64_t c = c0 + py + i, r = r0 + 2 * px;  // r even: the pair stays inside one panel
    if (c < cols && r < rows)
      *reinterpret_cast<__half2*>(dst + ((r >> 8) * cols_pad + c) * H_PANEL + (r & 255)) =
          __floats2half2_rn(tile[2 * px][py + i], tile[2 * px + 1][py + i]);
  }
}

template <int KQ, int PARTS>
static int wtx_h_launch(const __half* X, int64_t rows_pad, int g, int64_t ncols, const __half* Fh, const float* inv_scale,
                        float* part, int64_t ldc, int nsplit, int rps, int flags, const NmfCtrl* ctrl,
                        cudaStream_t st) {
  using Cfg = HCfg<KQ, PARTS>;
  CUtensorMap xmap;
  const int64_t npanels = ceil_div(ncols, H_PANEL);
  if (int rc = make_tensor_map_2d(&xmap, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, H_PANEL, npanels * rows_pad, H_PANEL, 64,
                                  H_SR, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(wtx_h_kernel<KQ, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return SCRNA_ERR_CUDA;
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(ncols, Cfg::COLS), (unsigned)nsplit);
  wtx_h_kernel<KQ, PARTS><<<grid, H_THREADS, Cfg::SMEM_BYTES, st>>>(xmap, g, (int)rows_pad, ncols, Fh, inv_scale, part, ldc,
                                                                   rps, flags, ctrl);
  return last_launch_status();
}

static int h_rows_per_split(int g, int nsplit) {
  const int rps = (int)ceil_div(g, nsplit);
  return (int)(ceil_div(rps, H_ALIGN) * H_ALIGN);
}

static int h_nb(int kq, int parts) { return ((parts * kq + 15) / 16) * 16; }
static int h_cols_per_cta(int kq, int parts) {
  const int nb = h_nb(kq, parts);
  return (4 * nb <= 512 && kq <= 40 && 3 * (8 * H_SR * 128 + nb * 2 * H_SR) <= 226 * 1024 - 256) ? 512 : 256;
}

}  // namespace scrna

using namespace scrna;

#define SCRNA_DISPATCH_H(kq, ...)                           \
  switch (kq) {                                             \
    case 8: { constexpr int KQ = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int KQ = 16; __VA_ARGS__; } break; \
    case 24: { constexpr int KQ = 24; __VA_ARGS__; } break; \
    case 32: { constexpr int KQ = 32; __VA_ARGS__; } break; \
    case 40: { constexpr int KQ = 40; __VA_ARGS__; } break; \
    case 48: { constexpr int KQ = 48; __VA_ARGS__; } break; \
    case 56: { constexpr int KQ = 56; __VA_ARGS__; } break; \
    case 64: { constexpr int KQ = 64; __VA_ARGS__; } break; \
    default: return SCRNA_ERR_UNSUPPORTED;                  \
  }

extern "C" int scrna_nmf_h_supported(int kq) { return (kq >= 8 && kq <= 64 && (kq & 7) == 0) ? 1 : 0; }

extern "C" int scrna_nmf_h_nb(int kq, int parts) {
  return (scrna_nmf_h_supported(kq) && (parts == 2 || parts == 3)) ? h_nb(kq, parts) : SCRNA_ERR_BADARG;
}

extern "C" int64_t scrna_nmf_h_shadow_elems(int64_t rows, int kq, int parts) {
  if (!scrna_nmf_h_supported(kq) || rows <= 0 || (parts != 2 && parts != 3)) return SCRNA_ERR_BADARG;
  return ceil_div(rows, H_ALIGN) * H_ALIGN * (int64_t)h_nb(kq, parts);
}

extern "C" int64_t scrna_nmf_h_groups(int64_t rows) { return rows > 0 ? ceil_div(rows, H_GROUP) : SCRNA_ERR_BADARG; }

extern "C" int64_t scrna_f16_panel_rows(int64_t rows) { return rows > 0 ? ceil_div(rows, H_ALIGN) * H_ALIGN : SCRNA_ERR_BADARG; }

extern "C" int64_t scrna_f16_panel_elems(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return SCRNA_ERR_BADARG;
  return ceil_div(cols, H_PANEL) * scrna_f16_panel_rows(rows) * H_PANEL;
}

extern "C" int scrna_nmf_wtx_h_splits(int g, int64_t ncols, int kq, int parts) {
  if (g <= 0 || ncols <= 0 || !scrna_nmf_h_supported(kq) || (parts != 2 && parts != 3)) return SCRNA_ERR_BADARG;
  const int64_t tiles = ceil_div(ncols, h_cols_per_cta(kq, parts));
  const int64_t slots = (int64_t)num_sms();
  int max_split = g / 512;
  if (max_split > 64) max_split = 64;
  if (max_split < 1) max_split = 1;
  // Time model of a sweep: CTAs run one per SM in waves, a CTA costs its row span plus a fixed part (pipeline fill,
  // tensor-memory allocation, last drain, partial write) worth ~700 rows of streaming (measured on the 12 500-cell
  // shard of the 8-GPU run: 5 splits x 25 tiles beat 11 x 25, profiles/r02_split_counts_shard_12500.txt).
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= max_split; ++s) {
    const int rps = h_rows_per_split(g, s);
    if ((int)ceil_div(g, rps) != s) continue;  // only split counts whose spans are all non-empty
    const int64_t ctas = tiles * s;
    const int64_t waves = ceil_div(ctas, slots);
    const double cost = (double)waves * ((double)rps + 700.0);
    if (cost < best_cost * (1.0 - 1e-9)) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

extern "C" int scrna_nmf_wtx_partial_h(const void* X16, int64_t rows_pad, int g, int64_t ncols, const void* Fh,
                                       const float* inv_scale, int kq, int parts, float* part, int64_t ldc, int nsplit,
                                       int flags, const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!X16 || !Fh || !part || g <= 0 || ncols <= 0 || nsplit <= 0 || rows_pad < g || (rows_pad % H_ALIGN) ||
      ncols > ldc || (reinterpret_cast<uintptr_t>(X16) & 127) || (parts != 2 && parts != 3) ||
      ceil_div(ncols, H_PANEL) * rows_pad > 0x7fffffffLL)
    return SCRNA_ERR_BADARG;
  const int rps = h_rows_per_split(g, nsplit);
  if ((int64_t)rps * (nsplit - 1) >= g && nsplit > 1) return SCRNA_ERR_BADARG;  // an empty span
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const NmfCtrl* c = reinterpret_cast<const NmfCtrl*>(ctrl);
  const __half* Xh = reinterpret_cast<const __half*>(X16);
  const __half* F = reinterpret_cast<const __half*>(Fh);
  if (parts == 2) {
    SCRNA_DISPATCH_H(kq, return (wtx_h_launch<KQ, 2>(Xh, rows_pad, g, ncols, F, inv_scale, part, ldc, nsplit, rps,
                                                     flags & 0xff, c, st)));
  } else {
    SCRNA_DISPATCH_H(kq, return (wtx_h_launch<KQ, 3>(Xh, rows_pad, g, ncols, F, inv_scale, part, ldc, nsplit, rps,
                                                     flags & 0xff, c, st)));
  }
  return SCRNA_ERR_UNSUPPORTED;
}

extern "C" int scrna_nmf_h_shadow(const double* F64, int64_t ld, int64_t rows, int k, int kq, int parts,
                                  const double* gram_parts, int nblk, int kp, double* gram_out, void* Fh,
                                  float* inv_scale, const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!F64 || !Fh || !inv_scale || rows <= 0 || k <= 0 || k > kq || !scrna_nmf_h_supported(kq) ||
      (gram_out && (!gram_parts || nblk <= 0 || kp < k)) || (parts != 2 && parts != 3))
    return SCRNA_ERR_BADARG;
  const int groups = (int)ceil_div(rows, H_GROUP);
  const int gram_ctas = gram_out ? (int)ceil_div((int64_t)kp * kp, HS_THREADS >> 5) : 0;
  h_shadow_kernel<<<groups + gram_ctas, HS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      F64, ld, rows, k, kq, h_nb(kq, parts), parts, gram_parts, nblk, kp, gram_out, reinterpret_cast<__half*>(Fh),
      inv_scale, groups, reinterpret_cast<const NmfCtrl*>(ctrl));
  return last_launch_status();
}

extern "C" int scrna_f16_inexact(const float* X, int64_t count, int32_t* inexact, void* stream) {
  if (!X || !inexact || count <= 0 || (count & 3)) return SCRNA_ERR_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(inexact, 0, sizeof(int32_t), st) != cudaSuccess) return SCRNA_ERR_CUDA;
  f16_exact_kernel<<<num_sms() * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(X), count / 4, inexact);
  return last_launch_status();
}

extern "C" int scrna_narrow_f32_f16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* dst,
                                    void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || cols > ld_src) return SCRNA_ERR_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t rows_pad = scrna_f16_panel_rows(rows);
  if (cudaMemsetAsync(dst, 0, (size_t)scrna_f16_panel_elems(rows, cols) * 2, st) != cudaSuccess) return SCRNA_ERR_CUDA;
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
    const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
    int bx = (int)ceil_div(cols, 512);
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)nr);
    narrow_f16_kernel<<<grid, 256, 0, st>>>(src + r0 * ld_src, ld_src, nr, cols,
                                            reinterpret_cast<__half*>(dst) + r0 * H_PANEL, rows_pad);
  }
  return last_launch_status();
}

extern "C" int scrna_transpose_f32_f16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* dst,
                                       void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || cols > ld_src) return SCRNA_ERR_BADARG;
  const int64_t by = ceil_div(rows, 64);
  if (by > 65535) return SCRNA_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // the transposed matrix has `cols` rows and `rows` columns
  if (cudaMemsetAsync(dst, 0, (size_t)scrna_f16_panel_elems(cols, rows) * 2, st) != cudaSuccess) return SCRNA_ERR_CUDA;
  dim3 grid((unsigned)ceil_div(cols, 64), (unsigned)by);
  transpose_f16_kernel<<<grid, 256, 0, st>>>(src, ld_src, rows, cols, reinterpret_cast<__half*>(dst),
                                             scrna_f16_panel_rows(cols));
  return last_launch_status();
}
