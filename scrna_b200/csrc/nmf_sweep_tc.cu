// Tensor-core X sweep for sm_100a (tcgen05 + TMEM + tensor-map TMA), ranks kp = 16 .. 64.
//
//   part[s][t][j] = sum_{i in row span s} F[i][t] * X[i][j]          (F^T . X, same contract as K3)
//
// For kp <= 8 the sweep is HBM-bound on the FP32 pipe (nmf_sweep.cu, 99 % of the measured copy
// bandwidth).  From kp = 16 on it needs 2*kp flop per 4-byte element -- more than the FMA pipe can
// issue at HBM speed (measured 3.3 TB/s at kp=16, 1.9 TB/s at kp=32) -- so the contraction moves
// to the 5th-generation tensor cores with a 3xTF32 split that keeps fp32-level accuracy:
//
//      X = Xh + Xl,  F = Fh + Fl   (each part exactly representable in TF32)
//      F^T.X ~= Xh.[Fh | Fl] + Xl.Fh                  (the dropped Fl^T.Xl term is 2^-22 relative)
//
// Mapping.  A CTA owns 128*CPL consecutive output columns j (CPL = 4 for kp <= 32, 2 above) and one
// span of reduction rows i.  Per 8 rows and per column block c < CPL two MMAs are issued,
//      D_c[128 lanes][2kp] (+)= Ah_c[128][8] . [Fh | Fl][8][2kp]     (N = 2kp: main and correction halves)
//      D_c[128 lanes][ kp]  += Al_c[128][8] . Fh[8][kp]
// with  A (tensor memory): lane l, column q  <-  X[i0+q][j0 + CPL*l + c]  (hi or lo part),
//       B (shared memory): factor rows i0..i0+7 in the K-major canonical layout (no swizzle),
//       D (tensor memory): lane l of D_c = output column j0 + CPL*l + c.
// Every thread therefore keeps the coalesced vector access pattern of the SIMT sweep (CPL consecutive
// columns per lane) and no transposition is needed anywhere.
//
// Accuracy.  The tensor core truncates the fp32 accumulator at every accumulation (measured: 3.8e-8
// relative per tcgen05.mma, a bias that grows linearly with the chain length), so the accumulators are
// DRAINED every `drain` stages: tcgen05.ld -> fp32 round-to-nearest add into running totals in shared
// memory -> the next stage restarts the tensor-memory accumulators from zero.
//
// Pipeline (warp-specialised, 192 threads, one CTA per SM, a stage = 32 KB of X: 16 or 32 rows):
//      warp 4 lane 0 : TMA producer -- the stage's X tile as 256-column tensor-map boxes + its factor
//                      block (bulk copy) into a shared-memory ring, completion on an mbarrier (expect_tx);
//      warps 0-3     : read the staged rows (conflict-free vector LDS), split hi/lo, tcgen05.st the operand
//                      columns into a TMEM ring; drain the accumulators at period ends;
//      warp 5 lane 0 : issues the tcgen05.mma's of the stage and tcgen05.commit's it to the
//                      "smem slot free" / "TMEM slot free" / "accumulators complete" barriers.
//
// Algorithmic bytes per launch: rows * ldx * 4 (X once) + rows * kp * 8 * column tiles (factor, from L2).
// Replaces the same BLAS products as nmf_sweep.cu (sklearn _nmf.py:380-381, 537-538, 644).
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace scrna {

constexpr int TC_STEP = 8;        // reduction rows per tcgen05.mma.kind::tf32 (K = 8)
constexpr int TC_ROW_ALIGN = 32;  // row spans and operand shadows are padded to this many rows
constexpr int TC_THREADS = 192;   // 4 converter/drain warps + producer warp + MMA warp
constexpr uint32_t TF32_MASK = 0xffffe000u;

template <int KP>
struct TcCfg {
  static_assert(KP % 16 == 0 && KP >= 16 && KP <= 64, "tcgen05 tf32 with M=128 needs N % 16 == 0");
  static constexpr int CPL = KP <= 32 ? 4 : 2;            // output columns per TMEM lane
  static constexpr int COLS = 128 * CPL;                  // output columns per CTA
  static constexpr int KSTEPS = 8 / CPL;                  // MMA K-steps per pipeline stage
  static constexpr int STAGE_ROWS = KSTEPS * TC_STEP;     // 16 rows x 2 KB or 32 rows x 1 KB: 32 KB of X
  static constexpr int XBYTES = STAGE_ROWS * COLS * 4;
  static constexpr int TMEM_COLS = 512;                   // one CTA per SM owns the whole tensor memory
  static constexpr int ACC_COLS = CPL * 2 * KP;           // per column block: main | correction
  static constexpr int SLOT_COLS = KSTEPS * CPL * 16;     // operand columns of one stage (hi + lo per K-step)
  static constexpr int NT = (TMEM_COLS - ACC_COLS) / SLOT_COLS;  // TMEM operand ring depth
  static constexpr int BSTEP = KP * 64;                   // [Fh | Fl] rows of one K-step: 2 chunks x 2kp x 16 B
  static constexpr int BBYTES = KSTEPS * BSTEP;
  static constexpr int STAGE_BYTES = XBYTES + BBYTES;
  static constexpr int TOTAL_BYTES = KP * COLS * 4;       // running totals of the drained accumulators
  static constexpr int NS_FIT = (220 * 1024 - TOTAL_BYTES) / STAGE_BYTES;
  static constexpr int NS = NS_FIT > 8 ? 8 : NS_FIT;      // shared-memory ring depth
  static constexpr int SMEM_BYTES = NS * STAGE_BYTES + TOTAL_BYTES + 1024;
  static_assert(STAGE_ROWS <= 32 && TC_ROW_ALIGN % STAGE_ROWS == 0, "producer: one lane per staged row");
  static_assert(NT >= 2 && NS >= 3, "need pipelined operand rings");
};

// ---------------------------------------------------------------------------------------------
// flags: 32 = X is exactly representable in TF32 (Xl == 0: the second product is skipped);
// bring-up probes only: 2 = skip the Xl pass, 4 = no MMAs, 8 = no X copies, 16 = no operand staging
// ---------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(TC_THREADS, 1)
wtx_tc_kernel(const __grid_constant__ CUtensorMap xmap, int g, int64_t ncols, const float* __restrict__ Fs,
              float* __restrict__ part, int64_t ldc, int rows_per_split, int drain, int flags,
              const NmfCtrl* __restrict__ ctrl) {
  using Cfg = TcCfg<KP>;
  constexpr int CPL = Cfg::CPL;
  if (ctrl_done(ctrl)) return;
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * Cfg::NS + 2 * Cfg::NT + 2];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  const uint32_t full_x = smem_u32(&bars[0]);
  const uint32_t empty_x = smem_u32(&bars[Cfg::NS]);
  const uint32_t tmem_full = smem_u32(&bars[2 * Cfg::NS]);
  const uint32_t tmem_empty = smem_u32(&bars[2 * Cfg::NS + Cfg::NT]);
  const uint32_t accum_full = smem_u32(&bars[2 * Cfg::NS + 2 * Cfg::NT]);
  const uint32_t accum_empty = smem_u32(&bars[2 * Cfg::NS + 2 * Cfg::NT + 1]);

  const int64_t col0 = (int64_t)blockIdx.x * Cfg::COLS;
  const int r_begin = blockIdx.y * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > g) r_end = g;
  const int nstages = r_end > r_begin ? (r_end - r_begin + Cfg::STAGE_ROWS - 1) / Cfg::STAGE_ROWS : 0;

  if (tid == 0) {
    for (int i = 0; i < Cfg::NS; ++i) {
      mbar_init(full_x + 8 * i, 1);
      mbar_init(empty_x + 8 * i, 1);
    }
    for (int i = 0; i < Cfg::NT; ++i) {
      mbar_init(tmem_full + 8 * i, 128);
      mbar_init(tmem_empty + 8 * i, 1);
    }
    mbar_init(accum_full, 1);
    mbar_init(accum_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tmem_a0 = tmem_base + Cfg::ACC_COLS;  // operand ring starts after the accumulators

  if (warp == 4) {
    // ===== TMA producer: one tensor copy per 256-column box of the stage + the stage's factor block.
    // Rows past g and columns past ncols are zero-filled by the tensor map; spans are multiples of the
    // stage height, so a stage never straddles two spans.
    if (lane == 0) {
      int s = 0;
      uint32_t ph_s = 0;
      for (int i = 0; i < nstages; ++i) {
        const int r0 = r_begin + i * Cfg::STAGE_ROWS;
        const uint32_t dst = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t bar = full_x + 8 * s;
        mbar_wait(empty_x + 8 * s, ph_s ^ 1u);
        mbar_expect_tx(bar, ((flags & 8) ? 0u : (uint32_t)Cfg::XBYTES) + Cfg::BBYTES);
        if (!(flags & 8)) {
#pragma unroll
          for (int h = 0; h < Cfg::COLS / 256; ++h)
            tma_load_2d(dst + h * (Cfg::STAGE_ROWS * 1024), &xmap, (int)col0 + h * 256, r0, bar);
        }
        bulk_g2s(dst + Cfg::XBYTES, Fs + (int64_t)r0 * (2 * KP), Cfg::BBYTES, bar);
        if (++s == Cfg::NS) { s = 0; ph_s ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptors: D=f32, A=B=tf32, A from TMEM (K-major), B K-major, M=128, N = 2kp / kp
      constexpr uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
      constexpr uint32_t idesc_2n = idesc_base | ((uint32_t)((2 * KP) >> 3) << 17);
      constexpr uint32_t idesc_1n = idesc_base | ((uint32_t)(KP >> 3) << 17);
      constexpr uint32_t lbo = 2u * KP * 16u, sbo = 128u;  // K chunks 2kp*16 B apart, 8-row groups 128 B apart
      int period = 0;  // completed drain periods
      int pos = 0;     // stage index within the current period
      int s = 0, b = 0;
      uint32_t ph_b = 0;  // phase parity of the TMEM ring
      for (int i = 0; i < nstages; ++i) {
        const bool first = pos == 0;
        const bool last = pos + 1 == drain || i + 1 == nstages;
        if (first && period > 0) mbar_wait(accum_empty, (period - 1) & 1);  // previous drain has read D
        // the converters' arrival also orders the stage's bulk copies (they waited on full_x) before us
        mbar_wait(tmem_full + 8 * b, ph_b);
        tc_fence_after();
        const uint32_t bsm = smem_base + s * Cfg::STAGE_BYTES + Cfg::XBYTES;
        if (!(flags & 4)) {
#pragma unroll
          for (int u = 0; u < Cfg::KSTEPS; ++u) {
            const uint64_t bdesc = tc_smem_desc(bsm + u * Cfg::BSTEP, lbo, sbo);
            const uint32_t ahi = tmem_a0 + b * Cfg::SLOT_COLS + u * (CPL * 16), alo = ahi + CPL * 8;
            const uint32_t acc = (first && u == 0) ? 0u : 1u;
#pragma unroll
            for (int c = 0; c < CPL; ++c) tc_mma_tf32_ts(tmem_base + c * 2 * KP, ahi + c * 8, bdesc, idesc_2n, acc);
            if (!(flags & (2 | 32))) {  // 32: X is exactly representable in TF32, Xl == 0
#pragma unroll
              for (int c = 0; c < CPL; ++c) tc_mma_tf32_ts(tmem_base + c * 2 * KP, alo + c * 8, bdesc, idesc_1n, 1u);
            }
          }
        }
        tc_commit(empty_x + 8 * s);
        tc_commit(tmem_empty + 8 * b);
        if (last) {
          tc_commit(accum_full);
          ++period;
          pos = 0;
        } else {
          ++pos;
        }
        if (++s == Cfg::NS) s = 0;
        if (++b == Cfg::NT) { b = 0; ph_b ^= 1u; }
      }
    }
  } else {
    // ===== converters (warps 0-3): smem rows -> hi/lo -> TMEM operand ring; drains =====
    using VecT = typename std::conditional<CPL == 4, float4, float2>::type;
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    uint8_t* const smem_al = tc_smem_raw + (smem_base - smem_u32(tc_smem_raw));
    VecT* const totals = reinterpret_cast<VecT*>(smem_al + Cfg::NS * Cfg::STAGE_BYTES);
    const int64_t j = col0 + CPL * (int64_t)tid;
    float* const out = part + ((int64_t)blockIdx.y * KP) * ldc + j;
    int period = 0;

    // accumulators (main + correction) -> running totals, or the split partial at the end of the span
    auto drain_accumulators = [&](bool final_drain) {
      mbar_wait(accum_full, period & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t0 = 0; t0 < KP; t0 += 8) {
        float m[CPL][8], e[CPL][8];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          tc_ld8(tmem_base + lane_field + c * 2 * KP + t0, m[c]);
          tc_ld8(tmem_base + lane_field + c * 2 * KP + KP + t0, e[c]);
        }
        tc_wait_ld();
        if (t0 + 8 == KP) {  // every accumulator column has been read: the MMA warp may overwrite D
          tc_fence_before();
          mbar_arrive(accum_empty);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          float v[CPL];
#pragma unroll
          for (int c = 0; c < CPL; ++c) v[c] = m[c][t] + e[c][t];
          VecT* slot = totals + (t0 + t) * 128 + tid;
          if (period > 0) {
            const VecT o = *slot;
            const float* of = reinterpret_cast<const float*>(&o);
#pragma unroll
            for (int c = 0; c < CPL; ++c) v[c] += of[c];
          }
          VecT w;
          float* wf = reinterpret_cast<float*>(&w);
#pragma unroll
          for (int c = 0; c < CPL; ++c) wf[c] = v[c];
          if (!final_drain) *slot = w;
          else if (j < ncols) *reinterpret_cast<VecT*>(out + (int64_t)(t0 + t) * ldc) = w;
        }
      }
      ++period;
    };

    bool pending = false;
    const bool exact = (flags & 32) != 0;
    int s = 0, b = 0, pos = 0;
    uint32_t ph_s = 0, ph_b = 0;
    for (int i = 0; i < nstages; ++i) {
      mbar_wait(full_x + 8 * s, ph_s);
      // staged as 256-column boxes: [box][row][256]; this thread's CPL columns live in box tid*CPL/256
      const VecT* xs = reinterpret_cast<const VecT*>(smem_al + s * Cfg::STAGE_BYTES) +
                       ((tid * CPL) / 256) * (Cfg::STAGE_ROWS * 256 / CPL) + (tid * CPL % 256) / CPL;
      bool slot_ready = false;
#pragma unroll
      for (int u = 0; u < Cfg::KSTEPS; ++u) {
        uint32_t hi[CPL * 8], lo[CPL * 8];
        if (!(flags & 16)) {
#pragma unroll
          for (int q = 0; q < TC_STEP; ++q) {
            VecT v = xs[(u * TC_STEP + q) * (256 / CPL)];
            const float* e = reinterpret_cast<const float*>(&v);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
              const float x = e[c];
              const uint32_t h = __float_as_uint(x) & TF32_MASK;
              hi[c * 8 + q] = h;
              lo[c * 8 + q] = __float_as_uint(x - __uint_as_float(h)) & TF32_MASK;
            }
          }
        }
        if (!slot_ready) {
          mbar_wait(tmem_empty + 8 * b, ph_b ^ 1u);
          tc_fence_after();
          slot_ready = true;
        }
        if (!(flags & 16)) {
          const uint32_t dst = tmem_a0 + b * Cfg::SLOT_COLS + u * (CPL * 16) + lane_field;
          tc_st<CPL * 8>(dst, hi);
          if (!exact) tc_st<CPL * 8>(dst + CPL * 8, lo);
        }
      }
      if (!(flags & 16)) tc_wait_st();
      tc_fence_before();
      mbar_arrive(tmem_full + 8 * b);

      // a period that ended with the previous stage is drained now, behind this stage's conversion,
      // so that the MMA completion latency is hidden
      if (pending) {
        drain_accumulators(false);
        pending = false;
      }
      if (i + 1 == nstages) drain_accumulators(true);
      else if (++pos == drain) {
        pending = true;
        pos = 0;
      }
      if (++s == Cfg::NS) { s = 0; ph_s ^= 1u; }
      if (++b == Cfg::NT) { b = 0; ph_b ^= 1u; }
    }
    if (nstages == 0 && j < ncols) {
      VecT z;
      float* zf = reinterpret_cast<float*>(&z);
#pragma unroll
      for (int c = 0; c < CPL; ++c) zf[c] = 0.f;
      for (int t = 0; t < KP; ++t) *reinterpret_cast<VecT*>(out + (int64_t)t * ldc) = z;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// [hi | lo] TF32 operand shadow of a component-major fp64 master (kp x ld) in the K-major canonical layout
// the MMA reads: out[((r/4) * 2kp + n) * 4 + r%4], n < kp: hi part of component n, n >= kp: lo part of
// component n - kp; rows padded to a multiple of 32 with zeros.
__global__ void tc_shadow_kernel(const double* __restrict__ F64, int64_t ld, int64_t rows, int k, int kp,
                                 float* __restrict__ out, int64_t rows_pad) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over rows_pad * kp
  if (idx >= rows_pad * kp) return;
  const int e = (int)(idx & 3);
  const int t = (int)((idx >> 2) % kp);
  const int64_t r4 = idx / (4 * (int64_t)kp);
  const int64_t r = r4 * 4 + e;
  float h = 0.f, l = 0.f;
  if (r < rows && t < k) {
    const double v = F64[(int64_t)t * ld + r];
    h = __uint_as_float(__float_as_uint((float)v) & TF32_MASK);
    l = __uint_as_float(__float_as_uint((float)(v - (double)h)) & TF32_MASK);
  }
  out[(r4 * 2 * kp + t) * 4 + e] = h;
  out[(r4 * 2 * kp + kp + t) * 4 + e] = l;
}

// inexact[0] |= 1 if any element of X carries mantissa bits below TF32's 10 (then Xl != 0 somewhere)
__global__ void tf32_exact_kernel(const uint4* __restrict__ X, int64_t n4, int* __restrict__ inexact) {
  uint32_t acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = X[i];
    acc |= (v.x | v.y | v.z | v.w) & ~TF32_MASK;
  }
  if (__any_sync(0xffffffffu, acc != 0) && (threadIdx.x & 31) == 0) atomicOr(inexact, 1);
}

template <int KP>
static int wtx_tc_launch(const float* X, int64_t ldx, int g, int64_t ncols, const float* Fs, float* part,
                         int64_t ldc, int nsplit, int rps, int drain, int flags, const NmfCtrl* ctrl,
                         cudaStream_t st) {
  using Cfg = TcCfg<KP>;
  CUtensorMap xmap;
  if (int rc = make_tensor_map_2d(&xmap, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ldx, g, ncols, 256, Cfg::STAGE_ROWS,
                                   CU_TENSOR_MAP_SWIZZLE_NONE))
    return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wtx_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return SCRNA_ERR_CUDA;
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(ncols, Cfg::COLS), (unsigned)nsplit);
  wtx_tc_kernel<KP><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(xmap, g, ncols, Fs, part, ldc, rps, drain, flags, ctrl);
  return last_launch_status();
}

static int tc_rows_per_split(int g, int nsplit) {
  int rps = (int)ceil_div(g, nsplit);
  return (int)(ceil_div(rps, TC_ROW_ALIGN) * TC_ROW_ALIGN);
}

}  // namespace scrna

using namespace scrna;

#define SCRNA_DISPATCH_TC(kp, CALL)                  \
  switch (kp) {                                      \
    case 16: { constexpr int KP = 16; CALL; } break; \
    case 32: { constexpr int KP = 32; CALL; } break; \
    case 48: { constexpr int KP = 48; CALL; } break; \
    case 64: { constexpr int KP = 64; CALL; } break; \
    default: return SCRNA_ERR_UNSUPPORTED;           \
  }

extern "C" int scrna_nmf_tc_supported(int kp) { return (kp == 16 || kp == 32 || kp == 48 || kp == 64) ? 1 : 0; }

extern "C" int64_t scrna_nmf_tc_shadow_elems(int64_t rows, int kp) {
  return ceil_div(rows, TC_ROW_ALIGN) * TC_ROW_ALIGN * 2 * kp;
}

extern "C" int scrna_nmf_wtx_tc_splits(int g, int64_t ncols, int kp) {
  if (g <= 0 || ncols <= 0 || !scrna_nmf_tc_supported(kp)) return SCRNA_ERR_BADARG;
  const int64_t tiles = ceil_div(ncols, kp <= 32 ? 512 : 256);
  const int64_t slots = (int64_t)num_sms();
  int max_split = g / 256;
  if (max_split > 64) max_split = 64;
  if (max_split < 1) max_split = 1;
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= max_split; ++s) {
    const int rps = tc_rows_per_split(g, s);
    const int eff_s = (int)ceil_div(g, rps);
    if (eff_s != s) continue;  // only split counts whose spans are all non-empty
    const int64_t ctas = tiles * s;
    const int64_t waves = ceil_div(ctas, slots);
    // wave efficiency against the cost of a split (a CTA prologue/drain and one more partial for the epilogue
    // to sum): measured on a 12 500-cell shard, 5-17 splits of 25 tiles beat 23-29 by 3-5 %
    const double score = (double)ctas / (double)(waves * slots) - 0.004 * s;
    if (score > best_score) {
      best_score = score;
      best = s;
    }
  }
  return best;
}

extern "C" int scrna_nmf_wtx_partial_tc(const float* X, int64_t ldx, int g, int64_t ncols, const float* Fs, int kp,
                                        float* part, int64_t ldc, int nsplit, int flags,
                                        const scrna_nmf_ctrl_t* ctrl, void* stream) {
  if (!X || !Fs || !part || g <= 0 || ncols <= 0 || nsplit <= 0 || (ldx & 3) || (ldc & 3) || (ncols & 3) ||
      ncols > ldx || ncols > ldc)
    return SCRNA_ERR_BADARG;
  const int rps = tc_rows_per_split(g, nsplit);
  if ((int64_t)rps * (nsplit - 1) >= g && nsplit > 1) return SCRNA_ERR_BADARG;  // an empty span
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const NmfCtrl* c = reinterpret_cast<const NmfCtrl*>(ctrl);
  int drain = (flags >> 8) & 0xffff;  // tuning override of the drain period (stages of 16 rows)
  if (drain <= 0) drain = kp <= 32 ? 32 : 16;  // 512 rows
  SCRNA_DISPATCH_TC(kp, return wtx_tc_launch<KP>(X, ldx, g, ncols, Fs, part, ldc, nsplit, rps, drain, flags & 0xff, c, st));
  return SCRNA_ERR_UNSUPPORTED;
}

extern "C" int scrna_tf32_inexact(const float* X, int64_t count, int32_t* inexact, void* stream) {
  if (!X || !inexact || count <= 0 || (count & 3)) return SCRNA_ERR_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(inexact, 0, sizeof(int32_t), st) != cudaSuccess) return SCRNA_ERR_CUDA;
  tf32_exact_kernel<<<num_sms() * 8, 256, 0, st>>>(reinterpret_cast<const uint4*>(X), count / 4, inexact);
  return last_launch_status();
}

extern "C" int scrna_nmf_tc_shadow(const double* F64, int64_t ld, int64_t rows, int k, int kp, float* out,
                                   void* stream) {
  if (!F64 || !out || rows <= 0 || k <= 0 || k > kp || !scrna_nmf_tc_supported(kp)) return SCRNA_ERR_BADARG;
  const int64_t rows_pad = ceil_div(rows, TC_ROW_ALIGN) * TC_ROW_ALIGN;
  const int64_t total = rows_pad * kp;
  tc_shadow_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      F64, ld, rows, k, kp, out, rows_pad);
  return last_launch_status();
}
