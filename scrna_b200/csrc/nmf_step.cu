// Fused half steps of the multiplicative-update iteration on the 16-bit path:
//
//   iteration  =  sweep(X^T)  ->  gstep  ->  sweep(X)  ->  cstep                       (four launches, no NCCL)
//
// gstep (gene factor; ONE launch per iteration on every rank, the cross-GPU exchange inside the kernel):
//   1. reduce this rank's sweep partials over the row splits -> its contribution to (X.C^T)^T, and its local
//      C.C^T from the cell epilogue's per-group Gram partials, into the rank's exchange buffer;
//   A. grid barrier + arrival flags stored into every peer's memory (NVLink peer stores, st.release.sys);
//   2. the genes are dealt to the ranks in groups of 128 rows: a rank sums ALL ranks' contributions for its
//      groups straight out of their exchange buffers (peer loads, fixed rank order: deterministic), applies the
//      MU ratio W *= XHt / (W.HHt + l1 + l2 W)  (sklearn _nmf.py:521-626; zero denominators -> float32 eps),
//      and STORES the new rows -- fp64 master, fp32 shadow, scaled fp16 operand shadow + its 2^-e, the group's
//      Gram partial -- into every rank's copy (peer stores): reduce-scatter, epilogue and all-gather in one;
//   B. grid barrier + arrival flags again: every rank's copy of the gene factor is complete;
//   3. G^T.G from the (now complete) per-group partials, for the cell half step.
//   With one rank the same kernel runs with itself as the only peer: the barriers are grid barriers.
// cstep (cell factor; cells are local to the shard, no communication): split-partial reduction, MU ratio,
//   shadows and per-group Gram partials in one launch, one CTA per 128-cell group.
//
// Every rank reads a peer's exchange buffer only between barriers A and B of the same call and writes it
// before A, so one buffer per rank suffices; flags and barrier counters only grow (sequence numbers), nothing
// is ever reset.  All waits are bounded and trap.  The peers' arenas come from torch's symmetric memory
// (same layout on every rank); `peers[q]` is the base address of rank q's arena in this rank's address space.
#include <cuda_fp16.h>

#include "common.cuh"

namespace scrna {

constexpr int ST_THREADS = 1024;  // up to 8 independent sub-blocks of 128 threads; a sub-block = one 128-row group
constexpr int ST_GROUP = 128;     // rows per scale group (= nmf_sweep_h.cu H_GROUP) = threads per sub-block
constexpr int ST_GLD = ST_GROUP + 1;
constexpr int ST_MAXSUB = ST_THREADS / ST_GROUP;
constexpr int ST_MAXPEERS = 16;
constexpr int ST_SMEM_BUDGET = 190 * 1024;

struct StepDst {          // one destination copy of a factor (own or a peer's)
  double* F64;            // kq x ld component-major master
  float* S32;             // rows x kq row-major fp32 shadow
  float* T32;             // kq x ld component-major fp32 shadow (cell factor only; may be null)
  uint4* Fh;              // fp16 operand shadow, 16-byte vectors [(r/8)][nb]
  float* inv;             // groups x kq
  double* gpart;          // groups x kq*kq Gram partials
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all CTAs of the (co-resident) grid; `counter` only grows: the caller passes the value it must reach.  ONE thread
// per CTA fences, at system scope: after the CTA barrier its fence is cumulative over the whole CTA's writes, so
// they are visible to the peers before the arrival flags stored after this barrier (a membar.sys by every thread
// measured 10x the cost: ncu showed the kernel waiting in ERRBAR).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu(counter) - target) < 0)
      if (clock64() - t0 > 4000000000LL) __trap();
  }
  __syncthreads();
}
__device__ __forceinline__ void sub_barrier(int sub) {   // the 128 threads of one sub-block
  asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "r"(ST_GROUP) : "memory");
}

// concurrent sub-blocks per CTA: each needs two kq x 129 fp64 panels (old rows, new rows) next to the kq x kq Gram
__host__ __device__ inline int step_nsub(int kq) {
  int n = (ST_SMEM_BUDGET - kq * kq * 8) / (2 * kq * ST_GLD * 8);
  return n > ST_MAXSUB ? ST_MAXSUB : (n < 1 ? 1 : n);
}

// One 128-row group of a factor, by ONE sub-block (128 threads, thread i = row i of the group, its k entries walked
// sequentially exactly like sklearn's row-wise arithmetic): MU ratio, group-local fp16 scales, shadows, Gram partial;
// results are stored into `ndst` destination copies.  sw / sv: this sub-block's kq x 129 panels.
//   num(t, row) -> numerator (X.H^T)[row][t];  gram_s: kq x kq of the OTHER factor (shared by the CTA).
template <class NumFn>
__device__ __forceinline__ void mu_group(int sub, double* sw, double* sv, float* sscale, const double* gram_s, int64_t grp,
                                         int64_t rows, int k, int kq, int nb, int parts, double l1, double l2,
                                         const double* __restrict__ Fold, int64_t ld, NumFn num, const StepDst* dst,
                                         int ndst, int* __restrict__ nan_flag) {
  const int i = threadIdx.x & (ST_GROUP - 1), lane = i & 31, warp = i >> 5;
  const int64_t row0 = grp * ST_GROUP, row = row0 + i;
  const bool live = row < rows;
  // every load of the row is issued before anything is consumed
  for (int t = 0; t < k; ++t) {
    sw[t * ST_GLD + i] = live ? Fold[(int64_t)t * ld + row] : 0.0;
    sv[t * ST_GLD + i] = live ? num(t, row) : 0.0;          // numerator, replaced by the new value below
  }
  bool nan_seen = false;
  for (int t0 = 0; t0 < k; t0 += 8) {
    double den[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) den[u] = 0.0;
    for (int q = 0; q < k; ++q) {
      const double wq = sw[q * ST_GLD + i];
      const double* gq = gram_s + q * kq + t0;   // kq is a multiple of 8 and the Gram is zero-padded
#pragma unroll
      for (int u = 0; u < 8; ++u) den[u] += wq * gq[u];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + u;
      if (t < k) {
        const double w = sw[t * ST_GLD + i];
        double d = den[u];
        if (l1 > 0.0) d += l1;
        if (l2 > 0.0) d = d + l2 * w;
        if (d == 0.0) d = (double)kEpsMU;
        const double v = live ? w * (sv[t * ST_GLD + i] / d) : 0.0;
        nan_seen |= (v != v);
        sv[t * ST_GLD + i] = v;
      }
    }
  }
  if (nan_seen) atomicOr(nan_flag, SCRNA_FLAG_NAN);
  sub_barrier(sub);
  // group-local scale per component: largest |entry| of the 128 rows (one warp per component)
  for (int t = warp; t < kq; t += ST_GROUP / 32) {
    double mx = 0.0;
    if (t < k)
#pragma unroll
      for (int j = 0; j < ST_GROUP / 32; ++j) mx = fmax(mx, fabs(sv[t * ST_GLD + lane + 32 * j]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      // mx * 2^e in [2^13, 2^14); e clamped so that 2^e and 2^-e stay inside fp32
      int ex = 0;
      if (mx > 0.0 && mx < 1e300) {
        ex = 14 - (ilogb(mx) + 1);
        ex = ex > 114 ? 114 : (ex < -100 ? -100 : ex);
      }
      sscale[t] = __int_as_float((127 + ex) << 23);
      const float inv = __int_as_float((127 - ex) << 23);
      for (int d = 0; d < ndst; ++d) dst[d].inv[grp * kq + t] = inv;
    }
  }
  sub_barrier(sub);
  // masters (component-major: the sub-block's threads are consecutive rows) and the row-major fp32 shadow
  if (live) {
    for (int t = 0; t < k; ++t) {
      const double v = sv[t * ST_GLD + i];
      for (int d = 0; d < ndst; ++d) {
        dst[d].F64[(int64_t)t * ld + row] = v;
        if (dst[d].T32) dst[d].T32[(int64_t)t * ld + row] = (float)v;
      }
    }
    for (int t0 = 0; t0 < kq; t0 += 4) {
      float4 o;
      o.x = t0 + 0 < k ? (float)sv[(t0 + 0) * ST_GLD + i] : 0.f;
      o.y = t0 + 1 < k ? (float)sv[(t0 + 1) * ST_GLD + i] : 0.f;
      o.z = t0 + 2 < k ? (float)sv[(t0 + 2) * ST_GLD + i] : 0.f;
      o.w = t0 + 3 < k ? (float)sv[(t0 + 3) * ST_GLD + i] : 0.f;
      for (int d = 0; d < ndst; ++d) *reinterpret_cast<float4*>(dst[d].S32 + row * kq + t0) = o;
    }
  }
  // fp16 operand shadow: thread (8-row group rg = i / 8, component t = i % 8 + 8 j): consecutive threads write
  // consecutive 16-byte vectors; rows past the factor are zero
  {
    const int rg = i >> 3;
    const int64_t base = ((row0 >> 3) + rg) * nb;
    for (int t = i & 7; t < kq; t += 8) {
      union Pack { uint4 v; __half h[8]; } p1, p2, p3;
      const float sc = t < k ? sscale[t] : 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        double v = t < k ? sv[t * ST_GLD + rg * 8 + e] * (double)sc : 0.0;
        p1.h[e] = __double2half(v);
        v -= (double)__half2float(p1.h[e]);
        p2.h[e] = __double2half(v);
        v -= (double)__half2float(p2.h[e]);
        p3.h[e] = __double2half(v);
      }
      for (int d = 0; d < ndst; ++d) {
        uint4* o = dst[d].Fh + base;
        o[t] = p1.v;
        o[kq + t] = p2.v;
        if (parts == 3) o[2 * kq + t] = p3.v;
        if (parts * kq + t < nb) o[parts * kq + t] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  // Gram partial of the group's new rows: entry (a, b) by one thread, four independent chains over the rows
  for (int o = i; o < kq * kq; o += ST_GROUP) {
    const int a = o / kq, b = o - a * kq;
    double s = 0.0;
    if (a < k && b < k) {
      const double* ra = sv + a * ST_GLD;
      const double* rb = sv + b * ST_GLD;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
      for (int r = 0; r < ST_GROUP; r += 4) {
        s0 = fma(ra[r], rb[r], s0);
        s1 = fma(ra[r + 1], rb[r + 1], s1);
        s2 = fma(ra[r + 2], rb[r + 2], s2);
        s3 = fma(ra[r + 3], rb[r + 3], s3);
      }
      s = (s0 + s1) + (s2 + s3);
    }
    for (int d = 0; d < ndst; ++d) dst[d].gpart[grp * (int64_t)kq * kq + o] = s;
  }
  sub_barrier(sub);   // the panels are reused by this sub-block's next group
}

struct GStepArgs {
  const unsigned long long* peers;   // device array: base address of every rank's arena (own included)
  int nranks, rank;
  long long off_xch;     // fp64 [kq*ldg | kq*kq]: this rank's contributions
  long long off_sync;    // u32 [ seq | grid counter | pad,pad | flagsA[16] | flagsB[16] ]
  long long off_G64, off_Gs32, off_Gh, off_Ginv, off_gpart, off_gtg;
  int g, k, kq, nb, parts;
  long long ldg;
  double l1, l2;
};

__global__ void __launch_bounds__(ST_THREADS, 1)
gstep_mu_kernel(GStepArgs a, const float* __restrict__ PA, int SA, const double* __restrict__ gpartC, int groupsC,
                int* __restrict__ nan_flag) {
  extern __shared__ double st_smem[];
  __shared__ StepDst dst[ST_MAXPEERS];
  __shared__ float sscale_all[ST_MAXSUB][SCRNA_MAX_K];
  const int kq = a.kq, k = a.k;
  const int tid = threadIdx.x, sub = tid >> 7, nsub = step_nsub(kq);
  double* gram_s = st_smem;                                          // kq x kq   C.C^T summed over the ranks
  double* sw = gram_s + kq * kq + (size_t)sub * 2 * kq * ST_GLD;    // this sub-block's kq x 129 old rows
  double* sv = sw + (size_t)kq * ST_GLD;                            // ... and new rows
  char* me = reinterpret_cast<char*>(a.peers[a.rank]);
  unsigned int* sync = reinterpret_cast<unsigned int*>(me + a.off_sync);
  const unsigned int seq = sync[0];               // calls completed so far (bumped by CTA 0 at the very end)
  const unsigned int nct = gridDim.x;
  double* xch = reinterpret_cast<double*>(me + a.off_xch);
  const int64_t xn = (int64_t)kq * a.ldg;

  // ---- 1. local contributions -----------------------------------------------------------------------------
  for (int64_t i = (int64_t)blockIdx.x * ST_THREADS + tid; i < xn; i += (int64_t)nct * ST_THREADS) {
    double s = 0.0;
    for (int sp = 0; sp < SA; ++sp) s += (double)PA[(int64_t)sp * xn + i];   // fixed order
    xch[i] = s;
  }
  {
    const int warp_g = blockIdx.x * (ST_THREADS / 32) + (tid >> 5), lane = tid & 31;
    for (int o = warp_g; o < kq * kq; o += nct * (ST_THREADS / 32)) {
      const int r = o / kq, c = o - r * kq;
      double s = 0.0;
      if (r < k && c < k)
        for (int p = lane; p < groupsC; p += 32) s += gpartC[(int64_t)p * kq * kq + o];
      s = warp_sum(s);
      if (lane == 0) xch[xn + o] = s;
    }
  }
  // ---- A. everybody's contribution is in place ----------------------------------------------------------------
  grid_barrier(sync + 1, (2 * seq + 1) * nct);
  if (blockIdx.x == 0 && tid < a.nranks)
    st_release_sys(reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(a.peers[tid]) + a.off_sync) + 4 + a.rank, seq + 1);
  if (tid < a.nranks) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(sync + 4 + tid) - (seq + 1)) < 0)
      if (clock64() - t0 > 8000000000LL) __trap();
  }
  __syncthreads();

  // ---- 2. this rank's gene groups ----------------------------------------------------------------------------
  for (int o = tid; o < kq * kq; o += ST_THREADS) {
    double s = 0.0;
    for (int q = 0; q < a.nranks; ++q)
      s += reinterpret_cast<const double*>(reinterpret_cast<const char*>(a.peers[q]) + a.off_xch)[xn + o];
    gram_s[o] = s;
  }
  if (tid < a.nranks) {
    char* b = reinterpret_cast<char*>(a.peers[tid]);
    dst[tid].F64 = reinterpret_cast<double*>(b + a.off_G64);
    dst[tid].S32 = reinterpret_cast<float*>(b + a.off_Gs32);
    dst[tid].T32 = nullptr;
    dst[tid].Fh = reinterpret_cast<uint4*>(b + a.off_Gh);
    dst[tid].inv = reinterpret_cast<float*>(b + a.off_Ginv);
    dst[tid].gpart = reinterpret_cast<double*>(b + a.off_gpart);
  }
  __syncthreads();
  const int64_t groups = ((int64_t)a.g + ST_GROUP - 1) / ST_GROUP;
  const int64_t per = (groups + a.nranks - 1) / a.nranks;
  const int64_t first = (int64_t)a.rank * per;
  const int64_t last = first + per < groups ? first + per : groups;
  const double* Gold = reinterpret_cast<const double*>(me + a.off_G64);
  const unsigned long long* peers = a.peers;
  const int nr = a.nranks;
  const long long off_xch = a.off_xch, ldg = a.ldg;
  auto num = [=](int t, int64_t row) {
    double s = 0.0;
    for (int q = 0; q < nr; ++q)   // fixed rank order: every rank would compute the same bits
      s += reinterpret_cast<const double*>(reinterpret_cast<const char*>(peers[q]) + off_xch)[(int64_t)t * ldg + row];
    return s;
  };
  // groups are dealt to the CTAs first and to their sub-blocks second: a few groups spread over many SMs
  if (sub < nsub)
    for (int64_t grp = first + blockIdx.x + (int64_t)sub * nct; grp < last; grp += (int64_t)nct * nsub)
      mu_group(sub, sw, sv, sscale_all[sub], gram_s, grp, a.g, k, kq, a.nb, a.parts, a.l1, a.l2, Gold, a.ldg, num, dst,
               a.nranks, nan_flag);

  // ---- B. every rank's copy of the gene factor is complete ----------------------------------------------------
  grid_barrier(sync + 1, (2 * seq + 2) * nct);
  if (blockIdx.x == 0 && tid < a.nranks)
    st_release_sys(reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(a.peers[tid]) + a.off_sync) + 20 + a.rank, seq + 1);
  if (tid < a.nranks) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(sync + 20 + tid) - (seq + 1)) < 0)
      if (clock64() - t0 > 8000000000LL) __trap();
  }
  __syncthreads();

  // ---- 3. G^T.G for the cell half step ------------------------------------------------------------------------
  {
    const double* gp = reinterpret_cast<const double*>(me + a.off_gpart);
    double* gtg = reinterpret_cast<double*>(me + a.off_gtg);
    const int warp_g = blockIdx.x * (ST_THREADS / 32) + (tid >> 5), lane = tid & 31;
    for (int o = warp_g; o < kq * kq; o += nct * (ST_THREADS / 32)) {
      const int r = o / kq, c = o - r * kq;
      double s = 0.0;
      if (r < k && c < k)
        for (int64_t p = lane; p < groups; p += 32) s += gp[p * kq * kq + o];
      s = warp_sum(s);
      if (lane == 0) gtg[o] = s;
    }
  }
  // every CTA has read `seq` (they all passed barrier B): publish the next sequence number for the next launch
  if (blockIdx.x == 0 && tid == 0) sync[0] = seq + 1;
}

__global__ void __launch_bounds__(ST_THREADS, 1)
cstep_mu_kernel(double* __restrict__ C64, int64_t ldn, int64_t n, int k, int kq, int nb, int parts,
                const float* __restrict__ PB, int SB, const double* __restrict__ gtg, double l1, double l2,
                float* __restrict__ Cs32, float* __restrict__ Ct32, uint4* __restrict__ Ch, float* __restrict__ Cinv,
                double* __restrict__ gpartC, int* __restrict__ nan_flag) {
  extern __shared__ double st_smem[];
  __shared__ StepDst dst[1];
  __shared__ float sscale_all[ST_MAXSUB][SCRNA_MAX_K];
  const int tid = threadIdx.x, sub = tid >> 7, nsub = step_nsub(kq);
  double* gram_s = st_smem;
  double* sw = gram_s + kq * kq + (size_t)sub * 2 * kq * ST_GLD;
  double* sv = sw + (size_t)kq * ST_GLD;
  for (int o = tid; o < kq * kq; o += ST_THREADS) gram_s[o] = gtg[o];
  if (tid == 0) {
    dst[0].F64 = C64; dst[0].S32 = Cs32; dst[0].T32 = Ct32; dst[0].Fh = Ch; dst[0].inv = Cinv; dst[0].gpart = gpartC;
  }
  __syncthreads();
  const int64_t stride = (int64_t)kq * ldn;
  auto num = [=](int t, int64_t row) {
    double s = 0.0;
    const float* p = PB + (int64_t)t * ldn + row;
    for (int sp = 0; sp < SB; ++sp) s += (double)p[(int64_t)sp * stride];   // fixed order
    return s;
  };
  const int64_t groups = (n + ST_GROUP - 1) / ST_GROUP;
  if (sub < nsub)
    for (int64_t grp = blockIdx.x + (int64_t)sub * gridDim.x; grp < groups; grp += (int64_t)gridDim.x * nsub)
      mu_group(sub, sw, sv, sscale_all[sub], gram_s, grp, n, k, kq, nb, parts, l1, l2, C64, ldn, num, dst, 1, nan_flag);
}

static size_t step_smem(int kq) {
  return ((size_t)kq * kq + (size_t)step_nsub(kq) * 2 * kq * ST_GLD) * sizeof(double);
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_nmf_step_groups(int64_t rows) { return rows > 0 ? (int)ceil_div(rows, ST_GROUP) : SCRNA_ERR_BADARG; }

extern "C" int scrna_nmf_gstep_mu(const void* peers, int nranks, int rank, int64_t off_xch, int64_t off_sync,
                                  int64_t off_G64, int64_t off_Gs32, int64_t off_Gh, int64_t off_Ginv,
                                  int64_t off_gpart, int64_t off_gtg, int g, int64_t ldg, int k, int kq, int nb,
                                  int parts, double l1, double l2, const float* PA, int SA, const double* gpartC,
                                  int groupsC, int32_t* nan_flag, void* stream) {
  if (!peers || nranks < 1 || nranks > ST_MAXPEERS || rank < 0 || rank >= nranks || !PA || SA <= 0 || !gpartC ||
      groupsC <= 0 || !nan_flag || g <= 0 || k <= 0 || k > kq || kq > SCRNA_MAX_K || (kq & 7) || ldg < g)
    return SCRNA_ERR_BADARG;
  GStepArgs a;
  a.peers = reinterpret_cast<const unsigned long long*>(peers);
  a.nranks = nranks; a.rank = rank;
  a.off_xch = off_xch; a.off_sync = off_sync; a.off_G64 = off_G64; a.off_Gs32 = off_Gs32; a.off_Gh = off_Gh;
  a.off_Ginv = off_Ginv; a.off_gpart = off_gpart; a.off_gtg = off_gtg;
  a.g = g; a.k = k; a.kq = kq; a.nb = nb; a.parts = parts; a.ldg = ldg; a.l1 = l1; a.l2 = l2;
  const size_t smem = step_smem(kq);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(gstep_mu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return SCRNA_ERR_CUDA;
    attr = true;
  }
  // one co-resident CTA per SM: the kernel synchronises its own grid (the stream is otherwise idle between sweeps)
  gstep_mu_kernel<<<num_sms(), ST_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a, PA, SA, gpartC, groupsC,
                                                                                         nan_flag);
  return last_launch_status();
}

extern "C" int scrna_nmf_cstep_mu(double* C64, int64_t ldn, int64_t n, int k, int kq, int nb, int parts,
                                  const float* PB, int SB, const double* gtg, double l1, double l2, float* Cs32,
                                  float* Ct32, void* Ch, float* Cinv, double* gpartC, int32_t* nan_flag, void* stream) {
  if (!C64 || !PB || SB <= 0 || !gtg || !Cs32 || !Ch || !Cinv || !gpartC || !nan_flag || n <= 0 || ldn < n || k <= 0 ||
      k > kq || kq > SCRNA_MAX_K || (kq & 7))
    return SCRNA_ERR_BADARG;
  const size_t smem = step_smem(kq);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(cstep_mu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return SCRNA_ERR_CUDA;
    attr = true;
  }
  // one CTA per SM at most; its sub-blocks walk independent groups (all SMs busy before any CTA takes a second group)
  int64_t grid = ceil_div(n, ST_GROUP);
  if (grid > num_sms()) grid = num_sms();
  cstep_mu_kernel<<<(unsigned)grid, ST_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      C64, ldn, n, k, kq, nb, parts, PB, SB, gtg, l1, l2, Cs32, Ct32, reinterpret_cast<uint4*>(Ch), Cinv, gpartC, nan_flag);
  return last_launch_status();
}
