/* libscrna_b200 -- C ABI of the B200-native (sm_100a) scRNA clustering hot path.
 *
 * The reference (nicococo/scRNA) is pure Python and has NO FFI of its own: its hot
 * path is reached through Python classes that delegate the arithmetic to scikit-learn / scipy /
 * numpy.  Each entry point below therefore cites the reference call site (file:line under
 * /root/reference, or SP/ = the installed scikit-learn / scipy the reference calls into) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - every function returns SCRNA_OK (0) or a negative SCRNA_ERR_* code; nothing throws;
 *   - all data pointers are DEVICE pointers owned by the caller; the library never allocates;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call is
 *     asynchronous on that stream and safe to capture in a CUDA graph;
 *   - X is always genes x cells, row-major fp32, leading dimension `ldx` (elements, multiple of 4,
 *     columns [n, ldx) zero-filled) -- the reference layout (abstract_clustering.py:28-29) narrowed
 *     to fp32 storage;
 *   - factors: gene factor G is g x kp row-major, cell factor C is kp x ldc row-major, with
 *     kp = k rounded up to a multiple of 4 and padding rows/columns kept at zero.  Both exist as an
 *     fp64 master (updated by the epilogues) and an fp32 shadow (read by the X sweeps);
 *   - `ctrl` is a 64-byte device control block (scrna_nmf_ctrl_t).  When ctrl->done != 0 every
 *     kernel of the iteration returns immediately, so a loop can be enqueued (or graph-replayed)
 *     past convergence without host synchronisation.  ctrl may be NULL for one-shot calls.
 */
#ifndef SCRNA_B200_H
#define SCRNA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCRNA_OK 0
#define SCRNA_ERR_BADARG (-1)
#define SCRNA_ERR_CUDA (-2)
#define SCRNA_ERR_UNSUPPORTED (-3)

#define SCRNA_FLAG_NAN 1       /* a factor entry became NaN (nmf_clustering.py:33-38)          */
#define SCRNA_FLAG_ZERO_DEN 2  /* zero denominator in the target MU step (utils.py:199-200)    */

#define SCRNA_SOLVER_MU 0 /* SP/sklearn/decomposition/_nmf.py:521-723 (Frobenius multiplicative update) */
#define SCRNA_SOLVER_CD 1 /* SP/sklearn/decomposition/_cdnmf_fast.pyx:8-38 (coordinate descent)         */

#define SCRNA_MAX_K 64

typedef struct scrna_nmf_ctrl {
  int32_t iter;     /* completed iterations                                   */
  int32_t done;     /* 1 once the stop rule fired                             */
  int32_t n_iter;   /* iteration at which it fired (sklearn's n_iter)         */
  int32_t flags;    /* SCRNA_FLAG_*                                           */
  double viol_init; /* CD: violation of iteration 1 (_nmf.py:504-505)         */
  double viol_last; /* CD: violation of the last completed iteration          */
  double err_last;  /* MU / transfer: last error scalar                       */
  double reserved[3];
} scrna_nmf_ctrl_t;

/* ---- library / device ------------------------------------------------------------------ */
/* Version of the ABI (bumped on any signature change). */
int scrna_abi_version(void);
/* Fills sm_count / max dynamic smem per block of the current device. */
int scrna_device_info(int* sm_count, int* max_smem_per_block, int* cc_major, int* cc_minor);

/* ---- NMF iteration (NmfClustering.apply, nmf_clustering.py:20-42; NmfClustering_initW.apply,
 *      nmf_clustering.py:59-85; DaNmfClustering.apply :207-223) ------------------------------
 * One sklearn iteration (_nmf.py:491-502) = gene-factor half step + cell-factor half step, each
 *   sweep of X (K1 / K3)  ->  reduce partials  ->  [all-reduce across cell shards]  ->  epilogue.
 */

/* Recommended split counts for the two sweeps on the current device (grid = tiles x splits is
 * sized to a multiple of SM count x resident CTAs). */
int scrna_nmf_xht_splits(int g, int64_t ncols, int kp);
int scrna_nmf_wtx_splits(int g, int64_t ncols, int kp);

/* K1 xht_partial: part[s][i][t] = sum_{j in cell span s} X[i][j] * C32[t][j]   (X . C^T, reduces
 * over the contiguous axis).  Replaces safe_sparse_dot(X, Ht) in _update_coordinate_descent
 * (_nmf.py:381) / the numerator X.H^T of _multiplicative_update_w (_nmf.py:537-538).
 * part: nsplit x g x kp fp32.  variant: 0 = default kernel. */
int scrna_nmf_xht_partial(const float* X, int64_t ldx, int g, int64_t ncols, const float* C32,
                          int64_t ldc, int kp, float* part, int nsplit, int variant,
                          const scrna_nmf_ctrl_t* ctrl, void* stream);

/* K3 wtx_partial: part[s][t][j] = sum_{i in gene span s} G32[i][t] * X[i][j]   (G^T . X, cells own
 * their output).  Replaces safe_sparse_dot(X.T, W) (_nmf.py:381 via :499-501) / W^T.X of
 * _multiplicative_update_h (_nmf.py:644) / W.T.dot(trg_data) (utils.py:201).
 * part: nsplit x kp x ldc fp32.  The kernel is orientation-agnostic: called on the resident
 * transposed copy X^T (cells x genes) with the cell factor's row-major shadow it yields (X.C^T)^T,
 * i.e. K1 without a cross-lane reduction -- this is the default K1 path.
 * variant 0: scalar FFMA inner loop; variant >= 1: packed fma.rn.f32x2 (FFMA2, sm_100), bit-identical. */
int scrna_nmf_wtx_partial(const float* X, int64_t ldx, int g, int64_t ncols, const float* G32, int kp,
                          float* part, int64_t ldc, int nsplit, int variant,
                          const scrna_nmf_ctrl_t* ctrl, void* stream);

/* K1/K3 on the tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory, X staged by bulk
 * TMA): same contract and partial layout as scrna_nmf_wtx_partial for kp in {16, 32, 48, 64}, where the
 * FP32 pipe can no longer keep up with HBM.  The factor is passed as its 3xTF32 operand shadow Fs
 * ([hi | lo], scrna_nmf_tc_shadow_elems(g, kp) floats, K-major MMA operand layout
 * [(r/4)][n < 2kp][r%4]) written by scrna_nmf_tc_shadow; products are Xh.[Fh|Fl] + Xl.Fh accumulated in
 * fp32 and drained from tensor memory every 32 stages (512 rows) so that the accumulator's truncation
 * bias stays at the fp32 rounding level.
 * nsplit must come from scrna_nmf_wtx_tc_splits (every row span non-empty, spans multiples of 16 rows).
 * flags: 32 = X is TF32-exact (scrna_tf32_inexact reported 0); bits 1-4: bring-up probes; bits 8-23: drain
 * period override in stages. */
int scrna_nmf_tc_supported(int kp);
int64_t scrna_nmf_tc_shadow_elems(int64_t rows, int kp);
int scrna_nmf_wtx_tc_splits(int g, int64_t ncols, int kp);
int scrna_nmf_wtx_partial_tc(const float* X, int64_t ldx, int g, int64_t ncols, const float* Fs, int kp,
                             float* part, int64_t ldc, int nsplit, int flags, const scrna_nmf_ctrl_t* ctrl,
                             void* stream);
/* inexact[0] = 1 if any of the `count` (multiple of 4) fp32 values at X carries mantissa bits below TF32's
 * 10, else 0.  Raw counts < 2048 are exact: then Xl == 0 and the caller passes flag 32 to
 * scrna_nmf_wtx_partial_tc, which skips the Xl product. */
int scrna_tf32_inexact(const float* X, int64_t count, int32_t* inexact, void* stream);
/* [hi | lo] TF32 operand shadow of a component-major fp64 master (kp x ld, element [t*ld + r]). */
int scrna_nmf_tc_shadow(const double* F64, int64_t ld, int64_t rows, int k, int kp, float* Fs, void* stream);

/* K1/K3 on 16-bit storage (tcgen05.mma kind::f16 fed straight from the TMA-staged tile): when every element
 * of X is exactly representable in fp16 (raw counts <= 2048; scrna_f16_inexact reports 0) X and X^T are kept
 * as fp16, halving the bytes of the HBM-bound iteration.
 * Storage is PANEL-MAJOR: panels of 256 columns, element (i, j) at ((j/256)*rows_pad + i)*256 + j%256 with
 * rows_pad = scrna_f16_panel_rows(rows) (rows rounded up to 64) and all padding zero -- the rows a CTA streams
 * are one contiguous region of HBM per panel; scrna_f16_panel_elems(rows, cols) halves, 128-byte aligned base;
 * written by scrna_narrow_f32_f16 / scrna_transpose_f32_f16.
 * Same contract and partial layout as scrna_nmf_wtx_partial with kp = kq = k rounded up to 8.  The factor is
 * passed as its `parts`-part (2 or 3) fp16 operand shadow Fh (scrna_nmf_h_shadow_elems(g, kq, parts) halves,
 * K-major MMA operand layout [(r/8)][n < NB][r%8], n = part*kq + t, NB = scrna_nmf_h_nb) of the factor scaled per
 * component and per group of 128 rows by a power of two; inv_scale[(r/128)*kq + t] (scrna_nmf_h_groups(g) x kq
 * floats) undoes the scale when tensor memory is drained (every 128 rows, into register totals).  Both are
 * written by scrna_nmf_h_shadow / scrna_nmf_update / the fused step kernels.  x * part is exact in the fp32
 * accumulator.  nsplit from scrna_nmf_wtx_h_splits (row spans are multiples of 128). */
int scrna_nmf_h_supported(int kq);
int scrna_nmf_h_nb(int kq, int parts);
int64_t scrna_nmf_h_shadow_elems(int64_t rows, int kq, int parts);
int64_t scrna_nmf_h_groups(int64_t rows);
int64_t scrna_f16_panel_rows(int64_t rows);
int64_t scrna_f16_panel_elems(int64_t rows, int64_t cols);
int scrna_nmf_wtx_h_splits(int g, int64_t ncols, int kq, int parts);
int scrna_nmf_wtx_partial_h(const void* X16, int64_t rows_pad, int g, int64_t ncols, const void* Fh,
                            const float* inv_scale, int kq, int parts, float* part, int64_t ldc, int nsplit,
                            int flags, const scrna_nmf_ctrl_t* ctrl, void* stream);
/* fp16 operand shadow of a component-major fp64 master: per component and 128-row group, scale = 2^e with
 * max|F| * scale in [2^13, 2^14).  gram_out != NULL additionally finishes a kp x kp Gram (fixed-order sum of the
 * `nblk` partials gram_parts[p*kp*kp + ...]) in the same launch. */
int scrna_nmf_h_shadow(const double* F64, int64_t ld, int64_t rows, int k, int kq, int parts,
                       const double* gram_parts, int nblk, int kp, double* gram_out, void* Fh, float* inv_scale,
                       const scrna_nmf_ctrl_t* ctrl, void* stream);
/* inexact[0] = 1 if any of the `count` (multiple of 4) fp32 values is not exactly representable in fp16. */
int scrna_f16_inexact(const float* X, int64_t count, int32_t* inexact, void* stream);
/* row-major fp32 (rows x cols, ld_src) -> panel-major fp16 storage of the matrix, and of its transpose
 * (cols x rows); dst holds scrna_f16_panel_elems(rows, cols) resp. (cols, rows) halves, padding zeroed here. */
int scrna_narrow_f32_f16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* dst, void* stream);
int scrna_transpose_f32_f16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* dst, void* stream);

/* Fused half steps of the MU iteration on the 16-bit path: iteration = sweep(X^T) -> gstep -> sweep(X) -> cstep.
 * gstep (gene factor, one launch per iteration on every rank, the cross-GPU exchange INSIDE the kernel): reduces
 * this rank's sweep partials PA (SA x kq x ldg) and its cell-side Gram partials gpartC (groupsC x kq*kq) into the
 * rank's exchange buffer, synchronises with the peers through arrival flags stored into their memory, sums all
 * ranks' contributions for ITS gene groups (128 rows each, dealt contiguously over the ranks) with peer loads,
 * applies W *= XHt / (W.HHt + l1 + l2 W) (sklearn _nmf.py:521-626) and stores the new rows -- fp64 master, fp32
 * shadow, fp16 operand shadow + 2^-e, per-group Gram partial -- into EVERY rank's arena with peer stores, then
 * finishes G^T.G locally.  `peers`: device array of nranks 64-bit base addresses of the ranks' arenas (identical
 * layout; torch symmetric memory; one rank: its own arena); off_*: byte offsets inside an arena: xch fp64
 * [kq*ldg + kq*kq], sync u32[64] (zeroed once), G64 fp64 [kq*ldg], Gs32 fp32 [g*kq], Gh (scrna_nmf_h_shadow_elems
 * halves), Ginv fp32 [groups*kq], gpart fp64 [groups*kq*kq], gtg fp64 [kq*kq]; groups = scrna_nmf_step_groups(g).
 * cstep (cell factor, local): split-partial reduction, MU ratio, all shadows and per-group Gram partials. */
int scrna_nmf_step_groups(int64_t rows);
int scrna_nmf_gstep_mu(const void* peers, int nranks, int rank, int64_t off_xch, int64_t off_sync, int64_t off_G64,
                       int64_t off_Gs32, int64_t off_Gh, int64_t off_Ginv, int64_t off_gpart, int64_t off_gtg, int g,
                       int64_t ldg, int k, int kq, int nb, int parts, double l1, double l2, const float* PA, int SA,
                       const double* gpartC, int groupsC, int32_t* nan_flag, void* stream);
int scrna_nmf_cstep_mu(double* C64, int64_t ldn, int64_t n, int k, int kq, int nb, int parts, const float* PB, int SB,
                       const double* gtg, double l1, double l2, float* Cs32, float* Ct32, void* Ch, float* Cinv,
                       double* gpartC, int32_t* nan_flag, void* stream);

/* Sum the K1 partials over the splits in fixed order into fp64, component-major: out[t*ld + i]. */
int scrna_nmf_reduce_rows(const float* part, int nsplit, int g, int kp, double* out, int64_t ld,
                          const scrna_nmf_ctrl_t* ctrl, void* stream);
/* Sum the K3 partials over the splits in fixed order into fp64: out[t][j] (kp x ldc). */
int scrna_nmf_reduce_cols(const float* part, int nsplit, int kp, int64_t ldc, int64_t ncols,
                          double* out, const scrna_nmf_ctrl_t* ctrl, void* stream);

/* K4 gram: out[a][b] = sum_r F[r*rs + a*cs] * F[r*rs + b*cs], a,b < k, written as kp x kp fp64
 * (padding zero).  HHt = Ht^T.Ht / W^T.W of _nmf.py:380,536.  `partials` is scratch of
 * scrna_nmf_gram_blocks() * kp*kp doubles; the result is summed in fixed block order. */
int scrna_nmf_gram_blocks(void);
int scrna_nmf_gram(const double* F, int64_t rs, int64_t cs, int64_t rows, int k, int kp,
                   double* partials, double* out, const scrna_nmf_ctrl_t* ctrl, void* stream);

/* K2 / K3-epilogue: update one factor in place.  The fp64 master F64 and the products `num` are
 * component-major (kp x ld: element [t*ld + r], r < rows, t < k).
 *   gram : kp x kp fp64 Gram of the OTHER factor;
 *   solver MU: F *= num / (F.gram + l1 + l2*F), zero denominators -> float32 eps
 *              (_nmf.py:521-626, 629-723);
 *   solver CD: num -= l1; gram.diag += l2; for t in perm: projected-gradient step
 *              (_nmf.py:369-396 + _cdnmf_fast.pyx:8-38); perm = perms[(ctrl->iter*perm_stride +
 *              perm_offset) * kp .. +k) is the pre-drawn RandomState stream;
 *   S32  : fp32 row-major shadow (rows x kp, padding components zero) read by the next sweep;
 *   T32  : optional fp32 component-major shadow (kp x ld) read by scrna_residual / K1 variant 0;
 *   TC32 : optional [hi | lo] TF32 operand shadow read by scrna_nmf_wtx_partial_tc (rows and components
 *          beyond `rows` / `k` must have been zeroed once, e.g. by scrna_nmf_tc_shadow);
 *   Fh, inv_scale, kq, parts : optional fp16 operand shadow + per-component unscale factors read by
 *          scrna_nmf_wtx_partial_h (kq = k rounded up to 8); needs gram_partials -- the Gram finish and the
 *          shadow share one launch (scrna_nmf_h_shadow);
 *   viol_part: scrna_nmf_update_blocks(rows) doubles, per-block sums of |projected gradient|.
 */
int scrna_nmf_update_blocks(int64_t rows);
/* num_parts != NULL: the numerator is the fixed-order fp64 sum of `nsplit` fp32 sweep partials (each
 * kp x ld) -- the reduce step fused into the epilogue (single-GPU path; with cell shards the reduced
 * products must exist in the all-reduce buffer first, so `num` is used).
 * gram_out != NULL: additionally out = F_new^T F_new (kp x kp), the Gram the NEXT half step needs (K4
 * fused); gram_partials is scratch of scrna_nmf_update_blocks(rows) * kp*kp doubles. */
int scrna_nmf_update(int solver, double* F64, int64_t ld, int64_t rows, int k, int kp, const double* num,
                     const float* num_parts, int nsplit, const double* gram, double l1, double l2,
                     const int32_t* perms, int perm_stride, int perm_offset, float* S32, float* T32, float* TC32,
                     void* Fh, float* inv_scale, int kq, int parts, double* viol_part, double* gram_partials,
                     double* gram_out, scrna_nmf_ctrl_t* ctrl, void* stream);
/* Build the fp32 shadows of a component-major fp64 master (initial factors). */
int scrna_nmf_shadows(const double* F64, int64_t ld, int64_t rows, int k, int kp, float* S32, float* T32,
                      void* stream);
/* dst[c][r] = src[r][c] (fp32): the resident transposed copy X^T (cells x genes) that lets K1 run as
 * the same output-owning stream as K3 (SURVEY.md §7.3 "keep X^T resident"). */
int scrna_transpose_f32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* dst,
                        int64_t ld_dst, void* stream);

/* End of one CD iteration: violation = sum(viol_a[0..na)) + sum(viol_b[0..nb)) in fixed order;
 * first iteration stores viol_init; done if viol_init == 0 or violation / viol_init <= tol
 * (_nmf.py:504-516); ++iter.  With tol < 0 only the counter advances (MU). */
int scrna_nmf_iter_end(scrna_nmf_ctrl_t* ctrl, const double* viol_a, int na, const double* viol_b,
                       int nb, double tol, int max_iter, void* stream);

/* Zero a control block. */
int scrna_nmf_ctrl_reset(scrna_nmf_ctrl_t* ctrl, void* stream);

/* fp64 master -> fp32 shadow (same layout, `count` elements). */
int scrna_cast_f64_f32(const double* src, float* dst, int64_t count, void* stream);
/* fp64 host-layout matrix (rows x cols, ld_src) -> fp32 (ld_dst >= cols, padding zeroed). */
int scrna_convert_pad_f64_f32(const double* src, int64_t ld_src, float* dst, int64_t ld_dst,
                              int64_t rows, int64_t cols, void* stream);

/* ---- pre-processing filters (cell_filter / gene_filter, sc3_clustering_impl.py:32-58) over a band of the fp64
 * host-layout matrix: col_ge[c] += #{r: X[r][c] >= thr} (accumulated over bands: zero it first), row_ge[r] =
 * #{c: X[r][c] >= thr}, row_gt0[r] = #{c: X[r][c] > 0}; NaN counts as 0 and raises nan_flag[0]. */
int scrna_filter_counts_f64(const double* X, int64_t ld, int64_t rows, int64_t cols, double thr, int32_t* col_ge,
                            int32_t* row_ge, int32_t* row_gt0, int32_t* nan_flag, void* stream);

/* ---- residuals (K7): sum_{i,j} |X - G.C|^p over the whole shard, p = 1 (utils.py:202, target
 * transfer error) or p = 2 (squared Frobenius, _nmf.py:866-879 convergence test of solver='mu').
 * partials: scrna_residual_blocks(ncols, kp) doubles; out: 1 double (fixed-order sum). */
int scrna_residual_blocks(int64_t ncols, int kp);
int scrna_residual(const float* X, int64_t ldx, int g, int64_t ncols, const float* G32, const float* C32,
                   int64_t ldc, int kp, int power, double* partials, double* out, void* stream);
/* the same over the panel-major fp16 storage of X (scrna_narrow_f32_f16) */
int scrna_residual_h(const void* X16, int64_t rows_pad, int g, int64_t ncols, const float* G32, const float* C32,
                     int64_t ldc, int kp, int power, double* partials, double* out, void* stream);

/* out[0] = scale * sum(v[0..n)) in a fixed order (block partial sums -> one scalar; used for the
 * sharded part of the CD violation and for the error scalars before their all-reduce). */
int scrna_sum_f64(const double* v, int n, double* out, double scale, void* stream);

/* ---- target transfer (utils.get_transferred_data_matrix, utils.py:187-230; DaNmfClustering
 *      .get_mixed_data / .calc_rejection, nmf_clustering.py:147-205, 225-272) ------------------ */

/* K6: one multiplicative step with W fixed: H[t][j] *= WtX[t][j] / sum_q WtW[t][q] * H[q][j]
 * (utils.py:201 with (W^T W).H in place of W^T.(W.H)); an exactly-zero denominator sets
 * SCRNA_FLAG_ZERO_DEN in ctrl->flags (utils.py:199-200 raises 'DA target: division by zero.'). */
int scrna_transfer_mu_step(double* H64, float* H32, const double* WtX, const double* WtW, int k, int kp,
                           int64_t ncols, int64_t ldc, scrna_nmf_ctrl_t* ctrl, void* stream);

/* One iteration of the transfer loop (utils.py:195-206) with the stop rule on the device.  scrna_transfer_iteration =
 * K6 + the L1 residual of the updated H (X: fp32 row-major with ldx, or the fp16 panel layout with ldx = rows_pad when
 * x_is_f16; n cells, ncols = n rounded up to the shard's zero padding); the caller all-reduces `out` across cell
 * shards if there are any, then scrna_transfer_stop turns it into err = out / denom (denom = genes x all target cells), applies |err_prev - err| / err_prev <= rel_err && err_prev >= err
 * (or ZERO_DEN, or max_iter) and sets ctrl->done / n_iter / err_last.  All kernels of later enqueued iterations return
 * at once when ctrl->done is set: the host enqueues a batch and reads the control block once per batch. */
int scrna_transfer_iteration(const void* X, int x_is_f16, int64_t ldx, int g, int64_t n, int64_t ncols, const float* G32,
                             double* H64, float* H32, const double* WtX, const double* WtW, int k, int kp, int64_t ldc,
                             double* partials, double* out, scrna_nmf_ctrl_t* ctrl, void* stream);
int scrna_transfer_stop(const double* res, double denom, double rel_err, int max_iter, scrna_nmf_ctrl_t* ctrl,
                        void* stream);

/* labels[j] = argmax_t H64[t][j], ties -> lowest t (np.argmax; utils.py:210, nmf_clustering.py:170). */
int scrna_argmax_cols(const double* H64, int k, int64_t ncols, int64_t ldc, int32_t* labels, void* stream);

/* K8: mixed = mix * W[:, labels] + (1 - mix) * X  and  new_trg = W[:, labels]  (W.H2 with one-hot
 * H2; nmf_clustering.py:185,200).  Any of mixed64 / mixed32 / newtrg64 may be NULL.  mixed32 is the
 * device-resident fp32 copy (leading dimension ld32, padding zeroed) the SC3 stage consumes. */
int scrna_mix_gather(const float* X, int64_t ldx, int g, int64_t ncols, const double* G64, int kp,
                     const int32_t* labels, double mix, double* mixed64, int64_t ld64, float* mixed32,
                     int64_t ld32, double* newtrg64, void* stream);
/* The same with the dense H instead of the one-hot H2 (get_mixed_data(use_H2=False), nmf_clustering.py:180-184):
 * mixed64 = mix . (W.H) + (1 - mix) . X and / or newtrg64 = W.H, both g x ld64 fp64; G64: g x kp row-major,
 * H64: kp x ldc component-major. */
int scrna_mix_dense(const float* X, int64_t ldx, int g, int64_t ncols, const double* G64, int k, int kp,
                    const double* H64, int64_t ldc, double mix, double* mixed64, int64_t ld64, double* newtrg64,
                    void* stream);

/* Per-cell rejection statistics in one sweep (nmf_clustering.py:236-271): out is 5 x ldo fp64:
 * column sums of X, L1 and squared-L2 residuals against W.H, then against W.H2. */
int scrna_reject_stats(const float* X, int64_t ldx, int g, int64_t ncols, const double* G64, int k, int kp,
                       const double* H64, int64_t ldc, const int32_t* labels, double* out, int64_t ldo,
                       void* stream);

/* ---- SC3 (scRNA/sc3_clustering_impl.py:106-260, scRNA/sc3_clustering.py:58-109); all fp64 ------------ */

/* K9 / K19 pairwise over the COLUMNS of X (F x N row-major, leading dimension ld): out is N x N.
 * op 0: Euclidean, sqrt(sum_f (X[f][i]-X[f][j])^2) summed in f order without FMA == scipy pdist
 *       (sc3_clustering_impl.py:129-130; also pdist(consensus) at :252 since the consensus is symmetric);
 * op 1: dot product Gram sum_f X[f][i]*X[f][j] (np.corrcoef at :124, spearmanr at :126). */
int scrna_pairwise_f64(const double* X, int64_t ld, int F, int N, int op, double* out, int64_t ldo, void* stream);
/* The row block [row_begin, row_end) x N of the same matrix, out[(i - row_begin)][j]: SC3 distances sharded by
 * row blocks over the GPUs (each rank its rows, then an all-gather); bit-identical to those rows of the full call. */
int scrna_pairwise_rows_f64(const double* X, int64_t ld, int F, int N, int op, int row_begin, int row_end,
                            double* out, int64_t ldo, void* stream);
/* X[:, j] -= mean_f X[f][j] in place; mean (N doubles) is returned (np.cov centring). */
int scrna_col_center_f64(double* X, int64_t ld, int F, int N, double* mean, void* stream);
/* out = 1 - corrcoef from the Gram G of centred columns, numpy op order: G/(F-1), /std_i, /std_j, clip. */
int scrna_corr_epilogue_f64(const double* G, int64_t ldg, int N, int F, double* out, int64_t ldo, void* stream);
/* K10 average ranks (scipy.stats.rankdata 'average') of every ROW of Xt (N cells x F genes). */
int scrna_rank_avg_f64(const double* Xt, int64_t ld, int N, int F, double* R, int64_t ldr, void* stream);
int scrna_transpose_f64(const double* src, int64_t ld_src, int64_t rows, int64_t cols, double* dst,
                        int64_t ld_dst, void* stream);

/* K11 pca: out = (dm - colmean) with +-inf -> nan, divided by the column nanstd unless all are zero
 * (sc3_clustering_impl.py:161-173).  sd: n doubles out; scratch: 1 int. */
int scrna_pca_prepare_f64(const double* dm, int64_t ld, int n, double* out, int64_t ldo, double* sd,
                          int32_t* scratch, void* stream);
/* K12 the reference's "spectral" matrix (sc3_clustering_impl.py:143-150); work: 2n+1 doubles. */
int scrna_laplacian_f64(const double* dm, int64_t ld, int n, double* out, int64_t ldo, double* work, void* stream);
/* K13 one-sided Jacobi SVD of M (n x n): right singular vectors, replaces scipy.linalg.svd
 * (sc3_clustering_impl.py:178).  init: AT = M^T, VT = I.  sweep: all column pairs once (round-robin),
 * `rotated` += rotations applied (0 after a sweep == converged).  Singular values^2 = row norms of AT. */
int scrna_jacobi_init_f64(const double* M, int64_t ldm, int n, double* AT, double* VT, int64_t ld, void* stream);
int scrna_jacobi_sweep_f64(double* AT, double* VT, int64_t ld, int n, double tol, int32_t* rotated, void* stream);
int scrna_row_sqnorm_f64(const double* AT, int64_t ld, int n, double* out, void* stream);
/* out[i][r] = VT[order[r]][i] : the first c right singular vectors as columns. */
int scrna_gather_vectors_f64(const double* VT, int64_t ld, int n, const int32_t* order, int c, double* out,
                             int64_t ldo, void* stream);

/* K13 at scale: block one-sided Jacobi on N = M^T (sc3_svd_block.cu).  W (n_pad x ld, n_pad = 32 *
 * scrna_bjacobi_blocks(n), rows >= n zero) starts as a copy of M; one call = one sweep (all block pairs once:
 * pair Gram -> 64 x 64 eigenproblem in shared memory -> rotation of the 64 rows, two GEMM-shaped kernels);
 * `rotated` += block pairs that still had a coupling above 1e-8 (0 after a sweep == converged: the rotations
 * of that sweep were applied and, by quadratic convergence, leave couplings at rounding level).  Afterwards
 * row j of W is sigma_j * v_j: scrna_row_sqnorm_f64 gives sigma^2, scrna_gather_scaled_f64 the unit vectors.
 * Columns whose squared norm is <= null_sq (numerically null: pass (1e-12 sigma_max)^2) are never rotated.
 * work: scrna_bjacobi_work_elems(n) doubles.  n > 32. */
int scrna_bjacobi_blocks(int n);
int scrna_bjacobi_splits(int n);
int64_t scrna_bjacobi_work_elems(int n);
int scrna_bjacobi_sweep_f64(double* W, int64_t ld, int n, double tol, double null_sq, double* work,
                            int32_t* rotated, void* stream);
/* out[i][r] = W[order[r]][i] * scale[order[r]], r < c. */
int scrna_gather_scaled_f64(const double* W, int64_t ld, int n, const int32_t* order, const double* scale, int c,
                            double* out, int64_t ldo, void* stream);

/* out[r][i] = W[order[r]][i] * scale[r], r < c (c x n): with scale_r = sqrt(s_r) / sigma_r the Gram of these rows
 * (scrna_pairwise_f64 op 1) is V.diag(s).V^T, the matrix sc3_clustering_impl.py:196-203 rebuilds from the kept
 * components. */
int scrna_gather_rows_scaled_f64(const double* W, int64_t ld, int n, const int32_t* order, const double* scale, int c,
                                 double* out, int64_t ldo, void* stream);

/* K14-K17 k-means as sklearn KMeans(init='k-means++', algorithm='lloyd') runs it
 * (sc3_clustering_impl.py:215-217).  state: 8 doubles {tol, potential, shift_tot, inertia, changed, n_empty, done,
 * iterations}.  scrna_km_prepare: X = V[:, :d] - column means, xsq = squared row norms, state[0] = tol
 * (1e-4 x mean column variance); work: scrna_km_prepare_work_elems(d) doubles. */
int64_t scrna_km_prepare_work_elems(int d);
int scrna_km_prepare(const double* V, int64_t ldv, int n, int d, double* X, int64_t ldx, double* xsq,
                     double* work, double* state, void* stream);
int scrna_kmeans_pp(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, double u0,
                    const double* uniforms, int ntrials, double* closest, double* cand_dist, double* cums,
                    double* pots, int32_t* cand, int32_t* idx, double* state, double* C, int64_t ldc,
                    void* stream);
/* Same with the first centre chosen by the caller: first = searchsorted(cdf, u0, side='right') for
 * cdf = cumsum(full(n, 1/n)) / cumsum[-1] (numpy RandomState.choice) -- a sequential n-term sum the host evaluates once
 * per n instead of one device thread walking it for every restart. */
int scrna_kmeans_pp_from(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, int first,
                         const double* uniforms, int ntrials, double* closest, double* cand_dist, double* cums,
                         double* pots, int32_t* cand, int32_t* idx, double* state, double* C, int64_t ldc,
                         void* stream);
/* k-means++ of all n_init restarts of a fit in one launch sequence (the restart is a grid dimension of every kernel):
 * firsts: n_init first centres (device int32), uniforms: n_init * (1 + (k-1)*ntrials) doubles in the reference's draw
 * order (the first of each restart is the one the host turned into `firsts`), closest / cums: n_init x n, cand_dist:
 * n_init x ntrials x n, pots / cand: n_init x 8, idx: n_init x k, state: n_init x 8, C: n_init x k x ldc. */
int scrna_kmeans_pp_batched(const double* X, int64_t ldx, int n, int d, const double* xsq, int k, int n_init,
                            const int32_t* firsts, const double* uniforms, int ntrials, double* closest,
                            double* cand_dist, double* cums, double* pots, int32_t* cand, int32_t* idx, double* state,
                            double* C, int64_t ldc, void* stream);
int scrna_lloyd_assign(const double* X, int64_t ldx, int n, int d, const double* C, int64_t ldc, int k,
                       int32_t* labels, void* stream);
int64_t scrna_lloyd_sums_work_elems(int k, int d);
int scrna_lloyd_sums(const double* X, int64_t ldx, int n, int d, const int32_t* labels, int k, double* sums,
                     int64_t lds, int32_t* counts, double* work, void* stream);
int scrna_lloyd_finish(double* Cnew, const double* Cold, int64_t ldc, int k, int d, const int32_t* counts,
                       const int32_t* labels, int32_t* labels_old, int n, double* state, void* stream);
/* One gated Lloyd iteration: assign -> sums -> relocate -> finish, with sklearn's stop rule (_kmeans.py:700-722) on the
 * device: state[6] = 1 (labels unchanged) / 2 (centre shift <= tol) / 3 (max_iter), state[7] = iterations done.  Once
 * set, every kernel of a later enqueued iteration returns at once: the host enqueues batches and polls.  The host
 * swaps Ccur / Cnxt every call; the final centres are the Cnxt of iteration state[7].  Zero state[6..7] per run. */
int scrna_lloyd_iteration(const double* X, int64_t ldx, int n, int d, const double* Ccur, double* Cnxt, int64_t ldc,
                          int k, int32_t* labels, int32_t* labels_old, int32_t* counts, double* celld, double* work,
                          double* state, int max_iter, void* stream);
/* All n_init restarts of a fit in one launch set per Lloyd iteration: the assignment is one fp64 GEMM
 * X (n x d) . C^T (d x n_init*k) on the FP64 tensor cores (DMMA) with the argmin as its epilogue, the centre sums
 * read X once for all restarts; a restart whose stop rule has fired is skipped tile by tile.  Ccur / Cnxt:
 * n_init x k x ldc (swapped by the host every call), labels / labels_old: n_init x n, counts: n_init x k, celld:
 * n_init x n, state: n_init x 8 (per restart: [0] = tol, [6..7] zeroed before the first call), work:
 * scrna_lloyd_batched_work_elems doubles.  ldx, ldc even; X, C 16-byte aligned.
 * scrna_lloyd_batched_supported: n_init <= 16, k <= 48, n_init * k <= 400. */
int scrna_lloyd_batched_supported(int k, int n_init);
int64_t scrna_lloyd_batched_work_elems(int k, int d, int n_init);
int scrna_lloyd_iteration_batched(const double* X, int64_t ldx, int n, int d, const double* Ccur, double* Cnxt,
                                  int64_t ldc, int k, int n_init, int32_t* labels, int32_t* labels_old,
                                  int32_t* counts, double* celld, double* work, double* state, int max_iter,
                                  void* stream);
int scrna_km_cell_dist(const double* X, int64_t ldx, int n, int d, const double* C, int64_t ldc,
                       const int32_t* labels, double* out, void* stream);
int scrna_km_inertia(const double* X, int64_t ldx, int n, int d, const double* C, int64_t ldc,
                     const int32_t* labels, double* celld, double* state, void* stream);
int scrna_km_relocate(const double* X, int64_t ldx, int n, int d, const double* Cold, int64_t ldc, double* sums,
                      int64_t lds, int32_t* counts, const int32_t* labels, double* celld, int k, void* stream);

/* K18 co-association counts (sc3_clustering_impl.py:222-241) and consensus = M / cnt (sc3_clustering.py:105). */
int scrna_coassoc_accum(const int32_t* labels, int n, int32_t* M, int64_t ld, void* stream);
int scrna_consensus_scale(const int32_t* M, int64_t ldm, int n, double cnt, double* C, int64_t ldc, void* stream);
/* K20 + K21 complete linkage (NN-chain, SciPy tie-breaking) and cut_tree(n_clusters)
 * (sc3_clustering_impl.py:254-256).  D (n x n) is destroyed; Z: (n-1) x 4 SciPy linkage matrix out;
 * merges: (n-1) x 3 scratch; iwork: 8n + 4(2n-1) ints. */
int scrna_hclust_complete(double* D, int64_t ld, int n, int n_clusters, int32_t* labels, double* Z, double* merges,
                          int32_t* iwork, void* stream);

/* ---- evaluation: kernel-target alignment (scRNA/utils.py:66-152; cmd_target.py:254-260 picks the mixture by it) --
 * After Kx = X^T X (scrna_pairwise_f64 op 1): rowstats gives the row sums and the diagonal; align returns
 * out = {<Kxn, Kyn>, <Kxn, Kxn>, <Kyn, Kyn>} for Kn_ij = (K_ij - mean_i - mean_j + tot) * scale_i * scale_j with the
 * label kernel Ky_ij = [l_i == l_j] formed on the fly (centring / normalisation are closed forms of the row means,
 * the reference uses dense n^3 products).  partials: 3 * scrna_kta_blocks() doubles. */
int scrna_kta_blocks(void);
int scrna_kta_rowstats_f64(const double* K, int64_t ld, int n, double* rowsum, double* diag, void* stream);
/* Gaussian kernel from the linear Gram, in place: K_ij <- exp(-(diag_i - 2 K_ij + diag_j) / param), diag = the Gram's
 * diagonal copied beforehand (get_kernel(type='rbf'), scRNA/utils.py:108-124; unsupervised_acc_kta(kernel='rbf')). */
int scrna_kta_rbf_f64(double* K, int64_t ld, int n, const double* diag, double param, void* stream);
int scrna_kta_align_f64(const double* K, int64_t ld, int n, const int32_t* labels, const double* mean_x,
                        const double* scale_x, double tot_x, const double* mean_y, const double* scale_y, double tot_y,
                        double* partials, double* out, void* stream);

/* ---- K13 at the BASELINE size: leading eigenvectors of the Gram A = M^T M (sc3_eigh.cu) -- the right singular vectors
 * scipy.linalg.svd yields at sc3_clustering_impl.py:178 -- by Householder tridiagonalisation + bisection + inverse
 * iteration (LAPACK dsytrd / dstebz / dstein / dormtr restated for the device).
 * scrna_sytrd_f64: A (n x n, FULL symmetric, row-major) -> d (n), e (n-1 used), tau (n); on exit row c of A holds the
 *   reflector of column c (entries (c, n), leading 1 stored); work: scrna_sytrd_work_elems(n) doubles.
 * scrna_stebz_top_f64: lam[j] = j-th LARGEST eigenvalue of the tridiagonal (d, e), j < c; [gl, gu] Gershgorin bounds.
 * scrna_stein_f64: X (n x c, row-major [row][vector]) unit eigenvectors for lam; work 4*n*c doubles, iwork n*c bytes.
 * scrna_ormtr_f64: ZT (c x n, one vector per row) <- Q z for every row, Q from scrna_sytrd_f64's reflectors. */
int64_t scrna_sytrd_work_elems(int n);
int scrna_sytrd_f64(double* A, int64_t ld, int n, double* d, double* e, double* tau, double* work, void* stream);
int scrna_stebz_top_f64(const double* d, const double* e, int n, int c, double gl, double gu, double pivmin,
                        double* lam, void* stream);
int scrna_stein_f64(const double* d, const double* e, int n, int c, const double* lam, double tnorm, double* work,
                    int8_t* iwork, double* X, int iters, void* stream);
int scrna_ormtr_f64(const double* A, int64_t ld, int n, const double* tau, double* ZT, int c, void* stream);
/* The same in blocks of 32 reflectors (compact WY: two FP64 tensor-core GEMMs per block instead of 32 dot / axpy
 * passes); work: scrna_ormtr_work_elems(n, c) doubles; needs n even and ZT 16-byte aligned, else
 * SCRNA_ERR_UNSUPPORTED (use scrna_ormtr_f64). */
int64_t scrna_ormtr_work_elems(int n, int c);
int scrna_ormtr_blocked_f64(const double* A, int64_t ld, int n, const double* tau, double* ZT, int c, double* work,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCRNA_B200_H */
