"""ctypes binding of libscrna_b200.so (include/scrna_b200.h).

The library is the product: there is NO fallback.  If the shared object is missing, or a compute
entry point is called without a CUDA device, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libscrna_b200.so")

SCRNA_OK = 0
SOLVER_MU = 0
SOLVER_CD = 1
FLAG_NAN = 1
FLAG_ZERO_DEN = 2
MAX_K = 64

_ERRORS = {-1: "bad argument", -2: "CUDA error", -3: "unsupported configuration"}

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p

# name -> argtypes (restype is always int)
_SIGNATURES = {
    "scrna_abi_version": [],
    "scrna_device_info": [c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_nmf_xht_splits": [c_int, c_i64, c_int],
    "scrna_nmf_wtx_splits": [c_int, c_i64, c_int],
    "scrna_nmf_xht_partial": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_i64, c_int, c_ptr, c_int, c_int, c_ptr, c_ptr],
    "scrna_nmf_wtx_partial": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_int, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr],
    "scrna_nmf_tc_supported": [c_int],
    "scrna_nmf_wtx_tc_splits": [c_int, c_i64, c_int],
    "scrna_nmf_wtx_partial_tc": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_int, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr],
    "scrna_nmf_tc_shadow": [c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr],
    "scrna_nmf_tc_shadow_elems": [c_i64, c_int],
    "scrna_tf32_inexact": [c_ptr, c_i64, c_ptr, c_ptr],
    "scrna_nmf_reduce_rows": [c_ptr, c_int, c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr],
    "scrna_nmf_reduce_cols": [c_ptr, c_int, c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "scrna_nmf_gram_blocks": [],
    "scrna_nmf_gram": [c_ptr, c_i64, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_nmf_update_blocks": [c_i64],
    "scrna_nmf_update": [c_int, c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr, c_dbl, c_dbl,
                         c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr,
                         c_ptr, c_ptr],
    "scrna_nmf_h_supported": [c_int],
    "scrna_nmf_h_nb": [c_int, c_int],
    "scrna_nmf_h_shadow_elems": [c_i64, c_int, c_int],
    "scrna_nmf_wtx_h_splits": [c_int, c_i64, c_int, c_int],
    "scrna_nmf_wtx_partial_h": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_int, c_int, c_ptr, c_i64, c_int, c_int,
                                c_ptr, c_ptr],
    "scrna_nmf_h_shadow": [c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                           c_ptr],
    "scrna_f16_inexact": [c_ptr, c_i64, c_ptr, c_ptr],
    "scrna_nmf_h_groups": [c_i64],
    "scrna_nmf_step_groups": [c_i64],
    "scrna_nmf_gstep_mu": [c_ptr, c_int, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_i64,
                           c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_ptr, c_int, c_ptr, c_int, c_ptr, c_ptr],
    "scrna_nmf_cstep_mu": [c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int, c_ptr, c_int, c_ptr, c_dbl, c_dbl, c_ptr,
                           c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_f16_panel_rows": [c_i64],
    "scrna_f16_panel_elems": [c_i64, c_i64],
    "scrna_narrow_f32_f16": [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr],
    "scrna_transpose_f32_f16": [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr],
    "scrna_residual_h": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr],
    "scrna_nmf_shadows": [c_ptr, c_i64, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr],
    "scrna_transpose_f32": [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr],
    "scrna_nmf_iter_end": [c_ptr, c_ptr, c_int, c_ptr, c_int, c_dbl, c_int, c_ptr],
    "scrna_nmf_ctrl_reset": [c_ptr, c_ptr],
    "scrna_cast_f64_f32": [c_ptr, c_ptr, c_i64, c_ptr],
    "scrna_convert_pad_f64_f32": [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64, c_ptr],
    "scrna_residual_blocks": [c_i64, c_int],
    "scrna_residual": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr],
    "scrna_filter_counts_f64": [c_ptr, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_sum_f64": [c_ptr, c_int, c_ptr, c_dbl, c_ptr],
    "scrna_transfer_mu_step": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_i64, c_i64, c_ptr, c_ptr],
    "scrna_argmax_cols": [c_ptr, c_int, c_i64, c_i64, c_ptr, c_ptr],
    "scrna_mix_gather": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_int, c_ptr, c_dbl, c_ptr, c_i64, c_ptr, c_i64,
                         c_ptr, c_ptr],
    "scrna_mix_dense": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_int, c_int, c_ptr, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_ptr],
    "scrna_reject_stats": [c_ptr, c_i64, c_int, c_i64, c_ptr, c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_i64,
                           c_ptr],
    # ---- SC3 ----
    "scrna_pairwise_f64": [c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_i64, c_ptr],
    "scrna_pairwise_rows_f64": [c_ptr, c_i64, c_int, c_int, c_int, c_int, c_int, c_ptr, c_i64, c_ptr],
    "scrna_col_center_f64": [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr],
    "scrna_corr_epilogue_f64": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr],
    "scrna_rank_avg_f64": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr],
    "scrna_transpose_f64": [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr],
    "scrna_pca_prepare_f64": [c_ptr, c_i64, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr],
    "scrna_laplacian_f64": [c_ptr, c_i64, c_int, c_ptr, c_i64, c_ptr, c_ptr],
    "scrna_jacobi_init_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_i64, c_ptr],
    "scrna_jacobi_sweep_f64": [c_ptr, c_ptr, c_i64, c_int, c_dbl, c_ptr, c_ptr],
    "scrna_row_sqnorm_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr],
    "scrna_gather_vectors_f64": [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_i64, c_ptr],
    "scrna_bjacobi_blocks": [c_int],
    "scrna_bjacobi_splits": [c_int],
    "scrna_bjacobi_work_elems": [c_int],
    "scrna_bjacobi_sweep_f64": [c_ptr, c_i64, c_int, c_dbl, c_dbl, c_ptr, c_ptr, c_ptr],
    "scrna_sytrd_work_elems": [c_int],
    "scrna_sytrd_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_stebz_top_f64": [c_ptr, c_ptr, c_int, c_int, c_dbl, c_dbl, c_dbl, c_ptr, c_ptr],
    "scrna_stein_f64": [c_ptr, c_ptr, c_int, c_int, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_int, c_ptr],
    "scrna_ormtr_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_int, c_ptr],
    "scrna_ormtr_work_elems": [c_int, c_int],
    "scrna_ormtr_blocked_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ptr],
    "scrna_gather_scaled_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr],
    "scrna_gather_rows_scaled_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr],
    "scrna_kta_blocks": [],
    "scrna_kta_rowstats_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr],
    "scrna_kta_rbf_f64": [c_ptr, c_i64, c_int, c_ptr, c_dbl, c_ptr],
    "scrna_kta_align_f64": [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr],
    "scrna_km_prepare": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_kmeans_pp": [c_ptr, c_i64, c_int, c_int, c_ptr, c_int, c_dbl, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                        c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    "scrna_kmeans_pp_batched": [c_ptr, c_i64, c_int, c_int, c_ptr, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr,
                                c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    "scrna_kmeans_pp_from": [c_ptr, c_i64, c_int, c_int, c_ptr, c_int, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                             c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr],
    "scrna_lloyd_assign": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_int, c_ptr, c_ptr],
    "scrna_lloyd_sums": [c_ptr, c_i64, c_int, c_int, c_ptr, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr],
    "scrna_lloyd_sums_work_elems": [c_int, c_int],
    "scrna_lloyd_finish": [c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr],
    "scrna_transfer_iteration": [c_ptr, c_int, c_i64, c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_i64,
                                 c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_transfer_stop": [c_ptr, c_dbl, c_dbl, c_int, c_ptr, c_ptr],
    "scrna_lloyd_iteration": [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_i64, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                              c_ptr, c_int, c_ptr],
    "scrna_lloyd_batched_supported": [c_int, c_int],
    "scrna_lloyd_batched_work_elems": [c_int, c_int, c_int],
    "scrna_lloyd_iteration_batched": [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr,
                                      c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr],
    "scrna_km_prepare_work_elems": [c_int],
    "scrna_km_cell_dist": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr],
    "scrna_km_inertia": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr],
    "scrna_km_relocate": [c_ptr, c_i64, c_int, c_int, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_int, c_ptr],
    "scrna_coassoc_accum": [c_ptr, c_int, c_ptr, c_i64, c_ptr],
    "scrna_consensus_scale": [c_ptr, c_i64, c_int, c_dbl, c_ptr, c_i64, c_ptr],
    "scrna_hclust_complete": [c_ptr, c_i64, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
}

# entry points that return a count / version instead of a status code
_VALUE_RETURNING = {"scrna_abi_version", "scrna_nmf_xht_splits", "scrna_nmf_wtx_splits", "scrna_nmf_gram_blocks",
                    "scrna_nmf_update_blocks", "scrna_residual_blocks", "scrna_nmf_tc_supported", "scrna_nmf_wtx_tc_splits",
                    "scrna_nmf_tc_shadow_elems", "scrna_bjacobi_blocks", "scrna_bjacobi_splits",
                    "scrna_bjacobi_work_elems", "scrna_lloyd_sums_work_elems", "scrna_kta_blocks",
                    "scrna_nmf_h_supported", "scrna_nmf_h_nb", "scrna_nmf_h_shadow_elems", "scrna_nmf_wtx_h_splits",
                    "scrna_f16_panel_rows", "scrna_f16_panel_elems", "scrna_nmf_h_groups", "scrna_nmf_step_groups", "scrna_sytrd_work_elems",
                    "scrna_lloyd_batched_supported", "scrna_lloyd_batched_work_elems", "scrna_km_prepare_work_elems",
                    "scrna_ormtr_work_elems"}


_RETURNS_I64 = {"scrna_sytrd_work_elems", "scrna_nmf_tc_shadow_elems", "scrna_nmf_h_shadow_elems", "scrna_f16_panel_rows", "scrna_f16_panel_elems",
                "scrna_nmf_h_groups", "scrna_bjacobi_work_elems", "scrna_lloyd_sums_work_elems",
                "scrna_lloyd_batched_work_elems", "scrna_km_prepare_work_elems", "scrna_ormtr_work_elems"}


class ScrnaError(RuntimeError):
    pass


_lib = None


def exported_symbols():
    """Names every include/scrna_b200.h declaration must resolve to (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (idempotent).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScrnaError(
            "libscrna_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C scrna_b200/csrc`; there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = c_i64 if name in _RETURNS_I64 else c_int
    _lib = lib
    return lib


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def call(name, *args):
    """Call a status-returning entry point; tensors are passed as their device pointers."""
    lib = load()
    rc = getattr(lib, name)(*[_ptr(a) if (a is None or hasattr(a, "data_ptr")) else a for a in args])
    if name in _VALUE_RETURNING:
        if rc < 0:
            raise ScrnaError("%s failed: %s" % (name, _ERRORS.get(rc, rc)))
        return rc
    if rc != SCRNA_OK:
        detail = ""
        if rc == -2:
            try:
                import torch
                torch.cuda.synchronize()
            except Exception as e:  # surfaces the sticky CUDA error text
                detail = " (%s)" % e
        raise ScrnaError("%s failed: %s%s" % (name, _ERRORS.get(rc, rc), detail))
    return rc


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise ScrnaError("scrna_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch


def current_stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
