"""Typed wrappers over the C ABI (include/scrna_b200.h) taking torch CUDA tensors.

`CudaKernels` is the ONLY compute provider of the package.  The solvers talk to it through this
small method set so that the host-side logic (sharding, all-reduce packing, stop rules) can be
exercised in tests with a stand-in provider; the product never constructs anything else.
"""
import os

from . import _lib


class CudaKernels:
    name = "cuda"

    def __init__(self):
        self.torch = _lib.require_cuda()
        _lib.load()
        self.launches = 0  # number of kernels launched through the C ABI (bench.py: gpu_launches)
        self.sweep_events = None  # bench.py sets a list: (start, stop, X bytes) CUDA events per sweep launch
        self.h_flags = int(os.environ.get("SCRNA_B200_H_FLAGS", "0"))  # bring-up probes of the 16-bit sweep

    def _s(self):
        return self.torch.cuda.current_stream().cuda_stream

    def _call(self, name, *args):
        self.launches += 1
        return _lib.call(name, *args)

    # ---- sizing -----------------------------------------------------------------------------
    def xht_splits(self, g, ncols, kp):
        return _lib.call("scrna_nmf_xht_splits", g, ncols, kp)

    def wtx_splits(self, g, ncols, kp):
        return _lib.call("scrna_nmf_wtx_splits", g, ncols, kp)

    def gram_blocks(self):
        return _lib.call("scrna_nmf_gram_blocks")

    def update_blocks(self, rows):
        return _lib.call("scrna_nmf_update_blocks", rows)

    def residual_blocks(self, ncols, kp):
        return _lib.call("scrna_residual_blocks", ncols, kp)

    # ---- NMF iteration ------------------------------------------------------------------------
    def xht_partial(self, X, ldx, g, ncols, C32, ldc, kp, part, nsplit, ctrl, variant=0):
        self._call("scrna_nmf_xht_partial", X, ldx, g, ncols, C32, ldc, kp, part, nsplit, variant, ctrl, self._s())

    def wtx_partial(self, X, ldx, g, ncols, G32, kp, part, ldc, nsplit, ctrl, variant=2):
        ev = self.sweep_events
        if ev is not None:
            a = self.torch.cuda.Event(enable_timing=True)
            b = self.torch.cuda.Event(enable_timing=True)
            a.record()
        self._call("scrna_nmf_wtx_partial", X, ldx, g, ncols, G32, kp, part, ldc, nsplit, variant, ctrl, self._s())
        if ev is not None:
            b.record()
            ev.append((a, b, int(g) * int(ncols) * 4))

    def tc_supported(self, kp):
        return bool(_lib.call("scrna_nmf_tc_supported", kp))

    def wtx_tc_splits(self, g, ncols, kp):
        return _lib.call("scrna_nmf_wtx_tc_splits", g, ncols, kp)

    def tc_shadow_elems(self, rows, kp):
        return _lib.call("scrna_nmf_tc_shadow_elems", rows, kp)

    def wtx_partial_tc(self, X, ldx, g, ncols, Fs, kp, part, ldc, nsplit, ctrl, flags=0):
        ev = self.sweep_events
        if ev is not None:
            a = self.torch.cuda.Event(enable_timing=True)
            b = self.torch.cuda.Event(enable_timing=True)
            a.record()
        self._call("scrna_nmf_wtx_partial_tc", X, ldx, g, ncols, Fs, kp, part, ldc, nsplit, flags, ctrl, self._s())
        if ev is not None:
            b.record()
            ev.append((a, b, int(g) * int(ncols) * 4))

    # ---- 16-bit storage path (nmf_sweep_h.cu) ---------------------------------------------------
    def h_supported(self, kq):
        return bool(_lib.call("scrna_nmf_h_supported", kq))

    def h_nb(self, kq, parts):
        return _lib.call("scrna_nmf_h_nb", kq, parts)

    def h_shadow_elems(self, rows, kq, parts):
        return _lib.call("scrna_nmf_h_shadow_elems", rows, kq, parts)

    def wtx_h_splits(self, g, ncols, kq, parts):
        return _lib.call("scrna_nmf_wtx_h_splits", g, ncols, kq, parts)

    def h_groups(self, rows):
        return _lib.call("scrna_nmf_h_groups", rows)

    def gstep_mu(self, peers, nranks, rank, off, g, ldg, k, kq, nb, parts, l1, l2, PA, SA, gpartC, groupsC, nan_flag):
        """Fused gene half step of the MU iteration (nmf_step.cu); `off`: dict of arena byte offsets."""
        self._call("scrna_nmf_gstep_mu", peers, nranks, rank, off["xch"], off["sync"], off["G64"], off["Gs32"], off["Gh"],
                   off["Ginv"], off["gpart"], off["gtg"], g, ldg, k, kq, nb, parts, float(l1), float(l2), PA, SA,
                   gpartC, groupsC, nan_flag, self._s())

    def cstep_mu(self, C64, ldn, n, k, kq, nb, parts, PB, SB, gtg, l1, l2, Cs32, Ct32, Ch, Cinv, gpartC, nan_flag):
        self._call("scrna_nmf_cstep_mu", C64, ldn, n, k, kq, nb, parts, PB, SB, gtg, float(l1), float(l2), Cs32, Ct32,
                   Ch, Cinv, gpartC, nan_flag, self._s())

    def panel_rows(self, rows):
        return _lib.call("scrna_f16_panel_rows", rows)

    def panel_elems(self, rows, cols):
        return _lib.call("scrna_f16_panel_elems", rows, cols)

    def wtx_partial_h(self, X16, rows_pad, g, ncols, Fh, inv_scale, kq, parts, part, ldc, nsplit, ctrl, flags=0):
        ev = self.sweep_events
        if ev is not None:
            a = self.torch.cuda.Event(enable_timing=True)
            b = self.torch.cuda.Event(enable_timing=True)
            a.record()
        self._call("scrna_nmf_wtx_partial_h", X16, rows_pad, g, ncols, Fh, inv_scale, kq, parts, part, ldc, nsplit,
                   flags | self.h_flags, ctrl, self._s())
        if ev is not None:
            b.record()
            ev.append((a, b, int(g) * int(ncols) * 2))

    def h_shadow(self, F64, ld, rows, k, kq, parts, gram_parts, nblk, kp, gram_out, Fh, inv_scale, ctrl=None):
        self._call("scrna_nmf_h_shadow", F64, ld, rows, k, kq, parts, gram_parts, nblk, kp, gram_out, Fh, inv_scale,
                   ctrl, self._s())

    def f16_exact(self, X):
        """True if every element of the fp32 tensor is exactly representable in fp16."""
        flag = self.torch.zeros(1, dtype=self.torch.int32, device=X.device)
        self._call("scrna_f16_inexact", X, X.numel(), flag, self._s())
        return int(flag.item()) == 0

    def narrow_f16(self, src, ld_src, rows, cols, dst):
        self._call("scrna_narrow_f32_f16", src, ld_src, rows, cols, dst, self._s())

    def transpose_f16(self, src, ld_src, rows, cols, dst):
        self._call("scrna_transpose_f32_f16", src, ld_src, rows, cols, dst, self._s())

    def tf32_exact(self, X):
        """True if every element of the fp32 tensor is exactly representable in TF32."""
        flag = self.torch.zeros(1, dtype=self.torch.int32, device=X.device)
        self._call("scrna_tf32_inexact", X, X.numel(), flag, self._s())
        return int(flag.item()) == 0

    def tc_shadow(self, F64, ld, rows, k, kp, Fs):
        self._call("scrna_nmf_tc_shadow", F64, ld, rows, k, kp, Fs, self._s())

    def reduce_rows(self, part, nsplit, g, kp, out, ld, ctrl):
        self._call("scrna_nmf_reduce_rows", part, nsplit, g, kp, out, ld, ctrl, self._s())

    def reduce_cols(self, part, nsplit, kp, ldc, ncols, out, ctrl):
        self._call("scrna_nmf_reduce_cols", part, nsplit, kp, ldc, ncols, out, ctrl, self._s())

    def gram(self, F64, rs, cs, rows, k, kp, partials, out, ctrl):
        self.launches += 1  # two kernels per call
        self._call("scrna_nmf_gram", F64, rs, cs, rows, k, kp, partials, out, ctrl, self._s())

    def update(self, solver, F64, ld, rows, k, kp, num, gram, l1, l2, perms, perm_stride, perm_offset, S32, T32,
               viol_part, ctrl, num_parts=None, nsplit=0, gram_partials=None, gram_out=None, TC32=None, Fh=None,
               inv_scale=None, kq=0, parts=3):
        if gram_out is not None or Fh is not None:
            self.launches += 1
        self._call("scrna_nmf_update", solver, F64, ld, rows, k, kp, num, num_parts, nsplit, gram, float(l1),
                   float(l2), perms, perm_stride, perm_offset, S32, T32, TC32, Fh, inv_scale, int(kq), int(parts), viol_part,
                   gram_partials, gram_out, ctrl, self._s())

    def shadows(self, F64, ld, rows, k, kp, S32, T32):
        self._call("scrna_nmf_shadows", F64, ld, rows, k, kp, S32, T32, self._s())

    def transpose(self, src, ld_src, rows, cols, dst, ld_dst):
        self._call("scrna_transpose_f32", src, ld_src, rows, cols, dst, ld_dst, self._s())

    def iter_end(self, ctrl, viol_a, na, viol_b, nb, tol, max_iter):
        self._call("scrna_nmf_iter_end", ctrl, viol_a, na, viol_b, nb, float(tol), int(max_iter), self._s())

    def ctrl_reset(self, ctrl):
        self._call("scrna_nmf_ctrl_reset", ctrl, self._s())

    def cast_f64_f32(self, src, dst, count):
        self._call("scrna_cast_f64_f32", src, dst, count, self._s())

    def convert_pad(self, src, ld_src, dst, ld_dst, rows, cols):
        self._call("scrna_convert_pad_f64_f32", src, ld_src, dst, ld_dst, rows, cols, self._s())

    def residual(self, X, ldx, g, ncols, G32, C32, ldc, kp, power, partials, out):
        self.launches += 1
        name = "scrna_residual_h" if X.dtype == self.torch.float16 else "scrna_residual"   # ldx = rows_pad for fp16
        self._call(name, X, ldx, g, ncols, G32, C32, ldc, kp, power, partials, out, self._s())

    def sum_f64(self, v, n, out, scale=1.0):
        self._call("scrna_sum_f64", v, n, out, float(scale), self._s())

    # ---- target transfer ------------------------------------------------------------------------
    def transfer_mu_step(self, H64, H32, WtX, WtW, k, kp, ncols, ldc, ctrl):
        self._call("scrna_transfer_mu_step", H64, H32, WtX, WtW, k, kp, ncols, ldc, ctrl, self._s())

    def transfer_iteration(self, X, ldx, g, n, ncols, G32, H64, H32, WtX, WtW, k, kp, ldc, partials, out, ctrl):
        self.launches += 3
        self._call("scrna_transfer_iteration", X, int(X.dtype == self.torch.float16), ldx, g, n, ncols, G32, H64, H32, WtX,
                   WtW, k, kp, ldc, partials, out, ctrl, self._s())

    def transfer_stop(self, out, denom, rel_err, max_iter, ctrl):
        self.launches += 1
        self._call("scrna_transfer_stop", out, float(denom), float(rel_err), int(max_iter), ctrl, self._s())

    def argmax_cols(self, H64, k, ncols, ldc, labels):
        self._call("scrna_argmax_cols", H64, k, ncols, ldc, labels, self._s())

    def mix_gather(self, X, ldx, g, ncols, G64, kp, labels, mix, mixed64, ld64, mixed32, ld32, newtrg64):
        self._call("scrna_mix_gather", X, ldx, g, ncols, G64, kp, labels, float(mix), mixed64, ld64, mixed32, ld32,
                   newtrg64, self._s())

    def mix_dense(self, X, ldx, g, ncols, G64, k, kp, H64, ldc, mix, mixed64, ld64, newtrg64):
        self.launches += 1
        self._call("scrna_mix_dense", X, ldx, g, ncols, G64, k, kp, H64, ldc, float(mix), mixed64, ld64, newtrg64,
                   self._s())

    def reject_stats(self, X, ldx, g, ncols, G64, k, kp, H64, ldc, labels, out, ldo):
        self._call("scrna_reject_stats", X, ldx, g, ncols, G64, k, kp, H64, ldc, labels, out, ldo, self._s())


_default = None


def default_kernels():
    """The process-wide CUDA provider (raises without a GPU or without the built library)."""
    global _default
    if _default is None:
        _default = CudaKernels()
    return _default
