"""CPU stand-in for scrna_b200.kernels.CudaKernels -- TEST INFRASTRUCTURE ONLY.

It lets the multi-process host logic of NmfSolver (cell sharding, all-reduce payload packing, the
stop rule agreeing across ranks, `first='G'|'C'` sequencing) run under the gloo backend on a box
without GPUs.  The arithmetic is numpy on the torch CPU tensors' memory and follows the oracle
(oracle/nmf_oracle.py); the product never imports this file.
"""
import numpy as np
import torch

from oracle import nmf_oracle as no


def _np(t):
    return None if t is None else t.numpy()


class FakeKernels:
    name = "fake-cpu"

    def __init__(self):
        self.torch = torch
        self.device = torch.device("cpu")
        self.launches = 0
        self.sweep_events = None

    # sizing
    def xht_splits(self, g, ncols, kp): return 2
    def wtx_splits(self, g, ncols, kp): return 3
    def gram_blocks(self): return 4
    def update_blocks(self, rows): return (rows + 127) // 128
    def residual_blocks(self, ncols, kp): return 8
    def tc_supported(self, kp): return False
    def h_supported(self, kq): return False

    @staticmethod
    def _done(ctrl):
        return ctrl is not None and int(ctrl[1]) != 0

    def convert_pad(self, src, ld_src, dst, ld_dst, rows, cols):
        d = _np(dst); d[:] = 0
        d[:rows, :cols] = _np(src)[:rows, :cols].astype(np.float32)

    def transpose(self, src, ld_src, rows, cols, dst, ld_dst):
        _np(dst)[:cols, :rows] = _np(src)[:rows, :cols].T

    def shadows(self, F64, ld, rows, k, kp, S32, T32):
        F = _np(F64)
        if S32 is not None:
            s = _np(S32); s[:] = 0; s[:rows, :k] = F[:k, :rows].T.astype(np.float32)
        if T32 is not None:
            _np(T32)[:k, :rows] = F[:k, :rows].astype(np.float32)

    def wtx_partial(self, X, ldx, g, ncols, G32, kp, part, ldc, nsplit, ctrl, variant=0):
        if self._done(ctrl): return
        Xn, Gn, P = _np(X), _np(G32), _np(part)
        rps = (g + nsplit - 1) // nsplit
        P[:] = 0
        for s in range(nsplit):
            r0, r1 = s * rps, min(g, (s + 1) * rps)
            if r0 < r1:
                P[s, :, :ncols] = (Gn[r0:r1].astype(np.float64).T @ Xn[r0:r1, :ncols].astype(np.float64)).astype(np.float32)

    def reduce_cols(self, part, nsplit, kp, ldc, ncols, out, ctrl):
        if self._done(ctrl): return
        o = _np(out).reshape(kp, ldc)
        o[:] = _np(part).astype(np.float64).sum(axis=0)

    def gram(self, F64, rs, cs, rows, k, kp, partials, out, ctrl):
        if self._done(ctrl): return
        assert rs == 1
        F = _np(F64)[:k, :rows]
        o = _np(out).reshape(kp, kp); o[:] = 0
        o[:k, :k] = F @ F.T

    def update(self, solver, F64, ld, rows, k, kp, num, gram, l1, l2, perms, perm_stride, perm_offset, S32, T32,
               viol_part, ctrl, num_parts=None, nsplit=0, gram_partials=None, gram_out=None, TC32=None, Fh=None,
               inv_scale=None, kq=0, parts=3):
        if self._done(ctrl): return
        F = _np(F64)
        W = np.ascontiguousarray(F[:k, :rows].T)
        if num_parts is not None:
            N = _np(num_parts).astype(np.float64).sum(axis=0).reshape(kp, ld)[:k, :rows].T.copy()
        else:
            N = _np(num).reshape(kp, ld)[:k, :rows].T.copy()
        G = _np(gram).reshape(kp, kp)[:k, :k].copy()
        if solver == 1:
            it = int(ctrl[0]) if ctrl is not None else 0
            perm = _np(perms)[it * perm_stride + perm_offset, :k]
            G.flat[:: k + 1] += l2
            v = no.cd_sweep(W, G, N - l1, perm)
            vp = _np(viol_part); vp[:] = 0; vp[0] = v
        else:
            den = W @ G
            if l1 > 0: den += l1
            if l2 > 0: den = den + l2 * W
            den[den == 0] = no.EPSILON
            W = W * (N / den)
            if viol_part is not None: _np(viol_part)[:] = 0
        F[:k, :rows] = W.T
        self.shadows(F64, ld, rows, k, kp, S32, T32)
        if gram_out is not None:
            o = _np(gram_out).reshape(kp, kp); o[:] = 0
            o[:k, :k] = W.T @ W

    def iter_end(self, ctrl, viol_a, na, viol_b, nb, tol, max_iter):
        c = _np(ctrl)
        if c[1]: return
        d = c[4:].view(np.float64)
        v = (float(_np(viol_a)[:na].sum()) if viol_a is not None else 0.0) + \
            (float(_np(viol_b)[:nb].sum()) if viol_b is not None else 0.0)
        c[0] += 1
        it = int(c[0])
        if tol >= 0:
            d[1] = v
            if it == 1: d[0] = v
            if d[0] == 0.0 or v / d[0] <= tol:
                c[1] = 1; c[2] = it
        if not c[1] and it >= max_iter:
            c[1] = 1; c[2] = it

    def ctrl_reset(self, ctrl):
        _np(ctrl)[:] = 0

    def sum_f64(self, v, n, out, scale=1.0):
        _np(out)[0] = float(_np(v)[:n].sum()) * scale

    def residual(self, X, ldx, g, ncols, G32, C32, ldc, kp, power, partials, out):
        R = _np(X)[:, :ncols].astype(np.float64) - _np(G32).astype(np.float64) @ _np(C32)[:, :ncols].astype(np.float64)
        _np(out)[0] = np.abs(R).sum() if power == 1 else (R * R).sum()
