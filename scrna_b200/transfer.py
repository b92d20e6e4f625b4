"""Device-resident target transfer: H-only multiplicative updates with the dictionary W fixed,
the one-hot reconstruction and the mixed data (scRNA/utils.py:187-230, scRNA/nmf_clustering.py
:147-205, 225-272; SURVEY.md §8 rows a7-a9, kernels K5-K8).

Differences from the reference's arithmetic that do not change results beyond fp rounding:
  * W^T.X is computed once (the reference recomputes the constant every iteration, utils.py:201);
  * the denominator is (W^T W).H instead of W^T.(W.H) (same zeros, 1e-15 relative difference);
  * sum|X - W.H| is one streaming pass that never materialises W.H (the reference forms it 3x).
"""
import numpy as np

from . import _lib
from .nmf_solver import DeviceMatrix, pad_k


class TransferSession:
    def __init__(self, W, trg_data, kernels=None, comm=None):
        if kernels is None:
            from .kernels import default_kernels
            kernels = default_kernels()
        self.kn = kernels
        t = self.torch = kernels.torch
        W = np.ascontiguousarray(W, dtype=np.float64)
        self.W_host = W
        self.X = trg_data if isinstance(trg_data, DeviceMatrix) else DeviceMatrix.from_host(trg_data, kernels)
        X = self.X
        if W.shape[0] != X.g:
            raise ValueError("dictionary has %d genes, target data %d" % (W.shape[0], X.g))
        self.k = W.shape[1]
        self.kp = kp = pad_k(self.k)
        self.comm = comm if (comm is not None and comm.world > 1) else None
        dev = X.data.device
        self.ldc = ldc = X.ldx
        f64, f32 = t.float64, t.float32
        self.G64 = t.zeros((X.g, kp), dtype=f64, device=dev)
        self.G64[:, : self.k].copy_(t.from_numpy(W).to(dev))
        self.G32 = t.zeros((X.g, kp), dtype=f32, device=dev)
        kernels.cast_f64_f32(self.G64, self.G32, self.G64.numel())
        self.H64 = t.zeros((kp, ldc), dtype=f64, device=dev)
        self.H32 = t.zeros((kp, ldc), dtype=f32, device=dev)
        self.WtX = t.zeros((kp, ldc), dtype=f64, device=dev)
        self.WtW = t.zeros(kp * kp, dtype=f64, device=dev)
        self.gram_part = t.zeros(kernels.gram_blocks() * kp * kp, dtype=f64, device=dev)
        self.SB = kernels.wtx_splits(X.g, X.ncols, kp)
        self.PB = t.zeros((self.SB, kp, ldc), dtype=f32, device=dev)
        self.res_part = t.zeros(max(1, kernels.residual_blocks(X.ncols, kp)), dtype=f64, device=dev)
        self.res_out = t.zeros(1, dtype=f64, device=dev)
        self.ctrl = t.zeros(16, dtype=t.int32, device=dev)
        self.labels_dev = t.zeros(max(X.n, 1), dtype=t.int32, device=dev)
        self.n_total = X.n
        if self.comm is not None:
            cnt = t.tensor([float(X.n)], dtype=f64, device=dev)
            self.comm.all_reduce(cnt)
            self.n_total = int(cnt.item())
        # constants of the solve
        kernels.wtx_partial(X.data, X.ldx, X.g, X.ncols, self.G32, kp, self.PB, ldc, self.SB, None)
        kernels.reduce_cols(self.PB, self.SB, kp, ldc, X.ncols, self.WtX, None)
        kernels.gram(self.G64, kp, 1, X.g, self.k, kp, self.gram_part, self.WtW, None)

    def _set_H(self, H0):
        t, X = self.torch, self.X
        H0 = np.ascontiguousarray(H0, dtype=np.float64)
        if H0.shape != (self.k, X.n):
            raise ValueError("H init must be %d x %d" % (self.k, X.n))
        self.H64.zero_()
        self.H64[: self.k, : X.n].copy_(t.from_numpy(H0).to(X.data.device))
        self.kn.cast_f64_f32(self.H64, self.H32, self.H64.numel())

    def _flags(self):
        return int(self.ctrl.cpu()[3])

    def l1_error(self):
        """sum|X - W.H| / (g * n_t) over all shards (utils.py:202)."""
        X = self.X
        self.kn.residual(X.data, X.ldx, X.g, X.ncols, self.G32, self.H32, self.ldc, self.kp, 1, self.res_part,
                         self.res_out)
        if self.comm is not None:
            self.comm.all_reduce(self.res_out)
        return float(self.res_out.cpu()[0]) / float(X.g * self.n_total)

    POLL_EVERY = 4   # iterations enqueued between two reads of the control block

    def solve(self, H0, max_iter=100, rel_err=1e-3):
        """The reference loop (utils.py:195-206) from an injected H0.  Returns (n_iter, err).  The stop rule runs on
        the device (scrna_transfer_stop); the host enqueues POLL_EVERY iterations at a time and reads the control
        block once per batch -- iterations enqueued past the stop are no-ops."""
        self._set_H(H0)
        self.kn.ctrl_reset(self.ctrl)
        X = self.X
        denom = float(X.g * self.n_total)
        queued = 0
        raw = None
        while queued < max_iter:
            nb = min(self.POLL_EVERY, max_iter - queued)
            for _ in range(nb):
                self.kn.transfer_iteration(X.data, X.ldx, X.g, X.n, X.ncols, self.G32, self.H64, self.H32, self.WtX, self.WtW,
                                           self.k, self.kp, self.ldc, self.res_part, self.res_out, self.ctrl)
                if self.comm is not None:
                    self.comm.all_reduce(self.res_out)
                self.kn.transfer_stop(self.res_out, denom, rel_err, max_iter, self.ctrl)
            queued += nb
            raw = self.ctrl.cpu().numpy()
            if raw[3] & _lib.FLAG_ZERO_DEN:
                raise Exception('DA target: division by zero.')
            if raw[1]:
                break
        self.n_iter = int(raw[2]) if raw is not None else 0
        self.err = float(raw.view(np.float64)[4]) if raw is not None else None
        return self.n_iter, self.err

    def labels(self):
        X = self.X
        self.kn.argmax_cols(self.H64, self.k, X.n, self.ldc, self.labels_dev)
        return self.labels_dev[: X.n].cpu().numpy().astype(np.int64)

    def set_labels(self, labels):
        """Use the given per-cell labels (a caller-supplied one-hot H2) for `mixed` / `reject_stats`."""
        t = self.torch
        lab = np.ascontiguousarray(labels, dtype=np.int32).reshape(-1)
        if lab.size != self.X.n:
            raise ValueError("one label per target cell expected")
        self.labels_dev[: self.X.n].copy_(t.from_numpy(lab).to(self.labels_dev.device))

    def H_host(self):
        return self.H64[: self.k, : self.X.n].cpu().numpy().copy()

    def mixed(self, mix, want_newtrg=True, want_mixed32=False):
        """(mixed fp64 host, new_trg fp64 host[, mixed fp32 DeviceMatrix]); labels() must have run."""
        t, X = self.torch, self.X
        dev = X.data.device
        mixed64 = t.empty((X.g, X.n), dtype=t.float64, device=dev)
        newtrg = t.empty((X.g, X.n), dtype=t.float64, device=dev) if want_newtrg else None
        m32 = t.empty((X.g, X.ldx), dtype=t.float32, device=dev) if want_mixed32 else None
        self.kn.mix_gather(X.data, X.ldx, X.g, X.n, self.G64, self.kp, self.labels_dev, mix, mixed64, X.n,
                           m32, X.ldx if want_mixed32 else 0, newtrg)
        out = [mixed64.cpu().numpy(), newtrg.cpu().numpy() if want_newtrg else None]
        if want_mixed32:
            out.append(DeviceMatrix(m32, X.n))
        return tuple(out)

    def mixed_dense(self, mix):
        """(mix . W.H + (1 - mix) . X, W.H) as fp64 host arrays: get_mixed_data(use_H2=False)."""
        t, X = self.torch, self.X
        dev = X.data.device
        mixed64 = t.empty((X.g, X.n), dtype=t.float64, device=dev)
        newtrg = t.empty((X.g, X.n), dtype=t.float64, device=dev)
        self.kn.mix_dense(X.data, X.ldx, X.g, X.n, self.G64, self.k, self.kp, self.H64, self.ldc, mix, mixed64, X.n,
                          newtrg)
        return mixed64.cpu().numpy(), newtrg.cpu().numpy()

    def reject_stats(self):
        """5 x n_t: column sums, L1 / squared-L2 residuals vs W.H, then vs W.H2."""
        t, X = self.torch, self.X
        out = t.zeros((5, max(X.n, 1)), dtype=t.float64, device=X.data.device)
        self.kn.reject_stats(X.data, X.ldx, X.g, X.n, self.G64, self.k, self.kp, self.H64, self.ldc,
                             self.labels_dev, out, max(X.n, 1))
        return out.cpu().numpy()
