"""Device-side SC3 stages (fp64) sequenced from the host: distances, transformations, k-means ensemble,
consensus matrix and complete-linkage consensus clustering (kernels K9-K21).

Every array operation is a C-ABI kernel (include/scrna_b200.h); the host sequences launches, reads a
few scalars per k-means iteration / Jacobi sweep for the stop rules, and draws nothing: random inputs
(k-means++ uniforms) are passed in, pre-drawn from the reference's RNG stream.
"""
import contextlib
import time

import os

import numpy as np

from . import _lib


def kmeans_draw_count(k):
    """Uniform doubles one k-means++ init consumes from the legacy RandomState: choice(n, p) = 1, then
    (k-1) x uniform(size=2+int(log k))  (SP/sklearn/cluster/_kmeans.py:231,249)."""
    trials = 2 + int(np.log(k))
    return 1 + (k - 1) * trials, trials


def uniform_choice_cdf(n):
    """cdf of RandomState.choice(n, p=full(n, 1/n)): numpy accumulates p sequentially and normalises by the last entry
    (numpy/random/mtrand.pyx, `choice`); sklearn's k-means++ draws its first centre this way (_kmeans.py:231)."""
    cdf = np.full(n, 1.0 / n).cumsum()
    cdf /= cdf[-1]
    return cdf


def first_centres(cdf, uniforms):
    """choice's index for each uniform draw: searchsorted(cdf, u, side='right'), clipped like the device kernel."""
    return np.minimum(np.searchsorted(cdf, uniforms, side="right"), cdf.size - 1)


def same_clustering(l1, l2, k):
    """sklearn _is_same_clustering (_k_means_common.pyx:314-328)."""
    mapping = np.full(k, -1, dtype=np.int64)
    for a, b in zip(l1.tolist(), l2.tolist()):
        if mapping[a] == -1:
            mapping[a] = b
        elif mapping[a] != b:
            return False
    return True


class Sc3Ops:
    def __init__(self, kernels=None):
        if kernels is None:
            from .kernels import default_kernels
            kernels = default_kernels()
        self.kn = kernels
        self.torch = kernels.torch
        self.dev = self.torch.device("cuda", self.torch.cuda.current_device())
        self._uniform_cdf = {}
        self.stage_seconds = None   # bench.py sets a dict: synchronised wall seconds accumulated per stage
        self.svd_flops = 0.0
        self.jacobi_sweeps = 0
        self._last = None

    @contextlib.contextmanager
    def _stage(self, name):
        ss = self.stage_seconds
        if ss is None:
            yield
            return
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.torch.cuda.synchronize()
            ss[name] = ss.get(name, 0.0) + time.perf_counter() - t0

    # ---- plumbing ---------------------------------------------------------------------------------
    def _s(self):
        return self.torch.cuda.current_stream().cuda_stream

    def _call(self, name, *args, launches=1):
        self.kn.launches += launches
        return _lib.call(name, *args)

    def to_device(self, a):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a
        a = np.ascontiguousarray(a, dtype=np.float64)
        if not a.flags.writeable:              # read-only stage results / pp_data views: torch warns, nothing is written
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", UserWarning)
                return t.from_numpy(a).to(self.dev)
        return t.from_numpy(a).to(self.dev)

    def empty(self, *shape, dtype=None):
        t = self.torch
        return t.empty(shape, dtype=dtype or t.float64, device=self.dev)

    def zeros(self, *shape, dtype=None):
        t = self.torch
        return t.zeros(shape, dtype=dtype or t.float64, device=self.dev)

    # ---- a12 distances ------------------------------------------------------------------------------
    def _pairwise(self, Xd, F, N, op, out, comm=None):
        """N x N pairwise kernel; with `comm` (world > 1) each rank computes a contiguous block of rows and the
        blocks are all-gathered (SC3 distances shard by row blocks, X replicated)."""
        if comm is None or comm.world <= 1:
            self._call("scrna_pairwise_f64", Xd, N, F, N, op, out, N, self._s())
            return out
        from .nmf_solver import shard_bounds
        r0, r1 = shard_bounds(N, comm.world, comm.rank)
        block = out[r0:r1]                       # contiguous rows of the final matrix
        self._call("scrna_pairwise_rows_f64", Xd, N, F, N, op, r0, r1, block, N, self._s())
        bounds = [shard_bounds(N, comm.world, r) for r in range(comm.world)]
        comm.all_gather_rows(out, bounds)
        return out

    def distances(self, X, metric="euclidean", comm=None):
        """X: genes x cells (numpy or device fp64) -> device n x n distance matrix."""
        with self._stage("upload"):
            Xd = self.to_device(X)
        with self._stage("distances"):
            return self._distances(Xd, metric, comm)

    def _distances(self, Xd, metric, comm):
        F, N = Xd.shape
        out = self.empty(N, N)
        if metric == "euclidean":
            return self._pairwise(Xd, F, N, 0, out, comm)
        if metric == "pearson":
            work = Xd.clone()
        elif metric == "spearman":
            Xt = self.empty(N, F)
            self._call("scrna_transpose_f64", Xd, N, F, N, Xt, F, self._s())
            Rt = self.empty(N, F)
            self._call("scrna_rank_avg_f64", Xt, F, N, F, Rt, F, self._s())
            work = self.empty(F, N)
            self._call("scrna_transpose_f64", Rt, F, N, F, work, N, self._s())
        else:
            raise ValueError("metric must be 'euclidean', 'pearson' or 'spearman' (got %r)" % (metric,))
        mean = self.empty(N)
        self._call("scrna_col_center_f64", work, N, F, N, mean, self._s(), launches=2)
        G = self._pairwise(work, F, N, 1, self.empty(N, N), comm)
        self._call("scrna_corr_epilogue_f64", G, N, N, F, out, N, self._s())
        return out

    # ---- a13 transformations ------------------------------------------------------------------------
    def transform_matrix(self, D, method="pca"):
        with self._stage("transform"):
            return self._transform_matrix(D, method)

    def _transform_matrix(self, D, method):
        D = self.to_device(D)
        n = D.shape[0]
        out = self.empty(n, n)
        if method == "spectral":
            work = self.empty(2 * n + 1)
            self._call("scrna_laplacian_f64", D, n, n, out, n, work, self._s(), launches=3)
        else:
            sd = self.empty(n)
            scratch = self.zeros(1, dtype=self.torch.int32)
            self._call("scrna_pca_prepare_f64", D, n, n, out, n, sd, scratch, self._s(), launches=3)
            if int(scratch.cpu()[0]) == 0:
                print("All values are Zero!")
        return out

    EIGH_MIN_N = 6000    # from here on the tridiagonalisation solver replaces the Jacobi sweeps (16 x 4 n^3 FMA)

    def right_singular_vectors(self, M, components, max_sweeps=40, tol=1e-15, blocked=None, method="auto", comm=None):
        """First `components` right singular vectors of the n x n device matrix M (descending singular
        values); returns (V device n x components, singular values numpy n).
        method 'jacobi': one-sided Jacobi on M itself -- n > 64 the block algorithm (sc3_svd_block.cu: GEMM-shaped,
        FP64-pipe bound), smaller matrices the unblocked kernel; every parity test against the reference's vectors
        runs here.  'eigh': leading eigenvectors of M^T M by tridiagonalisation + bisection + inverse iteration
        (sc3_eigh.cu), a fraction of one Jacobi sweep; only the first `components` singular values are returned
        (the rest of the array is zero).  'auto': eigh from n = 6000 (the BASELINE 20 000-cell configuration)."""
        t = self.torch
        n = M.shape[0]
        self._last = None
        if method not in ("auto", "jacobi", "eigh"):
            raise ValueError("method must be 'auto', 'jacobi' or 'eigh'")
        if method == "eigh" or (method == "auto" and blocked is None and n >= self.EIGH_MIN_N):
            with self._stage("svd"):
                return self._rsv_eigh(M, components, comm)
        if blocked is None:
            blocked = n > 64
        if blocked:
            with self._stage("svd"):
                return self._rsv_blocked(M, components, max_sweeps, tol)
        AT = self.empty(n, n)
        VT = self.empty(n, n)
        self._call("scrna_jacobi_init_f64", M, n, n, AT, VT, n, self._s(), launches=2)
        rotated = self.zeros(1, dtype=t.int32)
        self.jacobi_sweeps = 0
        if n > 1:
            for _ in range(max_sweeps):
                rotated.zero_()
                self._call("scrna_jacobi_sweep_f64", AT, VT, n, n, float(tol), rotated, self._s(),
                           launches=((n + 1) // 2 * 2) - 1)
                self.jacobi_sweeps += 1
                if int(rotated.cpu()[0]) == 0:
                    break
        sq = self.empty(n)
        self._call("scrna_row_sqnorm_f64", AT, n, n, sq, self._s())
        vals = np.sqrt(np.maximum(sq.cpu().numpy(), 0.0))
        order = np.argsort(-vals, kind="stable").astype(np.int32)
        c = int(components)
        od = t.from_numpy(order[:c].copy()).to(self.dev)
        V = self.empty(n, c)
        self._call("scrna_gather_vectors_f64", VT, n, n, od, c, V, c, self._s())
        return V, vals[order]

    def _rsv_blocked(self, M, components, max_sweeps, tol):
        t = self.torch
        n = M.shape[0]
        n_pad = 32 * _lib.call("scrna_bjacobi_blocks", n)
        W = self.zeros(n_pad, n)
        W[:n].copy_(M)                       # N = M^T: column j of N is row j of M
        work = self.empty(_lib.call("scrna_bjacobi_work_elems", n))
        rotated = self.zeros(1, dtype=t.int32)
        rounds = (n_pad // 32 + 1) // 2 * 2 - 1
        self.jacobi_sweeps = 0
        # numerically-null threshold: |M|_F^2 >= sigma_max^2 bounds it from above, row norms from below
        sq0 = self.empty(n)
        self._call("scrna_row_sqnorm_f64", W, n, n, sq0, self._s())
        null_sq = 1e-24 * float(sq0.max().cpu())
        for _ in range(max_sweeps):
            rotated.zero_()
            self._call("scrna_bjacobi_sweep_f64", W, n, n, float(tol), null_sq, work, rotated, self._s(),
                       launches=3 * rounds)
            self.jacobi_sweeps += 1
            if int(rotated.cpu()[0]) == 0:
                break
        nb = n_pad // 32
        self.svd_flops = self.jacobi_sweeps * (nb * (nb - 1) / 2.0) * (2 * 64 * 64 * n) * 2.0   # Gram + apply, 2 flop/FMA
        sq = self.empty(n)
        self._call("scrna_row_sqnorm_f64", W, n, n, sq, self._s())
        vals = np.sqrt(np.maximum(sq.cpu().numpy(), 0.0))
        order = np.argsort(-vals, kind="stable").astype(np.int32)
        c = int(components)
        with np.errstate(divide="ignore"):
            inv = np.where(vals > 0, 1.0 / vals, 0.0)
        od = t.from_numpy(order[:c].copy()).to(self.dev)
        sc = t.from_numpy(inv).to(self.dev)
        V = self.empty(n, c)
        self._call("scrna_gather_scaled_f64", W, n, n, od, sc, c, V, c, self._s())
        self._last = (W, od, inv[order[:c]], vals[order][:c])   # for rebuilt_matrix()
        return V, vals[order]

    def _ormtr(self, A, n, tau, ZT, c):
        """Rows of ZT (c x n) <- Q z with Q from scrna_sytrd_f64's reflectors: in blocks of 32 reflectors as two
        FP64 tensor-core GEMMs per block (compact WY) when n is even, else reflector by reflector."""
        if n % 2 == 0 and os.environ.get("SCRNA_B200_ORMTR", "blocked") != "legacy":
            work = self.empty(_lib.call("scrna_ormtr_work_elems", int(n), int(c)))
            self._call("scrna_ormtr_blocked_f64", A, n, n, tau, ZT, int(c), work, self._s(),
                       launches=1 + 3 * ((n - 2 + 31) // 32))
        else:
            self._call("scrna_ormtr_f64", A, n, n, tau, ZT, int(c), self._s())

    def _rsv_eigh(self, M, components, comm=None):
        """Leading `components` eigenpairs of A = M^T M (sc3_eigh.cu).  The host does O(n) bookkeeping only:
        Gershgorin bounds of the tridiagonal matrix and dstein's separation of numerically coincident eigenvalues."""
        t = self.torch
        n = M.shape[0]
        c = int(components)
        if n < 3 or c > n:
            raise ValueError("eigh path needs n >= 3 and components <= n")
        A = self.empty(n, n)
        with self._stage("svd.gram"):
            self._pairwise(M, n, n, 1, A, comm)                                             # A = M^T M (row blocks over the ranks)
        d, e, tau = self.empty(n), self.zeros(n), self.empty(n)
        work = self.empty(_lib.call("scrna_sytrd_work_elems", n))
        with self._stage("svd.tridiagonalise"):
            self._call("scrna_sytrd_f64", A, n, n, d, e, tau, work, self._s(), launches=6 * (n - 2) + (n - 2) // 32 + 1)
        dh, eh = d.cpu().numpy(), e.cpu().numpy()[: n - 1]
        off = np.zeros(n)
        off[:-1] += np.abs(eh)
        off[1:] += np.abs(eh)
        gl, gu = float(np.min(dh - off)), float(np.max(dh + off))
        tnorm = max(abs(gl), abs(gu))
        eps = np.finfo(np.float64).eps
        gl -= 2.0 * tnorm * eps * n + 2e-300
        gu += 2.0 * tnorm * eps * n + 2e-300
        pivmin = max(np.finfo(np.float64).tiny * max(1.0, float(np.max(eh * eh)) if n > 1 else 1.0), 1e-300)
        lam = self.empty(c)
        with self._stage("svd.bisection"):
            self._call("scrna_stebz_top_f64", d, e, n, c, gl, gu, pivmin, lam, self._s())
        lh = lam.cpu().numpy()                      # descending
        # dstein: numerically coincident eigenvalues are separated by 10 eps |lambda| so that inverse iteration does
        # not return the same vector twice
        lp = lh.copy()
        for j in range(c - 2, -1, -1):              # ascending walk: lp[j] must exceed lp[j+1]
            sep = 10.0 * eps * max(abs(lp[j + 1]), tnorm * eps)
            if lp[j] - lp[j + 1] < sep:
                lp[j] = lp[j + 1] + sep
        lam.copy_(t.from_numpy(lp).to(self.dev))
        nc = n * c
        work2 = self.empty(4 * nc)
        iwork = t.empty(nc, dtype=t.int8, device=self.dev)
        X = self.empty(n, c)
        with self._stage("svd.inverse_iteration"):
            self._call("scrna_stein_f64", d, e, n, c, lam, float(tnorm), work2, iwork, X, 3, self._s())
        del work2, iwork
        ZT = self.empty(c, n)
        self._call("scrna_transpose_f64", X, c, n, c, ZT, n, self._s())
        with self._stage("svd.back_transform"):
            if comm is None or comm.world <= 1:
                self._ormtr(A, n, tau, ZT, c)
            else:   # the vectors are independent: each rank back-transforms a block of them, then all-gather
                from .nmf_solver import shard_bounds
                bounds = [shard_bounds(c, comm.world, r) for r in range(comm.world)]
                c0, c1 = bounds[comm.rank]
                if c1 > c0:
                    self._ormtr(A, n, tau, ZT[c0:c1], c1 - c0)
                comm.all_gather_rows(ZT, bounds)
        V = X
        self._call("scrna_transpose_f64", ZT, n, c, n, V, c, self._s())
        vals = np.zeros(n)
        vals[:c] = np.sqrt(np.maximum(lh, 0.0))
        self.jacobi_sweeps = 0
        self.svd_flops = 2.0 * (float(n) ** 3) * (1.0 + 4.0 / 3.0) + 4.0 * float(n) * n * c   # Gram + sytrd + back-transform
        od = t.arange(c, dtype=t.int32, device=self.dev)
        self._last = (ZT, od, np.ones(c), vals[:c])   # rows are unit vectors: rebuilt_matrix scales them itself
        return V, vals

    def rebuilt_matrix(self, weights):
        """vecs . diag(weights) . vecs^T of the vectors just computed by the block solver (the first return value of
        the reference's `transformations`, sc3_clustering_impl.py:196-203) as a Gram on the device: rows
        sqrt(w_r) v_r, then the pairwise dot kernel -- instead of an n x c x n product on the host."""
        t = self.torch
        W, od, inv, _ = self._last
        n = W.shape[1]
        c = od.numel()
        w = np.sqrt(np.maximum(np.asarray(weights, dtype=np.float64)[:c], 0.0)) * inv
        Yt = self.empty(c, n)
        self._call("scrna_gather_rows_scaled_f64", W, n, n, od, t.from_numpy(w).to(self.dev), c, Yt, n, self._s())
        out = self.empty(n, n)
        self._call("scrna_pairwise_f64", Yt, n, c, n, 1, out, n, self._s())
        return out

    # ---- a14 k-means ----------------------------------------------------------------------------------
    def kmeans(self, V, d, k, draws, n_init=10, max_iter=10000, batched=None):
        """KMeans(n_clusters=k, n_init, max_iter, init='k-means++').fit_predict(V[:, :d]); `draws` are the
        n_init * kmeans_draw_count(k) uniforms of the reference's RNG stream.  Returns int64 host labels.
        batched: run the Lloyd iterations of all restarts in one launch set (None: whenever the kernels take the
        shape -- n_init <= 16, k <= 48); False keeps one restart after the other (any k <= 128)."""
        with self._stage("kmeans_ensemble"):
            return self._kmeans(V, d, k, draws, n_init, max_iter, batched)

    KM_POLL = 8   # Lloyd iterations enqueued between two polls of the device-side stop flags

    def _kmeans(self, V, d, k, draws, n_init, max_iter, batched=None):
        t = self.torch
        V = self.to_device(V)
        n, ldv = V.shape[0], V.stride(0)
        if V.stride(1) != 1 or ldv < d:
            V = V.contiguous()
            ldv = V.stride(0)
        per, trials = kmeans_draw_count(k)
        draws = np.ascontiguousarray(draws, dtype=np.float64)
        if draws.size < n_init * per:
            raise ValueError("need %d uniforms, got %d" % (n_init * per, draws.size))
        can_batch = bool(_lib.call("scrna_lloyd_batched_supported", int(k), int(n_init)))
        if batched is None:
            batched = can_batch
        elif batched and not can_batch:
            raise ValueError("the batched k-means kernels take n_init <= 16, k <= 48 and n_init * k <= 200")
        i32 = t.int32
        ld = ((d + 3) // 4) * 4                 # 32-byte rows: the batched assignment stages them with 16-byte copies
        X = self.empty(n, ld)
        xsq = self.empty(n)
        work = self.empty(_lib.call("scrna_km_prepare_work_elems", int(d)))
        R = int(n_init)
        state = self.zeros(R, 8)                # one control block per restart ([0] tol, [3] inertia, [6] done, [7] iterations)
        self._call("scrna_km_prepare", V, ldv, n, d, X, ld, xsq, work, state, self._s(), launches=3)
        state[1:, 0] = state[0, 0]
        U = t.from_numpy(draws).to(self.dev)
        closest = self.empty(R, n)
        cand_dist = self.empty(R, trials * n)
        cums = self.empty(R, n)
        pots = self.empty(R, 8)
        cand = self.zeros(R, 8, dtype=i32)
        idx = self.zeros(R, k, dtype=i32)
        bufs = (self.zeros(R, k, ld), self.zeros(R, k, ld))   # centres of all restarts, ping-pong
        labels = self.zeros(R, n, dtype=i32)
        labels_old = self.zeros(R, n, dtype=i32)
        labels_old.fill_(-1)
        counts = self.zeros(R, k, dtype=i32)
        celld = self.empty(R, n)
        esz = 8
        self.km_iters = 0
        # first centre of every restart: RandomState.choice(n, p=full(n, 1/n)) (SP/sklearn/cluster/_kmeans.py:231) is
        # searchsorted(cdf, u, side='right') on the sequentially accumulated cdf -- NumPy's own arithmetic, once per n
        cdf = self._uniform_cdf.get(n)
        if cdf is None:
            cdf = uniform_choice_cdf(n)
            self._uniform_cdf = {n: cdf}
        firsts = first_centres(cdf, draws[: R * per: per])
        # k-means++ of all restarts in one launch sequence (the restart is a grid dimension of every kernel)
        firsts_d = t.from_numpy(firsts.astype(np.int32)).to(self.dev)
        self._call("scrna_kmeans_pp_batched", X, ld, n, d, xsq, k, R, firsts_d, U, trials, closest, cand_dist, cums,
                   pots, cand, idx, state, bufs[0], ld, self._s(), launches=4 * k + 1)
        if batched:
            bwork = self.empty(_lib.call("scrna_lloyd_batched_work_elems", int(k), int(d), R))
            queued, st = 0, None
            while queued < max_iter:
                nb = min(self.KM_POLL, max_iter - queued)
                for b in range(nb):
                    cur, nxt = bufs[(queued + b) % 2], bufs[(queued + b + 1) % 2]
                    self._call("scrna_lloyd_iteration_batched", X, ld, n, d, cur, nxt, ld, k, R, labels, labels_old,
                               counts, celld, bwork, state, int(max_iter), self._s(), launches=6)
                queued += nb
                st = state.cpu().numpy()              # one poll per batch: every restart's stop flag
                if (st[:, 6] != 0.0).all():
                    break
        else:
            swork = self.empty(_lib.call("scrna_lloyd_sums_work_elems", k, d))
            for r in range(R):
                queued = 0
                while queued < max_iter:
                    nb = min(self.KM_POLL, max_iter - queued)
                    for b in range(nb):
                        cur, nxt = bufs[(queued + b) % 2][r], bufs[(queued + b + 1) % 2][r]
                        self._call("scrna_lloyd_iteration", X, ld, n, d, cur, nxt, ld, k, labels[r], labels_old[r],
                                   counts[r], celld[r], swork, state[r], int(max_iter), self._s(), launches=5)
                    queued += nb
                    if float(state[r, 6].cpu()) != 0.0:   # one poll per batch (the stop rule itself lives on the device)
                        break
            st = state.cpu().numpy()
        for r in range(R):
            done_iters = int(st[r, 7])
            self.km_iters += done_iters
            cur = bufs[done_iters % 2][r]           # the centres the last executed iteration of this restart wrote
            if st[r, 6] != 1.0:                      # not strict convergence: labels of the final centres
                self._call("scrna_lloyd_assign", X, ld, n, d, cur, ld, k, labels[r], self._s())
            self._call("scrna_km_inertia", X, ld, n, d, cur, ld, labels[r], celld[r], state[r], self._s(), launches=2)
        # restart selection on the host from ONE read-back (sklearn _kmeans.py:1534-1541)
        inert = state[:, 3].cpu().numpy()
        labs = labels.cpu().numpy()
        best = None
        for r in range(R):
            if best is None or (inert[r] < best[1] and not same_clustering(labs[r], best[0], k)):
                best = (labs[r].copy(), float(inert[r]))
        return best[0]

    # ---- a15 / a17 consensus --------------------------------------------------------------------------
    def coassoc_new(self, n):
        return self.zeros(n, n, dtype=self.torch.int32)

    def coassoc_add(self, M, labels):
        t = self.torch
        n = M.shape[0]
        if not isinstance(labels, t.Tensor):
            labels = t.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(self.dev)
        with self._stage("coassociation"):
            self._call("scrna_coassoc_accum", labels, n, M, n, self._s())

    def consensus_from_counts(self, M, cnt):
        n = M.shape[0]
        C = self.empty(n, n)
        self._call("scrna_consensus_scale", M, n, n, float(cnt), C, n, self._s())
        return C

    def consensus_clustering(self, C, n_clusters, comm=None):
        """pdist(C) -> complete linkage -> cut_tree(n_clusters).  Returns (labels numpy int64, device n x n
        row-distance matrix, linkage matrix Z numpy (n-1) x 4)."""
        t = self.torch
        C = self.to_device(C)
        n = C.shape[0]
        D = self.empty(n, n)
        with self._stage("consensus_row_distances"):
            self._pairwise(C, n, n, 0, D, comm)
        with self._stage("linkage"):
            return self._linkage(D, n, n_clusters)

    def _linkage(self, D, n, n_clusters):
        t = self.torch
        Dwork = D.clone()
        labels = self.zeros(n, dtype=t.int32)
        Z = self.zeros(max(n - 1, 1), 4)
        merges = self.zeros(max(n - 1, 1), 3)
        iwork = self.zeros(8 * n + 4 * (2 * n - 1), dtype=t.int32)
        if n >= 2:
            self._call("scrna_hclust_complete", Dwork, n, n, int(n_clusters), labels, Z, merges, iwork, self._s(),
                       launches=5)
        return labels.cpu().numpy().astype(np.int64), D, Z.cpu().numpy()


_ops = None


def default_ops():
    global _ops
    if _ops is None:
        _ops = Sc3Ops()
    return _ops
