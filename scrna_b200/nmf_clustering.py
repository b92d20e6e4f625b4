"""NMF clustering classes with the reference's API, running on the B200 kernels.

    NmfClustering        scRNA/nmf_clustering.py:11-46     NNDSVDar init + CD NMF, labels = argmax_k H
    NmfClustering_initW  scRNA/nmf_clustering.py:49-89     label-seeded dictionary, then CD NMF on X^T
    NmfClustering_fixW   scRNA/nmf_clustering.py:92-130    two fixed-factor solves (unused by the CLI)
    DaNmfClustering      scRNA/nmf_clustering.py:133-295   target transfer, mixed data, rejection stats

Same constructor / `apply` / `get_mixed_data` signatures and result attributes (`dictionary` g x k,
`data_matrix` k x n, `cluster_labels`, `intermediate_model`, `reject`, `mixed_data`); extra opt-in
keywords only: `solver` ('cd' reference solver | 'mu' sklearn multiplicative update), `init=(W0, H0)`
and `comm` (cell-sharded multi-GPU fit).  `alpha` / `l1` carry the semantics of the scikit-learn
the reference was written against (both factors, unscaled: l1_reg = alpha*l1, l2_reg = alpha*(1-l1);
SURVEY.md §A.2).
"""
import numpy as np
import scipy.stats as stats

from .abstract_clustering import AbstractClustering
from .nmf_solver import DeviceMatrix, NmfSolver, shard_bounds
from .transfer import TransferSession
from .utils import get_matching_gene_inds, transfer_h_init


def _regularization(alpha, l1_ratio):
    l1 = float(alpha) * float(l1_ratio)
    l2 = float(alpha) * (1.0 - float(l1_ratio))
    return l1, l2


def nndsvdar_init(X, k, random_state=0):
    """NMF(init='nndsvdar', random_state=0) start (SP/sklearn/decomposition/_nmf.py:214-366).

    SURVEY.md §8 row a4, phase 1: the init is produced on the host by the same library routine the
    reference calls and injected into the device solver ("identical init matrices"); the device
    randomized-SVD is the §8(f1) follow-up.  It is an input to the solver, not a fallback for it."""
    from sklearn.decomposition._nmf import _initialize_nmf
    return _initialize_nmf(np.asarray(X, dtype=np.float64), k, init="nndsvdar", random_state=random_state)


def _check_nan(W, H, alpha, k, l1, X):
    if np.any(np.isnan(H)):
        raise Exception('H contains NaNs (alpha={0}, k={1}, l1={2}, data={3}x{4}'.format(
            alpha, k, l1, X.shape[0], X.shape[1]))
    if np.any(np.isnan(W)):
        raise Exception('W contains NaNs (alpha={0}, k={1}, l1={2}, data={3}x{4}'.format(
            alpha, k, l1, X.shape[0], X.shape[1]))


DEVICE_INIT_MIN_ELEMENTS = 1 << 24   # below this the host routine costs milliseconds and is bit-faithful


class _Cells:
    """How the cell axis of a (genes x cells) host matrix maps onto the ranks of `comm`.

    cells='replicated' (default): every rank holds the same full matrix and works on its contiguous slice
    (shard_bounds); cell-indexed results are gathered, so every rank ends with the reference's full outputs.
    cells='sharded': `X` already is this rank's slice (atlas scale: the full matrix fits no single host);
    cell-indexed results stay local, gene-indexed ones are replicated."""

    def __init__(self, n_local_or_total, comm, cells):
        if cells not in ('replicated', 'sharded'):
            raise ValueError("cells must be 'replicated' or 'sharded'")
        self.comm = comm if (comm is not None and comm.world > 1) else None
        self.mode = cells
        n = int(n_local_or_total)
        if self.comm is None:
            self.c0, self.c1, self.n_total, self.bounds = 0, n, n, [(0, n)]
            return
        world, rank = self.comm.world, self.comm.rank
        if cells == 'replicated':
            self.n_total = n
            self.bounds = [shard_bounds(n, world, r) for r in range(world)]
        else:
            import torch
            cnt = torch.zeros(world, dtype=torch.int64)
            cnt[rank] = n
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and \
                self.comm.dist.get_backend(self.comm.group) == "nccl" else torch.device("cpu")
            cnt = cnt.to(dev)
            self.comm.all_reduce(cnt)
            sizes = [int(v) for v in cnt.cpu().tolist()]
            offs = np.concatenate([[0], np.cumsum(sizes)])
            self.n_total = int(offs[-1])
            self.bounds = [(int(offs[r]), int(offs[r + 1])) for r in range(world)]
        self.c0, self.c1 = self.bounds[rank]

    def local(self, A, axis=1):
        """This rank's slice of a cell-indexed host array (no-op when the data came in sharded)."""
        if self.comm is None or self.mode == 'sharded':
            return A
        return np.ascontiguousarray(A[:, self.c0:self.c1] if axis == 1 else A[self.c0:self.c1])

    def gather(self, A_local, kernels):
        """Full (r x n_total) host array from every rank's (r x n_local) block; identity for sharded input."""
        if self.comm is None or self.mode == 'sharded':
            return A_local
        t = kernels.torch
        loc = t.from_numpy(np.ascontiguousarray(A_local, dtype=np.float64))
        if t.cuda.is_available() and self.comm.dist.get_backend(self.comm.group) == "nccl":
            loc = loc.cuda()
        return self.comm.gather_cols(loc, self.bounds, self.n_total).cpu().numpy()

    @property
    def range(self):
        return (self.c0, self.c1, self.n_total) if self.comm is not None else None


def _fit_nmf(X, k, alpha, l1, max_iter, tol, solver, init, comm=None, cells='replicated'):
    """decomp.NMF(init='nndsvdar', solver='cd', shuffle=True, random_state=0).fit_transform(X).

    init: (W0, H0) | 'host' (sklearn's routine on the host, bit-identical to the reference's start) |
    'device' (same algorithm and RNG stream with every pass over X on the GPU, scrna_b200/nmf_init.py) |
    None = 'device' for matrices of at least 2^24 elements (where the host passes cost seconds to minutes)
    or cell-sharded input, else 'host'.
    comm: cells sharded over the ranks (one all-reduce per iteration); see `_Cells`.
    Returns the solver result with G (genes x k, replicated) and C (k x cells: all cells, or this rank's
    with cells='sharded')."""
    l1r, l2r = _regularization(alpha, l1)
    kn = _kernels()
    part = _Cells(X.shape[1], comm, cells)
    dm = DeviceMatrix.from_host(part.local(X), kn)
    XT = None
    if init is None:
        big = X.shape[0] * part.n_total >= DEVICE_INIT_MIN_ELEMENTS
        init = 'device' if (big or (part.comm is not None and cells == 'sharded')) else 'host'
    if isinstance(init, str):
        if init == 'device':
            from .nmf_init import nndsvdar_init_device
            W0, H0, XT = nndsvdar_init_device(dm, k, kn, seed=0, comm=part.comm, cell_range=part.range)
        elif init == 'host':
            if part.comm is not None and cells == 'sharded':
                raise ValueError("init='host' needs the whole matrix on every rank (cells='replicated')")
            W0, H0 = nndsvdar_init(X, k)
        else:
            raise ValueError("init must be (W0, H0), 'host' or 'device'")
    else:
        W0, H0 = init
    if H0.shape[1] == part.n_total and part.comm is not None:
        H0 = np.ascontiguousarray(H0[:, part.c0:part.c1])
    slv = NmfSolver(dm, k, kernels=kn, comm=part.comm, XT=XT)
    res = slv.fit(W0, H0, solver=solver, l1_G=l1r, l2_G=l2r, l1_C=l1r, l2_C=l2r, tol=tol,
                  max_iter=max_iter, first="G", seed=0, shuffle=True)
    res.C = part.gather(res.C, kn)
    return res


def _kernels():
    from .kernels import default_kernels
    return default_kernels()


class NmfClustering(AbstractClustering):
    num_cluster = -1
    dictionary = None
    data_matrix = None
    n_iter = None

    def __init__(self, data, gene_ids, num_cluster, labels=None):
        super(NmfClustering, self).__init__(data, gene_ids=gene_ids)
        self.num_cluster = num_cluster

    def apply(self, k=-1, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3, solver='cd', init=None, comm=None,
              cells='replicated'):
        """Reference signature (nmf_clustering.py:20) plus opt-in `solver`, `init`, and `comm` / `cells`: a
        TorchDistComm of the ranks that all call apply() -- the cells are then sharded over their GPUs (see
        `_Cells`); with cells='replicated' every rank ends with the same `dictionary`, `data_matrix`, labels."""
        if k == -1:
            k = self.num_cluster
        X = self.pre_processing()
        res = _fit_nmf(X, k, alpha, l1, max_iter, rel_err, solver, init, comm=comm, cells=cells)
        W, H = res.G, res.C
        self.cluster_labels = np.argmax(H, axis=0)
        _check_nan(W, H, alpha, k, l1, X)
        self.dictionary = W
        self.data_matrix = H
        self.n_iter = res.n_iter

    def print_reconstruction_error(self, X, W, H):
        R = X - W.dot(H)
        print(('  Elementwise absolute reconstruction error   : ', np.sum(np.abs(R)) / float(X.size)))
        print(('  Fro-norm reconstruction error               : ', np.sqrt(np.sum(R * R)) / float(X.size)))


def _one_hot(labels):
    """pd.get_dummies(labels): one column per sorted unique label (nmf_clustering.py:64)."""
    labels = np.asarray(labels)
    uniq = np.unique(labels)
    return (labels[:, None] == uniq[None, :]).astype(np.float64)


class NmfClustering_initW(AbstractClustering):
    num_cluster = -1
    dictionary = None
    data_matrix = None
    n_iter = None

    def __init__(self, data, gene_ids, num_cluster, labels):
        super(NmfClustering_initW, self).__init__(data, gene_ids=gene_ids)
        self.num_cluster = num_cluster
        self.labels = labels

    def apply(self, k=-1, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3, solver='cd', comm=None):
        """Reference signature (nmf_clustering.py:59) plus opt-in `solver` and `comm` (cells sharded over the ranks
        that all call apply() with the same data and labels; every rank ends with the same results)."""
        if k == -1:
            k = self.num_cluster
        X = self.pre_processing()
        g, n = X.shape
        onehot = _one_hot(np.asarray(self.labels)[self.remain_cell_inds]
                          if len(self.labels) == self.num_cells else self.labels)
        if onehot.shape[1] != k:
            raise ValueError('Array with wrong shape: %d unique labels but k=%d' % (onehot.shape[1], k))
        kn = _kernels()
        part = _Cells(n, comm, 'replicated')
        onehot = part.local(onehot, axis=0)
        dm = DeviceMatrix.from_host(part.local(X), kn)
        slv = NmfSolver(dm, k, kernels=kn, comm=part.comm)
        # (1) non_negative_factorization(X, H=onehot^T, update_H=False): dictionary from ZEROS with the
        #     cell factor fixed, no regularisation (nmf_clustering.py:66; _nmf.py:1224-1228)
        r1 = slv.fit(np.zeros((g, k)), onehot.T, solver=solver if solver == 'cd' else 'cd', tol=rel_err,
                     max_iter=max_iter, update_G=True, update_C=False, first="G", seed=0, shuffle=True)
        # (2) NMF(init='custom').fit_transform(X^T, W=onehot, H=dictionary^T): sklearn's W is the cell
        #     factor and is updated first (nmf_clustering.py:71-72)
        l1r, l2r = _regularization(alpha, l1)
        r2 = slv.fit(r1.G, onehot.T, solver=solver, l1_G=l1r, l2_G=l2r, l1_C=l1r, l2_C=l2r, tol=rel_err,
                     max_iter=max_iter, first="C", seed=0, shuffle=True)
        W = part.gather(r2.C, kn).T   # sklearn W: cells x k
        H = r2.G.T                    # sklearn H: k x genes
        self.cluster_labels = np.argmax(W, axis=1)
        _check_nan(W, H, alpha, k, l1, X)
        self.dictionary = H.T
        self.data_matrix = W.T
        self.n_iter = (r1.n_iter, r2.n_iter)

    print_reconstruction_error = NmfClustering.print_reconstruction_error


class NmfClustering_fixW(AbstractClustering):
    num_cluster = -1
    dictionary = None
    data_matrix = None

    def __init__(self, data, gene_ids, num_cluster, labels):
        super(NmfClustering_fixW, self).__init__(data, gene_ids=gene_ids)
        self.num_cluster = num_cluster
        self.labels = labels

    def apply(self, k=-1, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3):
        if k == -1:
            k = self.num_cluster
        X_t = self.pre_processing()                      # genes x cells
        g, n = X_t.shape
        onehot = _one_hot(self.labels)
        dm = DeviceMatrix.from_host(X_t, _kernels())
        slv = NmfSolver(dm, k)
        # dictionary with the cell factor fixed to the one-hot labels (nmf_clustering.py:110)
        r1 = slv.fit(np.zeros((g, k)), onehot.T, solver='cd', tol=rel_err, max_iter=max_iter,
                     update_G=True, update_C=False, first="G", seed=0, shuffle=True)
        # then the cell factor with that dictionary fixed (nmf_clustering.py:116)
        r2 = slv.fit(r1.G, np.zeros((k, n)), solver='cd', tol=rel_err, max_iter=max_iter,
                     update_G=False, update_C=True, first="C", seed=0, shuffle=True)
        learned_W = r2.C.T
        self.cluster_labels = np.argmax(learned_W, axis=1)
        _check_nan(onehot.T, r1.G, alpha, k, l1, X_t.T)
        self.dictionary = r1.G
        self.data_matrix = onehot.T


class DaNmfClustering(NmfClustering):
    reject = None
    transferability_score = 0.0
    transferability_percs = None
    transferability_rand_scores = None
    transferability_pvalue = 1.0
    src = None
    intermediate_model = None
    mixed_data = None

    def __init__(self, src, trg_data, trg_gene_ids, num_cluster, comm=None):
        """`comm` (opt-in, not in the reference): the target cells are sharded over the ranks that all construct
        this object with the same data (independent columns, one scalar all-reduce per transfer
        iteration for the stop rule); every rank ends with the same results."""
        super(DaNmfClustering, self).__init__(trg_data, gene_ids=trg_gene_ids, num_cluster=num_cluster, labels=[])
        self.src = src
        self.comm = comm if (comm is not None and comm.world > 1) else None

    def get_mixed_data(self, mix=0.0, reject_ratio=0., use_H2=True, calc_transferability=False, max_iter=100,
                       rel_err=1e-3, H_init=None, _target_cache=None):
        """Reference signature (nmf_clustering.py:147) plus opt-in `H_init` (inject the start instead of drawing
        it) and `_target_cache` (a dict shared by the caller between calls on the SAME target data: the uploaded
        target matrix is kept in it, so a (k, mixture) grid uploads it once -- scrna_b200/workflows.py)."""
        trg_data = self.pre_processing()
        trg_gene_ids = self.gene_ids[self.remain_gene_inds]
        src_gene_ids = self.src.gene_ids[self.src.remain_gene_inds].copy()
        inds1, inds2 = get_matching_gene_inds(src_gene_ids, trg_gene_ids)
        src_gene_ids = src_gene_ids[inds2]
        self.gene_ids = trg_gene_ids[inds1]
        trg_data = trg_data[inds1, :]
        assert np.all(src_gene_ids == self.gene_ids)
        assert self.src.dictionary is not None  # source data should always be pre-processed
        W = np.ascontiguousarray(self.src.dictionary[inds2, :], dtype=np.float64)
        k = W.shape[1]
        n_t = trg_data.shape[1]
        kn = _kernels()
        part = _Cells(n_t, self.comm, 'replicated')
        trg_dev = part.local(trg_data)
        if _target_cache is not None:
            key = (trg_data.shape, tuple(np.asarray(inds1[:8]).tolist()), int(np.asarray(inds1).sum()), part.c0, part.c1)
            if key not in _target_cache:
                _target_cache[key] = DeviceMatrix.from_host(trg_dev, kn)
            trg_dev = _target_cache[key]
        sess = TransferSession(W, trg_dev, kernels=kn, comm=part.comm)
        H0 = transfer_h_init(k, n_t) if H_init is None else np.asarray(H_init, dtype=np.float64)   # global RNG draw
        sess.solve(part.local(H0), max_iter=max_iter, rel_err=rel_err)
        H = part.gather(sess.H_host(), kn)
        labels = np.argmax(H, axis=0) if part.comm is not None else sess.labels()
        if part.comm is not None:
            sess.labels()
        H2 = np.zeros((k, n_t))
        H2[(labels, np.arange(n_t))] = 1
        self.cluster_labels = labels                      # == np.argmax(H, axis=0)
        self.intermediate_model = (W, H, H2)
        self.reject = self.calc_rejection(trg_data, W, H, H2, _session=sess, _part=part)

        if calc_transferability:   # nmf_clustering.py:175-183
            from .utils import get_transferability_score
            self.transferability_score, self.transferability_rand_scores, self.transferability_pvalue = \
                get_transferability_score(W, H, trg_data, max_iter=max_iter)
            self.transferability_percs = np.percentile(self.transferability_rand_scores, [25, 50, 75, 100])
            self.reject.append(('Transfer_Percentiles', self.transferability_percs))
            self.reject.append(('Transferability', self.transferability_score))
            self.reject.append(('Transferability p-value', self.transferability_pvalue))

        assert reject_ratio < 1.  # rejection of 100% (or more) does not make any sense
        if reject_ratio > 0.:
            # the reference slices with a float here (nmf_clustering.py:196-197) and fails the same way
            raise TypeError('slice indices must be integers or None or have an __index__ method')

        if use_H2:
            mixed_data, new_trg_data = sess.mixed(mix)
            mixed_data, new_trg_data = part.gather(mixed_data, kn), part.gather(new_trg_data, kn)
        else:
            mixed_data, new_trg_data = sess.mixed_dense(mix)      # mix . (W.H) + (1 - mix) . X on the device
            mixed_data, new_trg_data = part.gather(mixed_data, kn), part.gather(new_trg_data, kn)
        if np.any(trg_data < 0.0):
            print('Error! Negative values in target data!')
        if np.any(mixed_data < 0.0):
            print('Error! Negative values in reconstructed data!')
        return mixed_data, new_trg_data, trg_data

    def apply(self, k=-1, mix=0.0, reject_ratio=0., alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3,
              calc_transferability=False, solver='cd', H_init=None):
        if k == -1:
            k = self.num_cluster
        mixed_data, new_trg_data, trg_data = self.get_mixed_data(mix=mix, reject_ratio=reject_ratio,
                                                                 max_iter=max_iter, rel_err=rel_err,
                                                                 calc_transferability=calc_transferability,
                                                                 H_init=H_init)
        res = _fit_nmf(mixed_data, k, alpha, l1, max_iter, 1e-6, solver, None, comm=self.comm)
        self.dictionary = res.G
        self.data_matrix = res.C
        self.cluster_labels = np.argmax(res.C, axis=0)
        self.mixed_data = mixed_data
        self.n_iter = res.n_iter

    def calc_rejection(self, trg_data, W, H, H2, _session=None, _part=None):
        """Rejection scores (nmf_clustering.py:225-272): same list order and names.  The five
        genes x cells products of the reference are one fused sweep on the device; the k x n_t
        statistics of H stay on the host."""
        sess = _session if _session is not None else TransferSession(W, trg_data)
        if _session is None:
            sess._set_H(H)
            sess.set_labels(np.argmax(np.asarray(H2), axis=0))     # the caller's H2, not argmax(H), defines "W.H2"
        st = sess.reject_stats()
        if _part is not None:
            st = _part.gather(st, _kernels())      # per-cell statistics of every shard
        diffs = np.zeros(H2.shape[1])
        for c in range(self.src.num_cluster):
            inds = np.where(self.cluster_labels == c)[0]
            if inds.size > 0:
                min_h2 = np.min(H[:, inds])
                max_h2 = np.max(H[:, inds])
                foo = H[:, inds] - min_h2 / (max_h2 - min_h2)   # operator precedence as in the reference
                diffs[inds] = np.max(foo, axis=0) - np.min(foo, axis=0)
        sum_expr = st[0].copy()
        sum_expr -= np.min(sum_expr)
        sum_expr /= np.max(sum_expr)
        sum_expr = sum_expr + 1.0
        sum_expr /= np.max(sum_expr)
        weight = 1. - sum_expr
        reconstr_err = st[3].copy()
        reconstr_err -= np.min(reconstr_err)
        reconstr_err /= np.max(reconstr_err)
        final_values = weight * reconstr_err
        reject = [('Reconstr. Error', -final_values)]
        neg_entropy = stats.entropy(H)
        neg_entropy -= np.min(neg_entropy)
        neg_entropy /= np.max(neg_entropy)
        reject.append(('Kurtosis', stats.kurtosis(H, fisher=False, axis=0)))
        reject.append(('Entropy', -neg_entropy))
        reject.append(('Diffs', diffs))
        reject.append(('Dist L2 H', -st[2]))
        reject.append(('Dist L2 H2', -st[4]))
        reject.append(('Dist L1 H', -st[1]))
        reject.append(('Dist L1 H2', -st[3]))
        return reject

    def reject_classifier(self, K, kurts):
        """Binary keep / reject labels maximising the kernel-target alignment over the kurtosis ranking
        (scRNA/nmf_clustering.py:274-295; unused by the reference's CLI).  Host arithmetic on an n x n kernel."""
        from .utils import center_kernel, normalize_kernel, kta_align_binary
        sinds = np.argsort(kurts)
        K = normalize_kernel(center_kernel(K))
        max_kta, max_kta_ind = -1.0, -1
        for i in range(K.shape[1] - 2):
            labels = np.ones(kurts.size, dtype=int)
            labels[sinds[:i + 1]] = -1
            kta = kta_align_binary(K, labels)
            if kta > max_kta:
                max_kta, max_kta_ind = kta, i + 1
        labels = np.ones(kurts.size, dtype=int)
        labels[sinds[:max_kta_ind]] = -1
        return labels
