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
from .nmf_solver import DeviceMatrix, NmfSolver
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


def _fit_nmf(X, k, alpha, l1, max_iter, tol, solver, init, comm=None):
    """decomp.NMF(init='nndsvdar', solver='cd', shuffle=True, random_state=0).fit_transform(X).

    init: (W0, H0) | 'host' (sklearn's routine on the host, bit-identical to the reference's start) |
    'device' (same algorithm and RNG stream with every pass over X on the GPU, scrna_b200/nmf_init.py) |
    None = 'device' for matrices of at least 2^24 elements (where the host passes cost seconds to minutes)
    on a single GPU with k <= 54, else 'host'."""
    l1r, l2r = _regularization(alpha, l1)
    dm = DeviceMatrix.from_host(X, _kernels())
    XT = None
    if init is None:
        big = X.shape[0] * X.shape[1] >= DEVICE_INIT_MIN_ELEMENTS
        init = 'device' if (big and comm is None and k <= 54) else 'host'
    if isinstance(init, str):
        if init == 'device':
            from .nmf_init import nndsvdar_init_device
            W0, H0, XT = nndsvdar_init_device(dm, k, _kernels(), seed=0)
        elif init == 'host':
            W0, H0 = nndsvdar_init(X, k)
        else:
            raise ValueError("init must be (W0, H0), 'host' or 'device'")
    else:
        W0, H0 = init
    slv = NmfSolver(dm, k, comm=comm, XT=XT)
    res = slv.fit(W0, H0, solver=solver, l1_G=l1r, l2_G=l2r, l1_C=l1r, l2_C=l2r, tol=tol,
                  max_iter=max_iter, first="G", seed=0, shuffle=True)
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

    def apply(self, k=-1, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3, solver='cd', init=None):
        if k == -1:
            k = self.num_cluster
        X = self.pre_processing()
        res = _fit_nmf(X, k, alpha, l1, max_iter, rel_err, solver, init)
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

    def apply(self, k=-1, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3, solver='cd'):
        if k == -1:
            k = self.num_cluster
        X = self.pre_processing()
        g, n = X.shape
        onehot = _one_hot(np.asarray(self.labels)[self.remain_cell_inds]
                          if len(self.labels) == self.num_cells else self.labels)
        if onehot.shape[1] != k:
            raise ValueError('Array with wrong shape: %d unique labels but k=%d' % (onehot.shape[1], k))
        dm = DeviceMatrix.from_host(X, _kernels())
        slv = NmfSolver(dm, k)
        # (1) non_negative_factorization(X, H=onehot^T, update_H=False): dictionary from ZEROS with the
        #     cell factor fixed, no regularisation (nmf_clustering.py:66; _nmf.py:1224-1228)
        r1 = slv.fit(np.zeros((g, k)), onehot.T, solver=solver if solver == 'cd' else 'cd', tol=rel_err,
                     max_iter=max_iter, update_G=True, update_C=False, first="G", seed=0, shuffle=True)
        # (2) NMF(init='custom').fit_transform(X^T, W=onehot, H=dictionary^T): sklearn's W is the cell
        #     factor and is updated first (nmf_clustering.py:71-72)
        l1r, l2r = _regularization(alpha, l1)
        r2 = slv.fit(r1.G, onehot.T, solver=solver, l1_G=l1r, l2_G=l2r, l1_C=l1r, l2_C=l2r, tol=rel_err,
                     max_iter=max_iter, first="C", seed=0, shuffle=True)
        W = r2.C.T            # sklearn W: cells x k
        H = r2.G.T            # sklearn H: k x genes
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

    def __init__(self, src, trg_data, trg_gene_ids, num_cluster):
        super(DaNmfClustering, self).__init__(trg_data, gene_ids=trg_gene_ids, num_cluster=num_cluster, labels=[])
        self.src = src

    def get_mixed_data(self, mix=0.0, reject_ratio=0., use_H2=True, calc_transferability=False, max_iter=100,
                       rel_err=1e-3, H_init=None):
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
        sess = TransferSession(W, trg_data)
        H0 = transfer_h_init(k, n_t) if H_init is None else np.asarray(H_init, dtype=np.float64)
        sess.solve(H0, max_iter=max_iter, rel_err=rel_err)
        H = sess.H_host()
        labels = sess.labels()
        H2 = np.zeros((k, n_t))
        H2[(labels, np.arange(n_t))] = 1
        self.cluster_labels = labels                      # == np.argmax(H, axis=0)
        self.intermediate_model = (W, H, H2)
        self.reject = self.calc_rejection(trg_data, W, H, H2, _session=sess)

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
        else:
            new_trg_data = W.dot(H)
            mixed_data = mix * new_trg_data + (1. - mix) * trg_data
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
        res = _fit_nmf(mixed_data, k, alpha, l1, max_iter, 1e-6, solver, None)
        self.dictionary = res.G
        self.data_matrix = res.C
        self.cluster_labels = np.argmax(res.C, axis=0)
        self.mixed_data = mixed_data
        self.n_iter = res.n_iter

    def calc_rejection(self, trg_data, W, H, H2, _session=None):
        """Rejection scores (nmf_clustering.py:225-272): same list order and names.  The five
        genes x cells products of the reference are one fused sweep on the device; the k x n_t
        statistics of H stay on the host."""
        sess = _session if _session is not None else TransferSession(W, trg_data)
        if _session is None:
            sess._set_H(H)
            sess.labels()
        st = sess.reject_stats()
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
