"""Hot-path helpers with the reference's names and signatures (scRNA/utils.py:187-278).

On the north-star path: `get_transferred_data_matrix` (:187-230) and `get_matching_gene_inds` (:233-278).
"Next" rows of SURVEY.md §8f built on the same device kernels: `unsupervised_acc_kta` (:134-152, row f3) and
`get_transferability_score` (:155-184, row f4).  TSV / npz loaders and the rbf kernel (:11-63, 108-124) are IO /
unused by the CLI and stay out of scope.
"""
import numpy as np

from .transfer import TransferSession


def transfer_h_init(k, n_cells):
    """H init of the reference (utils.py:189-193): |N(0,1)| clamped to >= 1e-10, drawn with ONE
    np.random.randn(k, n) call on the global legacy RNG so the stream matches."""
    H = np.random.randn(k, n_cells)
    H[H < 0.] *= -1.
    H[H < 1e-10] = 1e-10
    return H


def get_transferred_data_matrix(W, trg_data, normalize_H2=False, max_iter=100, rel_err=1e-3,
                                H_init=None, _session=None):
    """Solve trg_data ~ W.H for H >= 0 with W fixed (multiplicative updates on the GPU).

    Returns (W, H, H2, err) exactly like the reference: H2 is the one-hot argmax of H and err the last
    mean absolute reconstruction error.  `H_init` (k x n_t) injects the start instead of drawing it.
    """
    W = np.asarray(W, dtype=np.float64)
    k = W.shape[1]
    sess = _session if _session is not None else TransferSession(W, trg_data)
    n_t = sess.X.n
    H0 = transfer_h_init(k, n_t) if H_init is None else np.asarray(H_init, dtype=np.float64)
    _, new_err = sess.solve(H0, max_iter=max_iter, rel_err=rel_err)
    H = sess.H_host()
    labels = sess.labels()
    H2 = np.zeros((k, n_t))
    H2[(labels, np.arange(n_t))] = 1
    if normalize_H2:
        # second loop of the reference (utils.py:213-229): the same update applied to H2
        keep = (sess.H_host(), sess.n_iter, sess.err)
        sess.solve(H2, max_iter=max_iter, rel_err=rel_err)
        H2 = sess.H_host()
        sess._set_H(keep[0])
        sess.n_iter, sess.err = keep[1], keep[2]
        sess.labels()
    return W, H, H2, new_err


def get_matching_gene_inds(src_gene_ids, trg_gene_ids):
    """Order-preserving intersection of gene ids in TARGET order; the first occurrence of a
    duplicated id wins on both sides (utils.py:233-278).  Returns (inds into trg, inds into src).
    Dictionary-based, O(G) instead of the reference's O(G^2) scans, same result."""
    src_gene_ids = np.asarray(src_gene_ids)
    trg_gene_ids = np.asarray(trg_gene_ids)
    if not np.unique(src_gene_ids).size == src_gene_ids.size:
        print(('\nWarning! (MTL gene ids) Gene ids are supposed to be unique. '
               'Only {0} of {1}  entries are unique.'.format(np.unique(src_gene_ids).shape[0],
                                                             src_gene_ids.shape[0])))
        print('Only first occurance will be used.\n')
    if not np.unique(trg_gene_ids).size == trg_gene_ids.size:
        print(('\nWarning! (Target gene ids) Gene ids are supposed to be unique. '
               'Only {0} of {1}  entries are unique.'.format(np.unique(trg_gene_ids).shape[0],
                                                             trg_gene_ids.shape[0])))
        print('Only first occurance will be used.\n')
    first_src, first_trg = {}, {}
    for i, s in enumerate(src_gene_ids.tolist()):
        first_src.setdefault(s, i)
    for i, s in enumerate(trg_gene_ids.tolist()):
        first_trg.setdefault(s, i)
    inds1, inds2 = [], []
    for s in trg_gene_ids.tolist():
        if s in first_src:
            inds1.append(first_trg[s])
            inds2.append(first_src[s])
    assert len(inds1) > 0
    return np.asarray(inds1, dtype=int), np.asarray(inds2, dtype=int)


def unsupervised_acc_kta(X, labels, kernel='linear', param=1.0, center=True, normalize=True):
    """Kernel-target alignment of the cells' linear kernel with the label kernel (scRNA/utils.py:134-152; the score
    cmd_target.py:254-260 maximises over the mixtures).  The linear kernel X^T X is the device Gram
    (scrna_pairwise_f64), centring / normalisation / the three Frobenius products are two O(n^2) device passes
    (csrc/eval_kta.cu) instead of the reference's dense n^3 products.

    `kernel='rbf'` (get_kernel, utils.py:108-124): exp(-|x_i - x_j|^2 / param) is formed in place from the same
    Gram.  A labeling with a single cluster makes the centred label kernel vanish and the reference returns
    rounding noise; 0.0 is returned here."""
    if kernel not in ('linear', 'rbf'):
        raise ValueError("kernel must be 'linear' or 'rbf'")
    from . import _lib
    from .sc3_device import default_ops
    ops = default_ops()
    t = ops.torch
    X = np.asarray(X, dtype=np.float64)
    labels = np.asarray(labels).astype(np.int64).reshape(-1)
    n = X.shape[1]
    if labels.size != n:
        raise ValueError("one label per cell expected")
    Xd = ops.to_device(X)
    K = ops.empty(n, n)
    ops._call("scrna_pairwise_f64", Xd, n, X.shape[0], n, 1, K, n, ops._s())
    rowsum, diag = ops.empty(n), ops.empty(n)
    ops._call("scrna_kta_rowstats_f64", K, n, n, rowsum, diag, ops._s())
    if kernel == 'rbf':
        ops._call("scrna_kta_rbf_f64", K, n, n, diag.clone(), float(param), ops._s())
        ops._call("scrna_kta_rowstats_f64", K, n, n, rowsum, diag, ops._s())
    rs, dg = rowsum.cpu().numpy(), diag.cpu().numpy()

    def params(rowsum_, diag_):
        """(row means, scales, total mean) of a kernel given its row sums and diagonal; None scales = the
        normalisation is numerically unstable (normalize_kernel's identity fallback, utils.py:71-73)."""
        if center:
            mean, tot = rowsum_ / float(n), rowsum_.sum() / float(n) ** 2
        else:
            mean, tot = np.zeros(n), 0.0
        d = diag_ - 2.0 * mean + tot
        if not normalize:
            return mean, np.ones(n), tot, d
        with np.errstate(all='ignore'):
            a = np.sqrt(d)
        if np.any(np.isnan(a)) or np.any(np.isinf(a)) or np.any(np.abs(a) <= 1e-16):
            return mean, None, tot, d
        return mean, 1.0 / a, tot, d

    counts = np.bincount(labels, minlength=int(labels.max()) + 1).astype(np.float64)
    mx, bx, tx, dx = params(rs, dg)
    my, by, ty, dy = params(counts[labels], np.ones(n))
    dev = lambda a: t.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(ops.dev)

    def frobenius(bx_, by_):
        part = ops.empty(3 * _lib.call("scrna_kta_blocks"))
        out = ops.empty(3)
        lab32 = t.from_numpy(labels.astype(np.int32)).to(ops.dev)
        ops._call("scrna_kta_align_f64", K, n, n, lab32, dev(mx), dev(bx_), float(tx), dev(my), dev(by_), float(ty),
                  part, out, ops._s(), launches=2)
        return out.cpu().numpy()          # <Kx, Ky>, <Kx, Kx>, <Ky, Ky>

    if bx is None and by is None:
        # normalize_kernel's identity fallback on both sides (utils.py:71-73): only the diagonals survive
        den = np.sqrt((dx * dx).sum() * (dy * dy).sum())
        return float((dx * dy).sum() / den) if den > 0 else 0.0
    if bx is None or by is None:
        # one side fell back to diag(d): <diag(d), Kn> = sum d_i (the normalised side has a unit diagonal), but
        # <Kn, Kn> is the FULL Frobenius norm of the normalised side -- from the device pass
        ones = np.ones(n)
        _, sxx, syy = frobenius(ones if bx is None else bx, ones if by is None else by)
        d_fb, s_ok = (dx, syy) if bx is None else (dy, sxx)
        den = np.sqrt((d_fb * d_fb).sum() * s_ok)
        return float(d_fb.sum() / den) if den > 0 else 0.0
    sxy, sxx, syy = frobenius(bx, by)
    den = np.sqrt(sxx * syy)
    return float(sxy / den) if den > 0 else 0.0


# ---- small host helpers of the reference's evaluation code (scRNA/utils.py:66-124), kept for API completeness:
# O(n^2) element-wise / closed forms instead of the reference's dense n^3 products; unused by the hot path ----------
def normalize_kernel(K):
    N = K.shape[0]
    a = np.sqrt(np.diag(K)).reshape((N, 1))
    if np.any(np.isnan(a)) or np.any(np.isinf(a)) or np.any(np.abs(a) <= 1e-16):
        print('Numerical instabilities.')
        C = np.eye(N)
    else:
        b = 1. / a
        C = b.dot(b.T)
    return K * C


def center_kernel(K):
    """K - 1K/N - K1/N + 1K1/N^2 via row / column means (the reference forms three N^3 products, utils.py:79-83)."""
    K = np.asarray(K, dtype=np.float64)
    return K - K.mean(axis=0, keepdims=True) - K.mean(axis=1, keepdims=True) + K.mean()


def kta_align_general(K1, K2):
    """<K1, K2>_F / sqrt(<K1, K1>_F <K2, K2>_F)  (utils.py:86-95, traces of products there)."""
    return float(np.sum(K1 * K2) / np.sqrt(np.sum(K1 * K1) * np.sum(K2 * K2)))


def kta_align_binary(K, y):
    """Alignment of K with the label kernel y y^T, y in {+1, -1}^m (utils.py:98-103)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    m = y.size
    return float(y.dot(K).dot(y) / (m * np.sqrt(np.sum(K * K))))


def get_kernel(X, Y, type='linear', param=1.0):
    """Kernel between the columns of X and Y (dims x examples), utils.py:106-124."""
    if type == 'linear':
        return X.T.dot(Y)
    if type == 'rbf':
        dx = np.sum(X * X, axis=0).reshape(-1, 1)
        dy = np.sum(Y * Y, axis=0).reshape(1, -1)
        return np.exp(-(dx - 2. * X.T.dot(Y) + dy) / param)
    return 1.0


def get_transferability_score(W, H, trg_data, reps=100, alpha=0.0, l1=0.75, max_iter=100, rel_err=1e-3):
    """Transferability of a source dictionary to a target data set (scRNA/utils.py:155-184; row f4 of SURVEY.md §8):
    `reps` target solves with the dictionary's gene rows permuted (the "no transfer" error distribution), the
    un-permuted solve, an NMF of the target itself (the best achievable error) and the error of the given (W, H).
    Returns (score, percs, p_value) like the reference.

    The target is uploaded once; every solve is the device path of `get_transferred_data_matrix` (K5-K7) and
    draws, in the reference's order, `np.random.permutation(genes)` then `np.random.randn(k, n_t)` from the global
    legacy RNG; the target NMF is `NmfClustering`'s solver with tol=1e-5 (its NNDSVDar init uses its own
    RandomState(0), as in scikit-learn)."""
    from .kernels import default_kernels
    from .nmf_clustering import _fit_nmf
    from .nmf_solver import DeviceMatrix
    W = np.ascontiguousarray(W, dtype=np.float64)
    trg_data = np.asarray(trg_data, dtype=np.float64)
    dm = DeviceMatrix.from_host(trg_data, default_kernels())
    errs = np.zeros((reps,))
    for i in range(errs.size):
        rand_gene_inds = np.random.permutation(W.shape[0])
        _, _, _, errs[i] = get_transferred_data_matrix(W[rand_gene_inds, :], dm, max_iter=max_iter, rel_err=rel_err)
    sess = TransferSession(W, dm)
    _, _, _, err_nonpermuted = get_transferred_data_matrix(W, dm, max_iter=max_iter, rel_err=rel_err, _session=sess)

    res = _fit_nmf(trg_data, W.shape[1], alpha, l1, max_iter, 0.00001, 'cd', 'host')
    best = TransferSession(res.G, dm)
    best._set_H(res.C)
    err_best = best.l1_error()                      # sum|X - W_best.H_best| / size
    sess._set_H(np.asarray(H, dtype=np.float64))
    err_curr = sess.l1_error()                      # sum|X - W.H| / size
    err_worst = np.max(errs)

    errs[errs < err_best] = err_best
    percs = 1.0 - (errs - err_best) / (err_worst - err_best)
    score = 1.0 - np.max([err_curr - err_best, 0]) / (err_worst - err_best)
    p_value = sum(errs < err_nonpermuted) / reps
    return score, percs, p_value
