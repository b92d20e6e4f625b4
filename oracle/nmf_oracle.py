"""CPU (numpy, float64) restatement of the reference's NMF hot path.

TEST INFRASTRUCTURE ONLY -- the oracle is the checker, never the product.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import it.
Nothing under scrna_b200/ imports anything from oracle/.

What is restated, and where it comes from (paths relative to /root/reference unless prefixed
SP/ = site-packages of scikit-learn 1.9.0 / scipy 1.18.1 -- the reference delegates its
arithmetic to those un-vendored, un-pinned libraries):

  cd_sweep                <- SP/sklearn/decomposition/_cdnmf_fast.pyx:8-38  (_update_cdnmf_fast)
  cd_half_step            <- SP/sklearn/decomposition/_nmf.py:369-396       (_update_coordinate_descent)
  nmf_cd                  <- SP/sklearn/decomposition/_nmf.py:399-518       (_fit_coordinate_descent)
  mu_update_w/h, nmf_mu   <- SP/sklearn/decomposition/_nmf.py:521-626,629-723,726-888 (Frobenius)
  era_regularization      <- scRNA/nmf_clustering.py:25-29 `alpha=`/`l1_ratio=` with the scikit-learn
                             0.20-0.23 semantics (unscaled)
  nndsvdar_init           <- SP/sklearn/decomposition/_nmf.py:214-366 (called, not restated: a4 phase 1)
  nmf_clustering_apply    <- scRNA/nmf_clustering.py:20-42
  nmf_clustering_initw_apply <- scRNA/nmf_clustering.py:59-85
  transfer_h_init         <- scRNA/utils.py:189-193
  transfer_mu             <- scRNA/utils.py:187-230   (get_transferred_data_matrix, normalize_H2=False)
  mixed_data              <- scRNA/nmf_clustering.py:169-200
  calc_rejection          <- scRNA/nmf_clustering.py:225-272
  matching_gene_inds      <- scRNA/utils.py:233-278

Pinning: the reference's own tests pin nothing at this boundary ("parity unpinned"
by the reference).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container under oracle/refshim.py: tests/golden/*.npz (made by tests/golden/make_golden.py),
checked by tests/test_oracle_nmf.py.
"""
import numpy as np

EPSILON = float(np.finfo(np.float32).eps)  # SP/sklearn/decomposition/_nmf.py:32


# --------------------------------------------------------------------------------------------
# coordinate descent (the reference's solver)
# --------------------------------------------------------------------------------------------
def cd_sweep(W, HHt, XHt, perm):
    """One pass of _update_cdnmf_fast over all rows of W (in place); returns the violation.

    Row-parallel restatement: rows are independent, components are visited in
    `perm` order, the inner product over r runs in index order, exactly as the Cython loop.
    """
    n, k = W.shape
    violation = 0.0
    for t in perm:
        t = int(t)
        grad = -XHt[:, t].copy()
        for r in range(k):
            grad += HHt[t, r] * W[:, r]
        pg = np.where(W[:, t] == 0.0, np.minimum(0.0, grad), grad)
        violation += float(np.abs(pg).sum())
        hess = HHt[t, t]
        if hess != 0.0:
            W[:, t] = np.maximum(W[:, t] - grad / hess, 0.0)
    return violation


def cd_sweep_rowmajor(W, HHt, XHt, perm):
    """Literal (slow, pure Python) transcription of the Cython loop order: rows outer, components
    inner.  Used on small cases to show cd_sweep() and the violation accumulation order agree."""
    n, k = W.shape
    violation = 0.0
    for i in range(n):
        for t in perm:
            t = int(t)
            grad = -XHt[i, t]
            for r in range(k):
                grad += HHt[t, r] * W[i, r]
            pg = min(0.0, grad) if W[i, t] == 0 else grad
            violation += abs(pg)
            hess = HHt[t, t]
            if hess != 0:
                W[i, t] = max(W[i, t] - grad / hess, 0.0)
    return violation


def cd_half_step(X, W, Ht, l1_reg, l2_reg, perm):
    """_update_coordinate_descent: Gram, product, -l1, +l2 on the diagonal, sweep."""
    k = Ht.shape[1]
    HHt = Ht.T @ Ht
    XHt = X @ Ht
    if l2_reg != 0.0:
        HHt.flat[:: k + 1] += l2_reg
    if l1_reg != 0.0:
        XHt -= l1_reg
    return cd_sweep(W, HHt, XHt, perm)


def nmf_cd(X, W0, H0, l1_W=0.0, l1_H=0.0, l2_W=0.0, l2_H=0.0, tol=1e-4, max_iter=200,
           update_H=True, shuffle=True, seed=0, perms=None, trace=None):
    """_fit_coordinate_descent.  X: rows x cols, W0: rows x k, H0: k x cols.

    The permutation stream comes from ONE RandomState(seed): W-perm1, H-perm1, W-perm2, ...
    (one draw per iteration when update_H=False).  `perms` overrides it (list of arrays).
    Returns (W, H, n_iter, violations) with violations[i] = violation of iteration i+1.
    """
    X = np.asarray(X, dtype=np.float64)
    W = np.ascontiguousarray(W0, dtype=np.float64).copy()
    Ht = np.ascontiguousarray(np.asarray(H0, dtype=np.float64).T).copy()
    k = W.shape[1]
    rng = np.random.RandomState(seed)
    pi = 0

    def next_perm():
        nonlocal pi
        if perms is not None:
            p = np.asarray(perms[pi], dtype=np.intp)
            pi += 1
            return p
        return rng.permutation(k) if shuffle else np.arange(k)

    violations = []
    violation_init = None
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        violation = cd_half_step(X, W, Ht, l1_W, l2_W, next_perm())
        if update_H:
            violation += cd_half_step(X.T, Ht, W, l1_H, l2_H, next_perm())
        violations.append(violation)
        if trace is not None:
            trace.append((W.copy(), Ht.T.copy()))
        if n_iter == 1:
            violation_init = violation
        if violation_init == 0:
            break
        if violation / violation_init <= tol:
            break
    return W, Ht.T.copy(), n_iter, np.asarray(violations)


def draw_cd_perms(k, n_draws, seed=0, shuffle=True):
    """The host-side pre-drawn permutation stream: what RandomState(seed) yields
    for `n_draws` successive permutation(k) calls."""
    rng = np.random.RandomState(seed)
    if not shuffle:
        return np.tile(np.arange(k, dtype=np.int32), (n_draws, 1))
    return np.stack([rng.permutation(k) for _ in range(n_draws)]).astype(np.int32)


# --------------------------------------------------------------------------------------------
# multiplicative update (sklearn solver='mu', Frobenius) -- the BASELINE.json headline metric
# --------------------------------------------------------------------------------------------
def mu_update_w(X, W, H, l1_reg, l2_reg):
    XHt = X @ H.T
    HHt = H @ H.T
    den = W @ HHt
    if l1_reg > 0:
        den += l1_reg
    if l2_reg > 0:
        den = den + l2_reg * W
    den[den == 0] = EPSILON
    W *= XHt / den
    return W


def mu_update_h(X, W, H, l1_reg, l2_reg):
    WtX = W.T @ X
    den = (W.T @ W) @ H
    if l1_reg > 0:
        den += l1_reg
    if l2_reg > 0:
        den = den + l2_reg * H
    den[den == 0] = EPSILON
    H *= WtX / den
    return H


def frobenius_error(X, W, H):
    """_beta_divergence(X, W, H, 2, square_root=True) == ||X - WH||_F."""
    R = X - W @ H
    return float(np.sqrt(max(np.vdot(R, R), 0.0)))


def nmf_mu(X, W0, H0, l1_W=0.0, l1_H=0.0, l2_W=0.0, l2_H=0.0, tol=1e-4, max_iter=200,
           update_H=True, trace=None):
    """_fit_multiplicative_update (beta_loss=2, gamma=1).  Returns (W, H, n_iter, errors) with
    errors = [(n_iter, ||X-WH||_F)] at the iterations where the convergence test ran."""
    X = np.asarray(X, dtype=np.float64)
    W = np.array(W0, dtype=np.float64, copy=True)
    H = np.array(H0, dtype=np.float64, copy=True)
    error_at_init = frobenius_error(X, W, H)
    previous_error = error_at_init
    errors = [(0, error_at_init)]
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        W = mu_update_w(X, W, H, l1_W, l2_W)
        if update_H:
            H = mu_update_h(X, W, H, l1_H, l2_H)
        if trace is not None:
            trace.append((W.copy(), H.copy()))
        if tol > 0 and n_iter % 10 == 0:
            error = frobenius_error(X, W, H)
            errors.append((n_iter, error))
            if (previous_error - error) / error_at_init < tol:
                break
            previous_error = error
    return W, H, n_iter, errors


# --------------------------------------------------------------------------------------------
# regularisation + init as the reference requests them
# --------------------------------------------------------------------------------------------
def era_regularization(alpha, l1_ratio):
    """decomp.NMF(alpha=, l1_ratio=) of the scikit-learn contemporary with the reference
    (regularization='both', unscaled): returns (l1_W, l1_H, l2_W, l2_H)."""
    l1 = float(alpha) * float(l1_ratio)
    l2 = float(alpha) * (1.0 - float(l1_ratio))
    return l1, l1, l2, l2


def nndsvdar_init(X, k, seed=0):
    """NMF(init='nndsvdar', random_state=0) initialisation.  Calls the installed scikit-learn routine
    (randomized SVD + NNDSVD + random fill); SURVEY.md §8 row a4 phase 1 injects this init."""
    from sklearn.decomposition._nmf import _initialize_nmf
    W0, H0 = _initialize_nmf(np.asarray(X, dtype=np.float64), k, init="nndsvdar", random_state=seed)
    return W0, H0


def nmf_clustering_apply(X, k, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3, init=None):
    """NmfClustering.apply after pre-processing (nmf_clustering.py:20-42): returns
    (dictionary W g x k, data_matrix H k x n, labels, n_iter)."""
    X = np.asarray(X, dtype=np.float64)
    W0, H0 = nndsvdar_init(X, k) if init is None else init
    l1W, l1H, l2W, l2H = era_regularization(alpha, l1)
    W, H, n_iter, _ = nmf_cd(X, W0, H0, l1W, l1H, l2W, l2H, tol=rel_err, max_iter=max_iter,
                             update_H=True, shuffle=True, seed=0)
    labels = np.argmax(H, axis=0)
    return W, H, labels, n_iter


def one_hot_labels(labels):
    """pd.get_dummies(labels): columns ordered by sorted unique label (nmf_clustering.py:64)."""
    labels = np.asarray(labels)
    uniq = np.unique(labels)
    return (labels[:, None] == uniq[None, :]).astype(np.float64), uniq


def nmf_clustering_initw_apply(X, labels, k, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3):
    """NmfClustering_initW.apply after pre-processing (nmf_clustering.py:59-85).

    (1) CD with H fixed = one-hot^T (k x n), W (g x k) from ZEROS, no regularisation
        (SP/sklearn/decomposition/_nmf.py:1224-1228; era `regularization=None`).
    (2) CD NMF on X^T (n x g) from (one-hot n x k, dictionary^T k x g) with regularisation.
    Returns (dictionary g x k, data_matrix k x n, labels, (n_iter1, n_iter2)).
    """
    X = np.asarray(X, dtype=np.float64)
    onehot, uniq = one_hot_labels(labels)
    assert onehot.shape[1] == k, "k must equal the number of unique labels"
    g, n = X.shape
    Wz = np.zeros((g, k))
    learned, _, n1, _ = nmf_cd(X, Wz, onehot.T, 0, 0, 0, 0, tol=rel_err, max_iter=max_iter,
                               update_H=False, shuffle=True, seed=0)
    l1W, l1H, l2W, l2H = era_regularization(alpha, l1)
    W, H, n2, _ = nmf_cd(X.T, onehot, learned.T, l1W, l1H, l2W, l2H, tol=rel_err,
                         max_iter=max_iter, update_H=True, shuffle=True, seed=0)
    out_labels = np.argmax(W, axis=1)
    return H.T.copy(), W.T.copy(), out_labels, (n1, n2)


# --------------------------------------------------------------------------------------------
# target transfer (the reference's only multiplicative-update loop)
# --------------------------------------------------------------------------------------------
def transfer_h_init(randn_draw):
    """utils.py:189-193 applied to a pre-drawn np.random.randn(k, n_t) block."""
    H = np.array(randn_draw, dtype=np.float64, copy=True)
    H[H < 0.0] *= -1.0
    H[H < 1e-10] = 1e-10
    return H


def transfer_mu(W, X, H0, max_iter=100, rel_err=1e-3, trace=None):
    """get_transferred_data_matrix(W, trg_data) with injected H0 (utils.py:187-230).

    Literal association order: W.T.dot(W.dot(H)).  Returns (H, H2, err, n_iter)."""
    W = np.asarray(W, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    H = np.array(H0, dtype=np.float64, copy=True)
    n_iter = 0
    err = 1e10
    new_err = None
    while n_iter < max_iter:
        n_iter += 1
        den = W.T.dot(W.dot(H))
        if np.any(den == 0.0):
            raise Exception('DA target: division by zero.')
        H *= W.T.dot(X) / den
        new_err = np.sum(np.abs(X - W.dot(H))) / float(X.size)
        if trace is not None:
            trace.append((H.copy(), new_err))
        if np.abs((err - new_err) / err) <= rel_err and err >= new_err:
            break
        err = new_err
    H2 = np.zeros_like(H)
    H2[(np.argmax(H, axis=0), np.arange(X.shape[1]))] = 1
    return H, H2, new_err, n_iter


def mixed_data(W, H, H2, X, mix, use_H2=True):
    """nmf_clustering.py:185-200 (reject_ratio == 0): returns (mixed, new_trg)."""
    new_trg = W.dot(H2) if use_H2 else W.dot(H)
    return mix * new_trg + (1.0 - mix) * X, new_trg


def calc_rejection(X, W, H, H2, cluster_labels, src_num_cluster):
    """DaNmfClustering.calc_rejection (nmf_clustering.py:225-272), quirks kept (operator precedence
    at :232)."""
    import scipy.stats as stats
    diffs = np.zeros(H2.shape[1])
    for c in range(src_num_cluster):
        inds = np.where(cluster_labels == c)[0]
        if inds.size > 0:
            min_h2 = np.min(H[:, inds])
            max_h2 = np.max(H[:, inds])
            foo = H[:, inds] - min_h2 / (max_h2 - min_h2)
            foo = np.max(foo, axis=0) - np.min(foo, axis=0)
            diffs[inds] = foo
    sum_expr = np.sum(X, axis=0)
    sum_expr -= np.min(sum_expr)
    sum_expr /= np.max(sum_expr)
    sum_expr = sum_expr + 1.0
    sum_expr /= np.max(sum_expr)
    weight = 1. - sum_expr
    reconstr_err = np.sum(np.abs(X - W.dot(H2)), axis=0)
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
    reject.append(('Dist L2 H', -np.sum((np.abs(X - W.dot(H)) ** 2.), axis=0)))
    reject.append(('Dist L2 H2', -np.sum((np.abs(X - W.dot(H2)) ** 2.), axis=0)))
    reject.append(('Dist L1 H', -np.sum(np.abs(X - W.dot(H)), axis=0)))
    reject.append(('Dist L1 H2', -np.sum(np.abs(X - W.dot(H2)), axis=0)))
    return reject


def matching_gene_inds(src_gene_ids, trg_gene_ids):
    """get_matching_gene_inds (utils.py:233-278): order-preserving intersection in TARGET order,
    first occurrence wins.  Returns (inds1 into target ids, inds2 into source ids)."""
    src_gene_ids = np.asarray(src_gene_ids)
    trg_gene_ids = np.asarray(trg_gene_ids)
    first_src = {}
    for i, s in enumerate(src_gene_ids.tolist()):
        first_src.setdefault(s, i)
    first_trg = {}
    for i, s in enumerate(trg_gene_ids.tolist()):
        first_trg.setdefault(s, i)
    inds1, inds2 = [], []
    for s in trg_gene_ids.tolist():          # duplicates in the target are listed repeatedly,
        if s in first_src:                   # each mapping to the first occurrence (utils.py:251-254)
            inds1.append(first_trg[s])
            inds2.append(first_src[s])
    assert len(inds1) > 0
    return np.asarray(inds1, dtype=int), np.asarray(inds2, dtype=int)
