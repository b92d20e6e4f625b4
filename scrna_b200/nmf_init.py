"""NNDSVDar initialisation with every pass over X on the device.

`NmfClustering.apply` starts from `sklearn.decomposition.NMF(init='nndsvdar', random_state=0)`
(scRNA/nmf_clustering.py:25-29), i.e. `_initialize_nmf` (SP/sklearn/decomposition/_nmf.py:214-366): a
randomized SVD (SP/sklearn/utils/extmath.py `_randomized_range_finder` / `_randomized_svd`: Gaussian test
matrix, 7 (or 4) LU-normalised power iterations, QR, SVD of the thin projection, sign flip), the NNDSVD split
of the singular vectors and a random fill of the zeros.  On the host that is ~16 sweeps of the 16 GB fp64
matrix at 20k x 100k; here the 16 products `X @ Q` / `X^T @ Q` (Q is (k+10) wide) run on the same X sweeps as
the NMF iterations (tcgen05 or FP32 pipe), and the host keeps only the O((g+n) k^2) tall-skinny factorizations
(LU / QR / SVD of (g or n) x (k+10) matrices) and the reference's RNG stream (RandomState(seed): the test
matrix, then the zero fills -- drawn in sklearn's order).
"""
import numpy as np
from scipy import linalg

EPS = 1e-6  # _initialize_nmf's zero threshold (_nmf.py:346-347)


def _svd_flip_v(u, v):
    """svd_flip(u, v, u_based_decision=False): signs from the largest |entry| of each row of v."""
    idx = np.argmax(np.abs(v), axis=1)
    signs = np.sign(v[np.arange(v.shape[0]), idx])
    return u * signs[np.newaxis, :], v * signs[:, np.newaxis]


def _svd_flip_u(u, v):
    idx = np.argmax(np.abs(u), axis=0)
    signs = np.sign(u[idx, np.arange(u.shape[1])])
    return u * signs[np.newaxis, :], v * signs[:, np.newaxis]


def _chunked(times, width):
    """`times` handles operands of at most `width` columns (the solver's rank): wider ones go in column chunks."""
    def f(Q):
        if Q.shape[1] == width:
            return times(Q)
        out = []
        for c in range(0, Q.shape[1], width):
            blk = np.zeros((Q.shape[0], width))
            w = min(width, Q.shape[1] - c)
            blk[:, :w] = Q[:, c: c + w]
            out.append(times(blk)[:, :w])
        return np.concatenate(out, axis=1)
    return f


def randomized_svd_device(prod, k, seed=0, n_oversamples=10):
    """sklearn `_randomized_svd(X, k, random_state=seed)` with `prod.x_times` / `prod.xt_times` as the only
    accesses to X (g x n).  Returns U (g x k), S (k), Vt (k x n)."""
    g = prod.X.g
    n = prod.cell_range[2] if getattr(prod, "cell_range", None) else prod.X.n
    rng = np.random.RandomState(seed)
    r = k + n_oversamples
    n_iter = 7 if k < 0.1 * min(g, n) else 4
    transpose = g < n                      # sklearn works on M = X^T when X has more columns than rows
    xt = getattr(prod, "xt_times_full", prod.xt_times)      # global products when the cells are sharded
    width = getattr(prod, "k", k + n_oversamples)          # the solver's rank; wider operands go in column chunks
    x_times = _chunked(prod.x_times, width)
    xt_times = _chunked(xt, width)
    if transpose:
        m_times, mt_times, m_cols = xt_times, x_times, g
    else:
        m_times, mt_times, m_cols = x_times, xt_times, n
    Q = rng.normal(size=(m_cols, r))
    lu = lambda A: linalg.lu(A, permute_l=True, check_finite=False)[0]
    if n_iter <= 2:
        lu = lambda A: A
    for _ in range(n_iter):
        Q = lu(m_times(Q))
        Q = lu(mt_times(Q))
    Q, _ = linalg.qr(m_times(Q), mode="economic", check_finite=False)
    B = mt_times(Q).T                       # Q^T M  (r x m_cols)
    Uhat, s, Vt = linalg.svd(B, full_matrices=False, lapack_driver="gesdd")
    U = Q @ Uhat
    if not transpose:
        U, Vt = _svd_flip_u(U, Vt)
        return U[:, :k], s[:k], Vt[:k, :]
    U, Vt = _svd_flip_v(U, Vt)
    return Vt[:k, :].T, s[:k], U[:, :k].T


def nndsvdar_from_svd(U, S, V, xmean, seed=0):
    """The NNDSVD split and the 'ar' random fill of `_initialize_nmf` (_nmf.py:313-359)."""
    k = S.shape[0]
    W = np.zeros_like(U)
    H = np.zeros_like(V)
    W[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])
    H[0, :] = np.sqrt(S[0]) * np.abs(V[0, :])
    for j in range(1, k):
        x, y = U[:, j], V[j, :]
        x_p, y_p = np.maximum(x, 0), np.maximum(y, 0)
        x_n, y_n = np.abs(np.minimum(x, 0)), np.abs(np.minimum(y, 0))
        x_p_nrm, y_p_nrm = linalg.norm(x_p), linalg.norm(y_p)
        x_n_nrm, y_n_nrm = linalg.norm(x_n), linalg.norm(y_n)
        m_p, m_n = x_p_nrm * y_p_nrm, x_n_nrm * y_n_nrm
        if m_p > m_n:
            u, v, sigma = x_p / x_p_nrm, y_p / y_p_nrm, m_p
        else:
            u, v, sigma = x_n / x_n_nrm, y_n / y_n_nrm, m_n
        lbd = np.sqrt(S[j] * sigma)
        W[:, j] = lbd * u
        H[j, :] = lbd * v
    W[W < EPS] = 0
    H[H < EPS] = 0
    rng = np.random.RandomState(seed)
    W[W == 0] = abs(xmean * rng.standard_normal(size=len(W[W == 0])) / 100)
    H[H == 0] = abs(xmean * rng.standard_normal(size=len(H[H == 0])) / 100)
    return W, H


def nndsvdar_init_device(dm, k, kernels=None, seed=0, XT=None, comm=None, cell_range=None):
    """(W0 g x k, H0 k x n) of NMF(init='nndsvdar', random_state=seed) for the device-resident matrix `dm`.
    Returns (W0, H0, XT): the transposed copy is handed back so that the NMF solver can reuse it.

    With `comm` and `cell_range = (c0, c1, n_total)` the matrix is one cell shard of n_total cells: the products
    are summed / gathered over the ranks, the tall-skinny host factorizations and the RNG stream are replicated
    (identical on every rank), and the GLOBAL H0 (k x n_total) is returned -- the caller takes its columns.
    Ranks above 54 (k + 10 oversamples > 64 components) run the (k + 10)-wide products in chunks of 64 columns."""
    from .nmf_solver import NmfSolver
    r = min(k + 10, 64)
    prod = NmfSolver(dm, r, kernels=kernels, XT=XT, comm=comm)
    n_total = dm.n
    if cell_range is not None:
        prod.set_cell_range(*cell_range)
        n_total = cell_range[2]
    U, S, Vt = randomized_svd_device(prod, k, seed=seed)
    xmean = prod.x_sum() / (float(dm.g) * float(n_total))
    W0, H0 = nndsvdar_from_svd(U, S, Vt, xmean, seed=seed)
    return W0, H0, prod.XT
