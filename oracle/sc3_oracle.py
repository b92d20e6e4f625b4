"""CPU (numpy / scipy / scikit-learn, float64) restatement of the reference's SC3 hot path.

TEST INFRASTRUCTURE ONLY -- the checker, never the product (see oracle/nmf_oracle.py header).

The reference's SC3 stages are thin wrappers over scipy / scikit-learn calls; the
restatement keeps those library calls where the reference has them and restates the surrounding
arithmetic, each function citing the reference lines it follows (paths under /root/reference):

  distances               <- scRNA/sc3_clustering_impl.py:106-132
  transformations         <- scRNA/sc3_clustering_impl.py:135-203   (returns only the vectors' block
                             the caller uses plus the discarded matrix)
  kmeans_draws / kmeans   <- scRNA/sc3_clustering_impl.py:206-219 -> sklearn KMeans(n_init=10,
                             max_iter=10000, init='k-means++') with the global-RNG stream made explicit
                             (SP/sklearn/cluster/_kmeans.py:180-279, 630-758, 1436-1562)
  build_consensus_matrix  <- scRNA/sc3_clustering_impl.py:222-241
  consensus_clustering    <- scRNA/sc3_clustering_impl.py:244-260
  sc3_apply               <- scRNA/sc3_clustering.py:58-109

Pinning: tests/golden/sc3.npz holds the per-stage outputs of the UNMODIFIED reference on the mixed
default target data (tests/golden/make_golden.py); tests/test_oracle_sc3.py checks this file against it.
"""
import numpy as np
import scipy.cluster.hierarchy as spc
import scipy.linalg as sl
import scipy.spatial.distance as dist
import scipy.stats as stats


def distances(data, metric='euclidean'):
    if metric == 'pearson':
        return 1. - np.corrcoef(data.T)
    if metric == 'spearman':
        X, _ = stats.spearmanr(data, axis=0)
        return 1. - X
    return dist.squareform(dist.pdist(data.T, metric=metric))


def euclidean_sequential(data):
    """pdist restated: plain sequential fp64 sum over genes, no FMA (SURVEY.md probe P8: bitwise equal
    to scipy).  O(n^2 g) in numpy broadcasting -- small cases only."""
    X = np.asarray(data, dtype=np.float64)
    g, n = X.shape
    acc = np.zeros((n, n))
    for f in range(g):
        d = X[f][:, None] - X[f][None, :]
        acc = acc + d * d
    return np.sqrt(acc)


def rank_average(data):
    """scipy.stats.rankdata(method='average') per column (spearmanr, SP/scipy/stats/_stats_py.py:5440)."""
    return np.apply_along_axis(stats.rankdata, 0, np.asarray(data, dtype=np.float64))


def pearson_from_columns(A):
    """np.corrcoef(A, rowvar=0) restated in the op order of numpy (SP/numpy/lib/_function_base_impl.py:
    2750-2766 cov, :3066-3071 corrcoef): centre, dot, scale by 1/(F-1), divide by the stds, clip."""
    A = np.asarray(A, dtype=np.float64)
    F = A.shape[0]
    Xc = A - A.mean(axis=0, keepdims=True)
    c = Xc.T @ Xc
    c *= np.true_divide(1, F - 1)
    d = np.diag(c)
    std = np.sqrt(d)
    c = c / std[:, None]
    c = c / std[None, :]
    return np.clip(c, -1, 1)


def pca_prepare(dm):
    """Column centring and nan-std scaling of sc3_clustering_impl.py:161-173."""
    dm = np.array(dm, dtype=np.float64)
    n = dm.shape[0]
    dm = dm - np.repeat(np.mean(dm, axis=0).reshape((1, n)), n, axis=0)
    np.place(dm, np.isinf(dm), np.nan)
    sd = np.nanstd(dm, axis=0)
    if not (np.sum(sd == 0) == len(sd)):
        dm = dm / np.repeat(sd.reshape((1, n)), n, axis=0)
    return dm


def laplacian_quirk(dm):
    """The reference's 'spectral' matrix (sc3_clustering_impl.py:143-150), quirks kept: degree = row
    sums of the DISTANCE matrix, L = D - A broadcasts the vector over rows."""
    dm = np.asarray(dm, dtype=np.float64)
    A = np.exp(-dm / np.max(dm))
    D = np.sum(dm, axis=1)
    L = D - A
    D1 = np.diag(D.__pow__(-0.5))
    D1[np.isinf(D)] = 0.0
    return D1.dot(L.dot(D1))


def transformations(dm, components=5, method='pca'):
    num_cells = dm.shape[0]
    m = laplacian_quirk(dm) if method == 'spectral' else pca_prepare(dm)
    _, vals, ev = sl.svd(m)
    vecs = ev.T
    vals = vals / np.sqrt(np.max((1, num_cells - 1)))
    inds = np.arange(components)
    D = np.diag(vals[inds])
    return vecs[:, inds].dot(D.dot(vecs[:, inds].T)), vecs[:, inds], vals


# ---- k-means with the RNG stream made explicit --------------------------------------------------
def kmeans_draw_count(k):
    """Uniform doubles one KMeans init consumes from the legacy RandomState:
    choice(n, p) = 1 draw, then (k-1) x uniform(size=2+int(log k))."""
    trials = 2 + int(np.log(k))
    return 1 + (k - 1) * trials, trials


def kmeans_pp(X, k, x_sq, draws):
    """_kmeans_plusplus with pre-drawn uniforms (SP/sklearn/cluster/_kmeans.py:180-279)."""
    n = X.shape[0]
    _, trials = kmeans_draw_count(k)
    p = np.full(n, 1.0 / n)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    first = int(cdf.searchsorted(draws[0], side='right'))   # RandomState.choice(n, p=uniform)
    idx = [first]
    closest = np.maximum(x_sq + x_sq[first] - 2.0 * (X @ X[first]), 0.0)
    pot = closest.sum()
    pos = 1
    for c in range(1, k):
        rand_vals = draws[pos:pos + trials] * pot
        pos += trials
        cand = np.searchsorted(np.cumsum(closest), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        d = np.maximum(x_sq[None, :] + x_sq[cand][:, None] - 2.0 * (X[cand] @ X.T), 0.0)
        np.minimum(closest, d, out=d)
        pots = d.sum(axis=1)
        best = int(np.argmin(pots))
        pot = pots[best]
        closest = d[best]
        idx.append(int(cand[best]))
    return X[idx].copy(), np.asarray(idx)


def lloyd(X, centers, tol, max_iter):
    """_kmeans_single_lloyd (SP/sklearn/cluster/_kmeans.py:630-758) incl. empty-cluster relocation
    (_k_means_common.pyx:167-212).  Returns (labels, inertia, centers, n_iter)."""
    n, d = X.shape
    k = centers.shape[0]
    centers = centers.copy()
    labels_old = np.full(n, -1, dtype=np.int32)
    labels = labels_old.copy()
    strict = False
    it = 0
    for it in range(max_iter):
        D = (centers * centers).sum(1)[None, :] - 2.0 * (X @ centers.T)
        labels = np.argmin(D, axis=1).astype(np.int32)
        w = np.bincount(labels, minlength=k).astype(np.float64)
        new = np.zeros_like(centers)
        np.add.at(new, labels, X)
        empty = np.where(w == 0)[0]
        if empty.size:
            distances = ((X - centers[labels]) ** 2).sum(axis=1)
            far = np.argpartition(distances, -empty.size)[:-empty.size - 1:-1]
            if np.max(distances) != 0:
                for e, f in zip(empty, far):
                    old = labels[f]
                    new[old] -= X[f]
                    new[e] = X[f]
                    w[e] = 1.0
                    w[old] -= 1.0
        amax = int(np.argmax(w))
        for j in range(k):
            if w[j] > 0:
                new[j] *= 1.0 / w[j]
            else:
                new[j] = new[amax]
        shift = np.sqrt(((new - centers) ** 2).sum(axis=1))
        centers = new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol:
            break
        labels_old = labels.copy()
    if not strict:
        D = (centers * centers).sum(1)[None, :] - 2.0 * (X @ centers.T)
        labels = np.argmin(D, axis=1).astype(np.int32)
    inertia = float((((X - centers[labels]) ** 2).sum(axis=1)).sum())
    return labels, inertia, centers, it + 1


def same_clustering(l1, l2, k):
    mapping = np.full(k, -1, dtype=np.int64)
    for a, b in zip(l1, l2):
        if mapping[a] == -1:
            mapping[a] = b
        elif mapping[a] != b:
            return False
    return True


def kmeans(V, k, draws, n_init=10, max_iter=10000):
    """KMeans(n_clusters=k, n_init=10, max_iter=10000).fit_predict(V) with `draws` = the n_init *
    kmeans_draw_count(k) uniforms the global RandomState would have produced."""
    X = np.array(V, dtype=np.float64)
    X = X - X.mean(axis=0)
    tol = np.mean(np.var(X, axis=0)) * 1e-4
    x_sq = (X * X).sum(axis=1)
    per, _ = kmeans_draw_count(k)
    best = None
    for i in range(n_init):
        c0, _ = kmeans_pp(X, k, x_sq, draws[i * per:(i + 1) * per])
        labels, inertia, centers, n_iter = lloyd(X, c0, tol, max_iter)
        if best is None or (inertia < best[1] and not same_clustering(labels, best[0], k)):
            best = (labels, inertia)
    return best[0]


def build_consensus_matrix(labels):
    labels = np.asarray(labels)
    if labels.ndim == 1:
        labels = labels[None, :]
    c = np.zeros((labels.shape[1], labels.shape[1]))
    for row in labels:
        c += (row[:, None] == row[None, :]).astype(np.float64)
    return c / float(labels.shape[0])


def consensus_clustering(consensus, n_components=5):
    cdm = dist.pdist(consensus)
    hclust = spc.complete(cdm)
    labels = spc.cut_tree(hclust, n_clusters=n_components).reshape(consensus.shape[0])
    return labels, dist.squareform(cdm), hclust


def nn_chain_complete(D):
    """scipy.cluster._hierarchy.nn_chain for method='complete' restated: returns the
    merges in NN-chain order as (x, y, dist) with x < y slot indices."""
    n = D.shape[0]
    D = np.array(D, dtype=np.float64)
    size = np.ones(n, dtype=np.int64)
    chain = []
    merges = []
    for _ in range(n - 1):
        if not chain:
            chain.append(int(np.nonzero(size > 0)[0][0]))
        while True:
            x = chain[-1]
            if len(chain) > 1:
                y = chain[-2]
                cur = D[x, y]
            else:
                y = -1
                cur = np.inf
            for i in range(n):
                if size[i] == 0 or i == x:
                    continue
                if D[x, i] < cur:
                    cur = D[x, i]
                    y = i
            if len(chain) > 1 and y == chain[-2]:
                break
            chain.append(y)
        chain.pop()
        chain.pop()
        if x > y:
            x, y = y, x
        merges.append((x, y, cur))
        nx, ny = size[x], size[y]
        size[x] = 0
        size[y] = nx + ny
        for i in range(n):
            if size[i] == 0 or i == y:
                continue
            v = max(D[i, x], D[i, y])
            D[i, y] = D[y, i] = v
    return merges


def sc3_apply(X, dists, transfs, k, pc_range, sub_sample=True):
    """SC3Clustering.apply (sc3_clustering.py:58-109) with consensus_mode 0 after pre-processing, drawing
    from the GLOBAL numpy RNG exactly where the reference does.  Returns (labels, consensus, label list)."""
    n = X.shape[1]
    D = [distances(X, m) for m in dists]
    V = []
    comps = pc_range[1]
    for d in D:
        for t in transfs:
            V.append(transformations(d, components=comps, method=t)[1])
    range_inds = list(range(pc_range[0], pc_range[1] + 1))
    if sub_sample and len(range_inds) > 15:
        range_inds = np.random.permutation(range_inds)[:15]
    per, _ = kmeans_draw_count(k)
    consensus = np.zeros((n, n))
    cnt = 0.
    all_labels = []
    for v in V:
        for dsel in range_inds:
            draws = np.random.random_sample(10 * per)
            lab = kmeans(v[:, :dsel], k, draws)
            all_labels.append(lab)
            consensus += build_consensus_matrix(lab)
            cnt += 1.
    consensus /= cnt
    labels, _, _ = consensus_clustering(consensus, k)
    return labels, consensus, all_labels
