"""Parity of the CUDA SC3 path (K9-K21, through the C ABI) against the oracle and the reference goldens.
Bar: Euclidean distances and consensus row distances bit-exact; correlation distances 1e-12; singular vectors
1e-8 up to sign; k-means labelings, consensus matrix and final labels identical."""
from functools import partial

import numpy as np
import pytest

from oracle import sc3_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(kernels):
    from scrna_b200.sc3_device import Sc3Ops
    return Sc3Ops(kernels)


@pytest.fixture(scope="module")
def sc3g(golden):
    return golden("sc3")


@pytest.fixture(scope="module")
def mixed(golden):
    return golden("transfer")["mixed"]


def test_euclidean_bit_exact(ops, mixed, sc3g):
    D = ops.distances(mixed, "euclidean").cpu().numpy()
    np.testing.assert_array_equal(D, sc3g["D_euclidean"])


@pytest.mark.parametrize("g,n", [(37, 5), (64, 64), (130, 129), (1000, 257), (3, 70)])
def test_euclidean_bit_exact_shapes(ops, g, n):
    rng = np.random.RandomState(g + n)
    X = rng.gamma(2.0, 3.0, size=(g, n)) * (rng.rand(g, n) < 0.7)
    np.testing.assert_array_equal(ops.distances(X, "euclidean").cpu().numpy(), so.distances(X, "euclidean"))


def test_pearson_spearman(ops, mixed, sc3g):
    P = ops.distances(mixed, "pearson").cpu().numpy()
    S = ops.distances(mixed, "spearman").cpu().numpy()
    np.testing.assert_allclose(P, sc3g["D_pearson"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(S, sc3g["D_spearman"], rtol=0, atol=1e-12)


def test_rank_average_with_ties(ops):
    rng = np.random.RandomState(1)
    X = rng.randint(0, 6, size=(300, 40)).astype(np.float64)       # many ties
    S = ops.distances(X, "spearman").cpu().numpy()
    np.testing.assert_allclose(S, so.distances(X, "spearman"), rtol=0, atol=1e-12)


@pytest.mark.parametrize("kind", ["gauss", "pca_bulk"])
@pytest.mark.parametrize("n", [131, 300, 1037])
def test_eigh_solver_vs_lapack(ops, n, kind):
    """The tridiagonalisation solver (sc3_eigh.cu: Gram, Householder tridiagonalisation with DMMA trailing updates,
    bisection, inverse iteration, back-transformation) that replaces the Jacobi sweeps from 6 000 cells on, against
    LAPACK: leading singular values, orthonormality, residual, and the individual vectors wherever the gap to the
    neighbours resolves them."""
    import scipy.linalg as sl
    rng = np.random.RandomState(n + 7)
    if kind == "gauss":
        M = rng.randn(n, n)
    else:
        X = rng.poisson(2.0 + 3.0 * (rng.rand(40, 1) < 0.3) * (rng.randint(0, 4, size=(1, n)) == 1), size=(40, n))
        M = so.pca_prepare(so.distances(np.log2(X + 1.0), "euclidean"))
    c = max(2, int(np.ceil(0.07 * n)))
    V, vals = ops.right_singular_vectors(ops.to_device(M), c, method="eigh")
    _, s, vt = sl.svd(M)
    np.testing.assert_allclose(vals[:c], s[:c], rtol=0, atol=1e-9 * s[0])
    assert np.all(vals[c:] == 0)
    V = V.cpu().numpy()
    np.testing.assert_allclose(V.T @ V, np.eye(c), rtol=0, atol=1e-8)
    assert np.linalg.norm(M.T @ (M @ V) - V * (vals[:c] ** 2)) / (s[0] ** 2) < 1e-9
    ref = vt[:c].T
    sgn = np.sign((V * ref).sum(axis=0))
    gaps = np.minimum(np.abs(np.diff(s[: c + 1] ** 2)), np.r_[np.inf, np.abs(np.diff(s[:c] ** 2))]) / s[0] ** 2
    ok = gaps > 1e-6
    assert ok.sum() >= c // 2
    np.testing.assert_allclose((V * sgn)[:, ok], ref[:, ok], rtol=0, atol=1e-7)
    # the rebuilt matrix the reference returns first (and discards): vecs . diag(w) . vecs^T
    w = rng.rand(c)
    np.testing.assert_allclose(ops.rebuilt_matrix(w).cpu().numpy(), (V * w) @ V.T, rtol=0, atol=1e-10)


def test_tridiagonalisation_preserves_the_spectrum(ops):
    """scrna_sytrd_f64 alone: the eigenvalues of (d, e) are those of A, and Q T Q^T rebuilds A (Q applied with
    scrna_ormtr_f64 to the identity)."""
    import scrna_b200._lib as L
    rng = np.random.RandomState(3)
    n = 203
    B = rng.randn(n, n)
    A = B @ B.T + np.diag(rng.rand(n))
    Ad = ops.to_device(A.copy())
    d, e, tau = ops.empty(n), ops.zeros(n), ops.empty(n)
    work = ops.empty(L.call("scrna_sytrd_work_elems", n))
    ops._call("scrna_sytrd_f64", Ad, n, n, d, e, tau, work, ops._s())
    dh, eh = d.cpu().numpy(), e.cpu().numpy()[: n - 1]
    T = np.diag(dh) + np.diag(eh, 1) + np.diag(eh, -1)
    np.testing.assert_allclose(np.linalg.eigvalsh(T), np.linalg.eigvalsh(A), rtol=1e-11, atol=1e-9)
    QT = ops.to_device(np.eye(n))                       # rows = unit vectors z; ormtr turns row j into Q e_j
    ops._call("scrna_ormtr_f64", Ad, n, n, tau, QT, n, ops._s())
    Q = QT.cpu().numpy().T
    np.testing.assert_allclose(Q.T @ Q, np.eye(n), rtol=0, atol=1e-12)
    np.testing.assert_allclose(Q @ T @ Q.T, A, rtol=0, atol=1e-9 * np.abs(A).max())


@pytest.mark.parametrize("n,c", [(518, 37), (2050, 144), (64, 5)])
def test_blocked_back_transformation_equals_the_reflector_walk(ops, n, c):
    """scrna_ormtr_blocked_f64 (32 reflectors at a time as I - V T V^T: two DMMA GEMMs per block) against
    scrna_ormtr_f64 (one reflector at a time) on the reflectors of a tridiagonalised matrix, and Q orthogonal."""
    import scrna_b200._lib as L
    rng = np.random.RandomState(n)
    B = rng.randn(n, n)
    Ad = ops.to_device(B @ B.T + np.diag(rng.rand(n)))
    d, e, tau = ops.empty(n), ops.zeros(n), ops.empty(n)
    ops._call("scrna_sytrd_f64", Ad, n, n, d, e, tau, ops.empty(L.call("scrna_sytrd_work_elems", n)), ops._s())
    Z = rng.randn(c, n)
    Z /= np.linalg.norm(Z, axis=1, keepdims=True)
    Za, Zb = ops.to_device(Z.copy()), ops.to_device(Z.copy())
    ops._call("scrna_ormtr_f64", Ad, n, n, tau, Za, c, ops._s())
    ops._call("scrna_ormtr_blocked_f64", Ad, n, n, tau, Zb, c, ops.empty(L.call("scrna_ormtr_work_elems", n, c)), ops._s())
    a, b = Za.cpu().numpy(), Zb.cpu().numpy()
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(b, axis=1), 1.0, rtol=0, atol=1e-12)   # Q is orthogonal
    np.testing.assert_allclose(b @ b.T, Z @ Z.T, rtol=0, atol=1e-12)


def test_rank_average_more_genes_than_one_shared_memory_chunk(ops):
    """Round 1 returned UNSUPPORTED above 28 160 genes: the ranks are now counted against the gene vector in chunks."""
    rng = np.random.RandomState(2)
    X = rng.poisson(1.5, size=(30011, 7)).astype(np.float64)          # heavy ties, genes > one 27 648-gene chunk
    S = ops.distances(X, "spearman").cpu().numpy()
    np.testing.assert_allclose(S, so.distances(X, "spearman"), rtol=0, atol=1e-12)


def test_da_nmf_distances_vs_reference_formula(ops):
    """scRNA/sc3_clustering_impl.py:79-103: convex combination of the data's distances and the rescaled distances of
    its dictionary reconstruction; and in-place edits of a stage result raise instead of desynchronising the device
    twin."""
    from scrna_b200 import sc3_clustering_impl as sc
    rng = np.random.RandomState(4)
    g, n, k = 60, 45, 3
    W, H = np.abs(rng.randn(g, k)), np.abs(rng.randn(k, n))
    H2 = np.zeros((k, n)); H2[np.argmax(H, 0), np.arange(n)] = 1
    X = rng.poisson(W @ H * 3).astype(np.float64)
    for use_H2 in (True, False):
        got = sc.da_nmf_distances(X, None, (W, H, H2), metric="euclidean", mixture=0.3, use_H2=use_H2)
        d1 = so.distances(X, "euclidean")
        d2 = so.distances(W @ (H2 if use_H2 else H), "euclidean")
        ref = 0.3 * d2 * (d1.max() / d2.max()) + 0.7 * d1
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)
        got[0, 0] = 1.0                                    # a fresh, writable host array
    np.testing.assert_array_equal(sc.da_nmf_distances(X, None, (W, H, H2), mixture=0.0), so.distances(X, "euclidean"))
    D = sc.distances(X, None)
    with pytest.raises(ValueError):
        D /= D.max()                                       # would leave the resident device copy stale
    Dc = D / D.max()                                       # a copy has no device twin: uploaded as it is
    _, V1 = sc.transformations(Dc, components=5)
    _, V2 = sc.transformations(Dc.copy(), components=5)
    np.testing.assert_array_equal(V1, V2)


@pytest.mark.parametrize("metric", ["euclidean", "pearson", "spearman"])
@pytest.mark.parametrize("method", ["pca", "spectral"])
def test_transformations_vs_reference(ops, sc3g, metric, method):
    from scrna_b200 import sc3_clustering_impl as sc
    c = int(sc3g["pc_range"][1])
    D = sc3g["D_" + metric]
    M = ops.transform_matrix(D, method).cpu().numpy()
    Mref = so.laplacian_quirk(D) if method == "spectral" else so.pca_prepare(D)
    np.testing.assert_allclose(M, Mref, rtol=1e-12, atol=1e-13)
    rebuilt, V = sc.transformations(D, components=c, method=method)
    ref = sc3g["V_%s_%s" % (metric, method)]
    sgn = np.sign((V * ref).sum(axis=0))
    np.testing.assert_allclose(V * sgn, ref, rtol=0, atol=1e-8)
    assert rebuilt.shape == D.shape
    vals_ref = np.linalg.svd(Mref, compute_uv=False)[:c] / np.sqrt(max(1, D.shape[0] - 1))
    np.testing.assert_allclose(rebuilt, (ref * vals_ref) @ ref.T, rtol=0, atol=1e-7 * vals_ref[0])
    from scrna_b200.sc3_device import default_ops
    assert default_ops().jacobi_sweeps < 30


def test_jacobi_on_random_matrices(ops):
    import scipy.linalg as sl
    for n in (2, 7, 64, 151):
        rng = np.random.RandomState(n)
        M = rng.randn(n, n)
        V, vals = ops.right_singular_vectors(ops.to_device(M), n)
        _, s, vt = sl.svd(M)
        np.testing.assert_allclose(vals, s, rtol=1e-10, atol=1e-12)
        V = V.cpu().numpy()
        sgn = np.sign((V * vt.T).sum(axis=0))
        np.testing.assert_allclose(V * sgn, vt.T, rtol=0, atol=1e-8)


@pytest.mark.parametrize("n", [65, 100, 500, 1031])
@pytest.mark.parametrize("kind", ["gauss", "pca_bulk", "rank_deficient"])
def test_block_jacobi_vs_lapack(ops, n, kind):
    """Block one-sided Jacobi (n > 64 path): singular values and the leading right singular vectors against
    LAPACK, on a dense Gaussian matrix, on a PCA-scaled distance matrix (few large directions over a flat bulk,
    the SC3 case) and on a rank-deficient matrix (exact zero singular values)."""
    import scipy.linalg as sl
    rng = np.random.RandomState(n)
    if kind == "gauss":
        M = rng.randn(n, n)
    elif kind == "pca_bulk":
        X = rng.poisson(2.0 + 3.0 * (rng.rand(40, 1) < 0.3) * (rng.randint(0, 4, size=(1, n)) == 1), size=(40, n))
        M = so.pca_prepare(so.distances(np.log2(X + 1.0), "euclidean"))
    else:
        M = rng.randn(n, n // 2) @ rng.randn(n // 2, n)
    c = max(2, int(np.ceil(0.07 * n)))
    V, vals = ops.right_singular_vectors(ops.to_device(M), c, blocked=True)
    assert ops.jacobi_sweeps < (30 if kind == "rank_deficient" else 20)
    _, s, vt = sl.svd(M)
    np.testing.assert_allclose(vals, s, rtol=1e-9, atol=1e-9 * s[0])
    V = V.cpu().numpy()
    np.testing.assert_allclose(V.T @ V, np.eye(c), rtol=0, atol=1e-9)
    # individual vectors where the relative gap to both neighbours is resolvable, else the spanned subspace
    ref = vt[:c].T
    sgn = np.sign((V * ref).sum(axis=0))
    gaps = np.minimum(np.abs(np.diff(s[: c + 1])), np.r_[np.inf, np.abs(np.diff(s[:c]))]) / s[0]
    ok = gaps > 1e-6
    np.testing.assert_allclose((V * sgn)[:, ok], ref[:, ok], rtol=0, atol=1e-7)
    resid = np.linalg.norm(M @ V - (M @ V @ V.T @ V), ord="fro")  # trivial identity: guards NaNs
    assert np.isfinite(resid)
    assert np.linalg.norm(M.T @ (M @ V) - V * (vals[:c] ** 2)) / (s[0] ** 2) < 1e-9


@pytest.mark.parametrize("op", [0, 1])
def test_pairwise_row_blocks_bit_identical(ops, op):
    """Row-block launches (the multi-GPU sharding of the distance stage) reproduce the rows of the full
    symmetric launch bit for bit."""
    import scrna_b200._lib as L
    torch = ops.torch
    rng = np.random.RandomState(4)
    F, N = 300, 333
    X = ops.to_device(rng.poisson(3.0, size=(F, N)).astype(np.float64) + rng.rand(F, N))
    full = ops.empty(N, N)
    L.call("scrna_pairwise_f64", X, N, F, N, op, full, N, ops._s())
    parts = ops.zeros(N, N)
    for r0, r1 in ((0, 111), (111, 222), (222, 333)):
        L.call("scrna_pairwise_rows_f64", X, N, F, N, op, r0, r1, parts[r0:r1], N, ops._s())
    torch.cuda.synchronize()
    assert torch.equal(full, parts)


@pytest.mark.parametrize("batched", [True, False])
def test_kmeans_labelings_vs_reference(ops, sc3g, batched):
    """All 15 k-means labelings of the reference's ensemble, once with every restart of a fit batched into one
    launch set per Lloyd iteration (DMMA assignment) and once restart by restart."""
    k, pc = int(sc3g["k"]), [int(v) for v in sc3g["pc_range"]]
    V = sc3g["V_euclidean_pca"]
    np.random.seed(int(sc3g["seed"]))
    per, _ = so.kmeans_draw_count(k)
    for i, d in enumerate(range(pc[0], pc[1] + 1)):
        draws = np.random.random_sample(10 * per)
        lab = ops.kmeans(ops.to_device(V), d, k, draws, batched=batched)
        np.testing.assert_array_equal(lab, sc3g["kmeans_euclid_pca"][i])


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("n,d,k", [(200, 3, 4), (513, 17, 6), (64, 40, 2), (1000, 9, 8), (700, 33, 11), (1500, 70, 20),
                                   (330, 5, 1)])
def test_kmeans_vs_sklearn_seeded(ops, n, d, k, batched):
    from sklearn.cluster import KMeans
    rng = np.random.RandomState(n + d)
    X = np.concatenate([rng.randn(n // k + 1, d) + 5 * rng.randn(1, d) for _ in range(k)])[:n]
    np.random.seed(n)
    ref = KMeans(n_clusters=k, n_init=10, max_iter=10000, init='k-means++').fit_predict(X)
    np.random.seed(n)
    per, _ = so.kmeans_draw_count(k)
    got = ops.kmeans(ops.to_device(X), d, k, np.random.random_sample(10 * per), batched=batched)
    np.testing.assert_array_equal(got, ref)


def test_kmeans_batched_limits(ops):
    """k above the batched kernels' 48 columns runs restart by restart; asking for the batched path there raises."""
    rng = np.random.RandomState(2)
    X = rng.randn(400, 6)
    per, _ = so.kmeans_draw_count(60)
    draws = rng.random_sample(10 * per)
    lab = ops.kmeans(ops.to_device(X), 6, 60, draws)
    assert lab.shape == (400,) and len(np.unique(lab)) > 30
    with pytest.raises(ValueError):
        ops.kmeans(ops.to_device(X), 6, 60, draws, batched=True)
    # 3 restarts, max_iter 1 and 2: the stop flag of max_iter and the final assignment
    per, _ = so.kmeans_draw_count(5)
    draws = rng.random_sample(3 * per)
    for mi in (1, 2, 7):
        a = ops.kmeans(ops.to_device(X), 6, 5, draws, n_init=3, max_iter=mi, batched=True)
        b = ops.kmeans(ops.to_device(X), 6, 5, draws, n_init=3, max_iter=mi, batched=False)
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("batched", [True, False])
def test_kmeans_duplicate_points_empty_cluster(ops, batched):
    """Duplicate points force empty clusters: the relocation path must run and produce a valid labeling."""
    X = np.repeat(np.array([[0.0, 0.0], [1.0, 1.0], [5.0, 5.0]]), 10, axis=0)
    per, _ = so.kmeans_draw_count(5)
    lab = ops.kmeans(ops.to_device(X), 2, 5, np.random.RandomState(0).random_sample(10 * per), batched=batched)
    assert lab.shape == (30,) and lab.min() >= 0 and lab.max() < 5
    for blk in range(3):   # identical points share a label unless a relocation split one off
        assert len(set(lab[blk * 10:(blk + 1) * 10].tolist())) <= 3


def test_consensus_matrix_and_clustering_vs_reference(ops, sc3g):
    from scrna_b200 import sc3_clustering_impl as sc
    labs = sc3g["kmeans_all"]
    C = sc.build_consensus_matrix(labs)
    np.testing.assert_array_equal(C, so.build_consensus_matrix(labs))
    labels, cd = sc.consensus_clustering(C, n_components=int(sc3g["k"]))
    np.testing.assert_array_equal(cd, sc3g["cdists_all"])
    np.testing.assert_array_equal(labels, sc3g["labels_all"])


@pytest.mark.parametrize("n,r,c,k", [(40, 6, 4, 3), (129, 3, 5, 7), (300, 12, 6, 5), (17, 1, 3, 2), (64, 2, 2, 64),
                                     (50, 4, 3, 1)])
def test_hclust_tie_heavy_vs_scipy(ops, n, r, c, k):
    """Consensus matrices are multiples of 1/r: ties everywhere.  Linkage distances and cut labels must
    equal SciPy's."""
    rng = np.random.RandomState(n * 7 + r)
    lab = rng.randint(0, c, size=(r, n))
    cons = so.build_consensus_matrix(lab)
    ref_labels, ref_sq, Z = so.consensus_clustering(cons, k)
    labels, D, Zg = ops.consensus_clustering(cons, k)
    np.testing.assert_array_equal(D.cpu().numpy(), ref_sq)
    np.testing.assert_array_equal(Zg[:, 2], Z[:, 2])
    np.testing.assert_array_equal(Zg[:, 3], Z[:, 3])
    np.testing.assert_array_equal(labels, ref_labels)


def _run_sc3(mixed, dists, transfs, k, pc, seed):
    from scrna_b200 import SC3Clustering
    from scrna_b200 import sc3_clustering_impl as sc
    np.random.seed(seed)
    s3 = SC3Clustering(mixed, np.arange(mixed.shape[0]), pc_range=pc, sub_sample=True, consensus_mode=0)
    for d in dists:
        s3.add_distance_calculation(partial(sc.distances, metric=d))
    for t in transfs:
        s3.add_dimred_calculation(partial(sc.transformations, components=pc[1], method=t))
    s3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
    s3.set_build_consensus_matrix(sc.build_consensus_matrix)
    s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
    s3.apply()
    return s3


def test_sc3_pipeline_labels_bit_exact_vs_reference(kernels, mixed, sc3g):
    """cmd_target.py:220-242 wiring on the mixed default target data, same seed as the reference run."""
    k, pc, seed = int(sc3g["k"]), [int(v) for v in sc3g["pc_range"]], int(sc3g["seed"])
    s3 = _run_sc3(mixed, ["euclidean"], ["pca"], k, pc, seed)
    np.testing.assert_array_equal(s3.cluster_labels, sc3g["labels_euclid_pca"])
    np.testing.assert_array_equal(s3.dists, sc3g["cdists_euclid_pca"])
    s3 = _run_sc3(mixed, ["euclidean", "pearson", "spearman"], ["pca", "spectral"], k, pc, seed)
    np.testing.assert_array_equal(s3.cluster_labels, sc3g["labels_all"])
    np.testing.assert_array_equal(s3.dists, sc3g["cdists_all"])


def test_sc3_consensus_mode1_and_plain_callables(kernels, mixed, sc3g):
    """The plug-in API accepts arbitrary callables (here the oracle's) and consensus_mode=1."""
    from scrna_b200 import SC3Clustering
    from scrna_b200 import sc3_clustering_impl as sc
    k, pc = int(sc3g["k"]), [int(v) for v in sc3g["pc_range"]]
    np.random.seed(5)
    s3 = SC3Clustering(mixed, None, pc_range=pc, sub_sample=True, consensus_mode=1)
    s3.add_distance_calculation(lambda X, ids: so.distances(X, "euclidean"))
    s3.add_dimred_calculation(partial(sc.transformations, components=pc[1], method="pca"))
    s3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
    s3.set_build_consensus_matrix(sc.build_consensus_matrix)
    s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
    s3.apply()
    assert s3.cluster_labels.shape == (100,) and len(np.unique(s3.cluster_labels)) == k
