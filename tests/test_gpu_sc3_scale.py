"""SC3 at a size where the oracle (LAPACK SVD, scipy pdist, sklearn KMeans on the host) takes minutes: size-independent
properties of every stage on 4 096 cells x 3 000 genes, and determinism of the whole pipeline."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, G, K = 4096, 3000, 6


@pytest.fixture(scope="module")
def data():
    rng = np.random.RandomState(42)
    centers = rng.gamma(2.0, 2.0, size=(G, K))
    lab = rng.randint(0, K, size=N)
    X = np.log2(rng.poisson(centers[:, lab] * 2.0 ** (0.3 * rng.randn(1, N))).astype(np.float64) + 1.0)
    return X, lab


def test_scale_stage_properties(kernels, data):
    from scrna_b200.sc3_device import Sc3Ops
    X, _ = data
    ops = Sc3Ops(kernels)
    torch = ops.torch
    D = ops.distances(X, "euclidean")
    assert float(D.diagonal().abs().max().cpu()) == 0.0
    assert torch.equal(D, D.T.contiguous())
    rng = np.random.RandomState(0)
    Dh = D[:64].cpu().numpy()
    for i, j in zip(rng.randint(0, 64, 40), rng.randint(0, N, 40)):
        # scipy's pdist sums (a-b)^2 sequentially without FMA; so does the kernel: bit-exact
        acc = 0.0
        for f in range(G):
            d = X[f, i] - X[f, j]
            acc = acc + d * d
        assert Dh[i, j] == np.sqrt(acc), (i, j)
    for metric in ("pearson", "spearman"):
        P = ops.distances(X, metric)
        assert float((P - P.T).abs().max().cpu()) < 1e-12 and float(P.diagonal().abs().max().cpu()) < 1e-12
        assert float(P.min().cpu()) >= 0.0 and float(P.max().cpu()) <= 2.0
    c = int(np.ceil(0.07 * N))
    M = ops.transform_matrix(D, "pca")
    V, vals = ops.right_singular_vectors(M, c)
    assert ops.jacobi_sweeps < 25 and np.all(np.diff(vals) <= 1e-9 * vals[0])
    eye = torch.eye(c, dtype=torch.float64, device=V.device)
    assert float((V.T @ V - eye).abs().max().cpu()) < 1e-9
    s2 = torch.from_numpy(vals[:c] ** 2).to(V.device)
    assert float((M.T @ (M @ V) - V * s2).norm().cpu()) / vals[0] ** 2 < 1e-9
    # ||M||_F^2 = sum sigma^2 (the rotations are orthogonal)
    assert abs(float((M * M).sum().cpu()) - float((vals ** 2).sum())) / float((vals ** 2).sum()) < 1e-10


def test_scale_pipeline_is_deterministic_and_finds_the_clusters(kernels, data):
    from functools import partial
    from sklearn.metrics import adjusted_rand_score
    from scrna_b200 import SC3Clustering
    from scrna_b200 import sc3_clustering_impl as sc
    X, lab = data

    def run():
        np.random.seed(5)
        s3 = SC3Clustering(X, np.arange(G), pc_range=[int(0.04 * N), int(np.ceil(0.07 * N))], sub_sample=True,
                           consensus_mode=0)
        s3.add_distance_calculation(partial(sc.distances, metric='euclidean'))
        s3.add_dimred_calculation(partial(sc.transformations, components=int(np.ceil(0.07 * N)), method='pca'))
        s3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=K))
        s3.set_build_consensus_matrix(sc.build_consensus_matrix)
        s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=K))
        s3.apply()
        return s3.cluster_labels

    a, b = run(), run()
    np.testing.assert_array_equal(a, b)
    assert adjusted_rand_score(lab, a) > 0.9
