"""The SC3 oracle (oracle/sc3_oracle.py) against the golden outputs of the reference itself
(tests/golden/sc3.npz from tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import sc3_oracle as so


@pytest.fixture(scope="module")
def sc3g(golden):
    return golden("sc3")


@pytest.fixture(scope="module")
def mixed(golden):
    return golden("transfer")["mixed"]


def test_distances_match_reference(mixed, sc3g):
    np.testing.assert_array_equal(so.distances(mixed, "euclidean"), sc3g["D_euclidean"])
    np.testing.assert_array_equal(so.euclidean_sequential(mixed), sc3g["D_euclidean"])   # bitwise, probe P8
    np.testing.assert_allclose(so.distances(mixed, "pearson"), sc3g["D_pearson"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(so.distances(mixed, "spearman"), sc3g["D_spearman"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(1 - so.pearson_from_columns(mixed), sc3g["D_pearson"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(1 - so.pearson_from_columns(so.rank_average(mixed)), sc3g["D_spearman"], rtol=0,
                               atol=1e-14)


@pytest.mark.parametrize("metric", ["euclidean", "pearson", "spearman"])
@pytest.mark.parametrize("method", ["pca", "spectral"])
def test_transformations_match_reference(sc3g, metric, method):
    c = int(sc3g["pc_range"][1])
    _, V, _ = so.transformations(sc3g["D_" + metric], components=c, method=method)
    ref = sc3g["V_%s_%s" % (metric, method)]
    # singular vectors are defined up to sign
    sgn = np.sign((V * ref).sum(axis=0))
    np.testing.assert_allclose(V * sgn, ref, rtol=0, atol=1e-9)


def test_full_pipeline_labels_match_reference(mixed, sc3g):
    k, pc, seed = int(sc3g["k"]), [int(v) for v in sc3g["pc_range"]], int(sc3g["seed"])
    np.random.seed(seed)
    labels, cons, labs = so.sc3_apply(mixed, ["euclidean"], ["pca"], k, pc)
    np.testing.assert_array_equal(labels, sc3g["labels_euclid_pca"])
    np.testing.assert_array_equal(np.array(labs), sc3g["kmeans_euclid_pca"])
    np.random.seed(seed)
    labels, cons, labs = so.sc3_apply(mixed, ["euclidean", "pearson", "spearman"], ["pca", "spectral"], k, pc)
    np.testing.assert_array_equal(labels, sc3g["labels_all"])
    np.testing.assert_array_equal(np.array(labs), sc3g["kmeans_all"])
    _, cd, _ = so.consensus_clustering(cons, k)
    np.testing.assert_array_equal(cd, sc3g["cdists_all"])


def test_nn_chain_restatement_matches_scipy(sc3g):
    import scipy.cluster.hierarchy as spc
    import scipy.spatial.distance as dist
    cons = so.build_consensus_matrix(sc3g["kmeans_all"])
    sq = dist.squareform(dist.pdist(cons))
    Z = spc.complete(dist.pdist(cons))
    merges = so.nn_chain_complete(sq)
    d = np.array([m[2] for m in merges])
    order = np.argsort(d, kind="mergesort")
    np.testing.assert_array_equal(d[order], Z[:, 2])
    # tie-heavy random case
    rng = np.random.RandomState(0)
    lab = rng.randint(0, 4, size=(6, 40))
    cons = so.build_consensus_matrix(lab)
    sq = dist.squareform(dist.pdist(cons))
    Z = spc.complete(dist.pdist(cons))
    d = np.array([m[2] for m in so.nn_chain_complete(sq)])
    np.testing.assert_array_equal(np.sort(d, kind="mergesort"), Z[:, 2])


def test_kmeans_restatement_matches_sklearn():
    from sklearn.cluster import KMeans
    rng = np.random.RandomState(4)
    X = np.concatenate([rng.randn(40, 5) + 4 * rng.randn(1, 5) for _ in range(4)])
    np.random.seed(11)
    ref = KMeans(n_clusters=4, n_init=10, max_iter=10000, init='k-means++').fit_predict(X)
    np.random.seed(11)
    per, _ = so.kmeans_draw_count(4)
    got = so.kmeans(X, 4, np.random.random_sample(10 * per))
    np.testing.assert_array_equal(got, ref)
