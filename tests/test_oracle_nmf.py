"""The oracle (oracle/nmf_oracle.py) against the golden outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from the unmodified reference)."""
import numpy as np
import pytest

from oracle import nmf_oracle as no

CLI = dict(alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)   # cmd_source.py:176-193 defaults


def test_nmf_clustering_bitwise(default_sim, golden):
    g = golden("nmf_source")
    W0, H0 = no.nndsvdar_init(default_sim["src"], 8)
    np.testing.assert_array_equal(W0, g["W0"])
    np.testing.assert_array_equal(H0, g["H0"])
    W, H, labels, n_iter = no.nmf_clustering_apply(default_sim["src"], 8, **CLI)
    assert n_iter == int(g["nmf_n_iter"]) == 179
    np.testing.assert_array_equal(W, g["nmf_W"])
    np.testing.assert_array_equal(H, g["nmf_H"])
    np.testing.assert_array_equal(labels, g["nmf_labels"])


def test_initw_bitwise(default_sim, golden):
    g = golden("nmf_source")
    D, M, labels, (n1, n2) = no.nmf_clustering_initw_apply(default_sim["src"], default_sim["src_labels"], 7, **CLI)
    assert (n1, n2) == (2, 163)
    np.testing.assert_array_equal(D, g["initw_dictionary"])
    np.testing.assert_array_equal(M, g["initw_data_matrix"])
    np.testing.assert_array_equal(labels, g["initw_labels"])


def test_mu_bitwise(default_sim, golden):
    g = golden("nmf_source")
    W, H, n_iter, errors = no.nmf_mu(default_sim["src"], g["W0"], g["H0"], 7.5, 7.5, 2.5, 2.5, tol=1e-4,
                                     max_iter=200)
    assert n_iter == int(g["mu_n_iter"])
    np.testing.assert_array_equal(W, g["mu_W"])
    np.testing.assert_array_equal(H, g["mu_H"])


def test_transfer_bitwise(default_sim, golden):
    g = golden("transfer")
    src = golden("nmf_source")
    H0 = no.transfer_h_init(g["randn"])
    H, H2, err, n_iter = no.transfer_mu(src["initw_dictionary"], default_sim["trg"], H0)
    np.testing.assert_array_equal(H, g["H"])
    np.testing.assert_array_equal(H2, g["H2"])
    labels = np.argmax(H, axis=0)
    np.testing.assert_array_equal(labels, g["labels"])
    mixed, new_trg = no.mixed_data(src["initw_dictionary"], H, H2, default_sim["trg"], 0.5)
    np.testing.assert_array_equal(mixed, g["mixed"])
    np.testing.assert_array_equal(new_trg, g["new_trg"])
    rej = no.calc_rejection(default_sim["trg"], src["initw_dictionary"], H, H2, labels, 7)
    assert [n for n, _ in rej] == list(g["rej_names"])
    for i, (_, v) in enumerate(rej):
        np.testing.assert_array_equal(v, g["rej_%d" % i])


def test_danmf_apply_bitwise(golden):
    g = golden("transfer")
    W, H, labels, _ = no.nmf_clustering_apply(g["mixed"], 7, alpha=10., l1=0.75, max_iter=4000, rel_err=1e-6)
    np.testing.assert_array_equal(W, g["apply_W"])
    np.testing.assert_array_equal(H, g["apply_H"])
    np.testing.assert_array_equal(labels, g["apply_labels"])


def test_cd_sweep_orders_agree():
    rng = np.random.RandomState(3)
    W = np.abs(rng.randn(40, 6))
    W[rng.rand(40, 6) < 0.3] = 0
    A = rng.randn(6, 6)
    HHt = A @ A.T
    XHt = rng.randn(40, 6)
    perm = rng.permutation(6)
    W1, W2 = W.copy(), W.copy()
    v1 = no.cd_sweep(W1, HHt, XHt, perm)
    v2 = no.cd_sweep_rowmajor(W2, HHt, XHt, perm)
    np.testing.assert_array_equal(W1, W2)
    assert abs(v1 - v2) <= 1e-12 * abs(v2)


def test_perm_stream_matches_sklearn_consumption():
    perms = no.draw_cd_perms(8, 6, seed=0)
    rng = np.random.RandomState(0)
    for i in range(6):
        np.testing.assert_array_equal(perms[i], rng.permutation(8))


def test_matching_gene_inds_first_occurrence():
    src = np.array(['a', 'b', 'c', 'b', 'd'])
    trg = np.array(['d', 'x', 'b', 'a', 'd'])
    i1, i2 = no.matching_gene_inds(src, trg)
    np.testing.assert_array_equal(trg[i1], src[i2])
    np.testing.assert_array_equal(i1, [0, 2, 3, 0])
    np.testing.assert_array_equal(i2, [4, 1, 0, 4])


def test_toy_golden_transfer(golden):
    g = golden("toys")
    H0 = no.transfer_h_init(g["t3_randn"])
    H, H2, err, n_iter = no.transfer_mu(g["t3_W"], g["t3_trg"], H0)
    np.testing.assert_array_equal(H, g["t3_H"])
    mixed, new = no.mixed_data(g["t3_W"], H, H2, g["t3_trg"], 0.25)
    np.testing.assert_array_equal(mixed, g["t3_mixed"])


@pytest.mark.skipif(not __import__("oracle.refshim", fromlist=["x"]).reference_available(),
                    reason="live reference only exists in the build container")
def test_oracle_against_live_reference_midsize():
    """A second, non-default shape checked against the reference run live (build container only)."""
    from oracle import refshim
    ref = refshim.import_reference()
    rng = np.random.RandomState(5)
    X = rng.poisson(rng.gamma(2.0, 1.0, size=(300, 1)) * np.ones((1, 220))).astype(float)
    nmf = ref.nmf_clustering.NmfClustering(X, np.arange(300), 5, None)
    nmf.apply(k=5, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3)
    W, H, labels, _ = no.nmf_clustering_apply(X, 5, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3)
    np.testing.assert_array_equal(W, nmf.dictionary)
    np.testing.assert_array_equal(H, nmf.data_matrix)
