"""BASELINE.json configs[0]/[1] end to end on the GPU: source dictionary (NmfClustering_initW) -> target
transfer + mixed data (DaNmfClustering) -> SC3 consensus clustering, all through the reference's class API,
compared with the reference's own outputs (tests/golden)."""
from functools import partial

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_source_to_target_to_sc3_labels_match_reference(kernels, default_sim, golden):
    from scrna_b200 import NmfClustering_initW, DaNmfClustering, SC3Clustering
    from scrna_b200 import sc3_clustering_impl as sc
    g_t, g_s = golden("transfer"), golden("sc3")
    gene_ids = default_sim["gene_ids"]
    src = NmfClustering_initW(default_sim["src"], gene_ids, 7, default_sim["src_labels"])
    src.apply(k=7, alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)
    # cmd_source.py:208-212 pickles the model, cmd_target.py:188-196 reloads it: must survive pickling
    import pickle
    src = pickle.loads(pickle.dumps(src))

    da = DaNmfClustering(src, default_sim["trg"].copy(), gene_ids, 7)
    np.random.seed(1)
    mixed, new_trg, trg = da.get_mixed_data(mix=0.5)
    np.testing.assert_array_equal(da.cluster_labels, g_t["labels"])
    assert np.linalg.norm(mixed - g_t["mixed"]) / np.linalg.norm(g_t["mixed"]) < 1e-4

    k, pc, seed = int(g_s["k"]), [int(v) for v in g_s["pc_range"]], int(g_s["seed"])
    for dists, transfs, key in ((["euclidean"], ["pca"], "labels_euclid_pca"),
                                (["euclidean", "pearson", "spearman"], ["pca", "spectral"], "labels_all")):
        np.random.seed(seed)
        s3 = SC3Clustering(mixed, gene_ids, pc_range=pc, sub_sample=True, consensus_mode=0)
        for d in dists:
            s3.add_distance_calculation(partial(sc.distances, metric=d))
        for t in transfs:
            s3.add_dimred_calculation(partial(sc.transformations, components=pc[1], method=t))
        s3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
        s3.set_build_consensus_matrix(sc.build_consensus_matrix)
        s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
        s3.apply()
        np.testing.assert_array_equal(s3.cluster_labels, g_s[key])


def test_source_model_pickles_like_the_reference(kernels, default_sim):
    """cmd_source.py:208-212 stores the fitted source object with np.savez(src=nmf) (a pickle) and cmd_target.py
    :188-189 loads it back: the object must carry no device state and keep the reference's attribute names."""
    import pickle
    from scrna_b200 import NmfClustering_initW
    nmf = NmfClustering_initW(default_sim["src"][:300, :200], default_sim["gene_ids"][:300], 7,
                              default_sim["src_labels"][:200])
    k = len(np.unique(default_sim["src_labels"][:200]))
    nmf.apply(k=k, alpha=10., l1=0.75, max_iter=200, rel_err=1e-3)
    nmf.cell_filter_list = None   # as the reference does before saving (lambdas do not pickle)
    nmf.gene_filter_list = None
    nmf.data_transf = None
    back = pickle.loads(pickle.dumps(nmf))
    for name in ("dictionary", "data_matrix", "cluster_labels", "pp_data", "remain_gene_inds", "remain_cell_inds",
                 "gene_ids", "num_cluster"):
        a, b = getattr(nmf, name), getattr(back, name)
        assert np.array_equal(np.asarray(a), np.asarray(b)), name


@pytest.mark.parametrize("name", ["trg_true", "src_300", "trg_log", "trg_random_labels", "one_cluster"])
def test_kta_score_vs_reference_golden(kernels, golden, name):
    """Row f3: unsupervised_acc_kta on the device against the reference's own scores (tests/golden/kta.npz)."""
    from scrna_b200.utils import unsupervised_acc_kta
    g = golden("kta")
    X, lab = g[name + "_X"].astype(np.float64), g[name + "_labels"]
    for c in (True, False):
        for nrm in (True, False):
            if name == "one_cluster" and c:
                assert unsupervised_acc_kta(X, lab, center=c, normalize=nrm) == 0.0   # reference: rounding noise
                continue
            ref = float(g["%s_kta_c%d_n%d" % (name, c, nrm)])
            got = unsupervised_acc_kta(X, lab, center=c, normalize=nrm)
            assert abs(got - ref) < 1e-10, (name, c, nrm, got, ref)


def test_kta_rbf_kernel_and_one_sided_fallback(kernels):
    """unsupervised_acc_kta(kernel='rbf') (scRNA/utils.py:108-124, 137-139) against the closed-form restatement, and
    the mixed case of normalize_kernel's identity fallback (one all-zero cell: its centred diagonal is fine but a
    constant-zero label kernel row is not) where the reference takes the FULL Frobenius norm of the normalised side."""
    from oracle import eval_oracle as eo
    from scrna_b200.utils import unsupervised_acc_kta
    rng = np.random.RandomState(9)
    X = rng.poisson(2.0, size=(40, 30)).astype(np.float64)
    lab = rng.randint(0, 3, size=30)
    Y = np.zeros((30, 3)); Y[np.arange(30), lab] = 1
    sq = (X * X).sum(0)
    for param in (50.0, 400.0):
        Kx = np.exp(-(sq[:, None] - 2 * X.T @ X + sq[None, :]) / param)
        Ky = Y @ Y.T
        for c in (True, False):
            for nrm in (True, False):
                a, b = (eo.center_kernel(Kx), eo.center_kernel(Ky)) if c else (Kx, Ky)
                if nrm:
                    a, b = eo.normalize_kernel(a), eo.normalize_kernel(b)
                ref = float((a * b).sum() / np.sqrt((a * a).sum() * (b * b).sum()))
                got = unsupervised_acc_kta(X, lab, kernel='rbf', param=param, center=c, normalize=nrm)
                assert abs(got - ref) < 1e-10, (param, c, nrm, got, ref)
    # linear kernel without centring and an all-zero cell: Kx has a zero diagonal entry -> identity fallback for Kx only
    X0 = X.copy(); X0[:, 4] = 0.0
    ref = eo.unsupervised_acc_kta(X0, lab, center=False, normalize=True)
    got = unsupervised_acc_kta(X0, lab, center=False, normalize=True)
    assert abs(got - ref) < 1e-10, (got, ref)


def test_reject_classifier_matches_reference_formula(kernels):
    """DaNmfClustering.reject_classifier (scRNA/nmf_clustering.py:274-295) against the formulas it is built from."""
    from oracle import eval_oracle as eo
    from scrna_b200 import DaNmfClustering, NmfClustering
    rng = np.random.RandomState(3)
    X = rng.poisson(2.0, size=(30, 25)).astype(np.float64)
    K = X.T @ X
    kurts = rng.rand(25)
    da = DaNmfClustering(NmfClustering(X, np.arange(30), 2), X, np.arange(30), 2)
    got = da.reject_classifier(K, kurts)
    Kn = eo.normalize_kernel(eo.center_kernel(K))
    sinds = np.argsort(kurts)
    best, best_i = -1.0, -1
    for i in range(K.shape[1] - 2):
        y = np.ones(25); y[sinds[:i + 1]] = -1
        kta = (Kn * np.outer(y, y)).sum() / (25 * np.sqrt((Kn * Kn).sum()))
        if kta > best:
            best, best_i = kta, i + 1
    ref = np.ones(25, dtype=int); ref[sinds[:best_i]] = -1
    np.testing.assert_array_equal(got, ref)
