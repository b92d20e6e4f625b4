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


def test_target_workflow_matches_the_reference_cli_loop(kernels, default_sim, golden, tmp_path):
    """Row f2: scrna_b200.workflows.target_workflow (cmd_target.py:179-260 as a library call; target uploaded once
    for the whole mixture grid) against tests/golden/workflow.npz, produced by the same loop over the UNMODIFIED
    reference classes: identical labels for every mixture, KTA to 1e-5, identical ARI, identical KTA-optimal mixture;
    the source model travels through the reference-compatible `.npz`."""
    from scrna_b200.workflows import source_workflow, target_workflow
    g = golden("workflow")
    gene_ids = default_sim["gene_ids"]
    src = source_workflow(default_sim["src"], gene_ids, default_sim["src_labels"], gene_ids, cluster_range=(7,),
                          fout=str(tmp_path / "src"))
    model, kta, _ = src[7]
    np.testing.assert_array_equal(model.cluster_labels, g["src_labels_out"])
    assert abs(kta - float(g["src_kta"])) < 1e-9
    assert (tmp_path / "src_c7.npz").exists() and (tmp_path / "src_c7.labels.tsv").exists()
    mixtures = [float(m) for m in g["mixtures"]]
    launches0 = kernels.launches
    np.random.seed(int(g["seed"]))
    out = target_workflow(str(tmp_path / "src_c7.npz"), default_sim["trg"], gene_ids, gene_ids, cluster_range=(7,),
                          mixtures=mixtures, labels=default_sim["trg_labels"], fout=str(tmp_path / "trg"))
    np.testing.assert_array_equal(out["trg_labels"].astype(np.int32), g["trg_labels"])
    # the mixed data carries our dictionary (within the 1e-3 per-factor budget of the reference's): KTA to 1e-5;
    # mixture 0 is the untouched target: exact
    np.testing.assert_allclose(out["accs"][0], g["accs"][0], rtol=0, atol=1e-5)
    assert abs(out["accs"][0, 0, 0] - g["accs"][0, 0, 0]) < 1e-10
    np.testing.assert_allclose(out["accs"][1], g["accs"][1], rtol=0, atol=1e-12)
    assert int(out["opt_mix_ind"][0]) == int(np.argmax(g["accs"][0, 0]))
    assert kernels.launches > launches0
    lab_file = np.loadtxt(str(tmp_path / "trg_c7.labels.transfercluster.tsv"), delimiter="\t")
    np.testing.assert_array_equal(lab_file[0].astype(np.int64), out["pred_labels"][7])


def test_device_filters_match_numpy_and_keep_the_nan_quirk(kernels, monkeypatch):
    """Row f4: cell_filter / gene_filter over row bands on the device (one pass, integer counts) return exactly the
    NumPy indices, and NaNs are still zeroed in place in the caller's matrix (sc3_clustering_impl.py:38-39, 51-52)."""
    from scrna_b200 import sc3_clustering_impl as sc
    rng = np.random.RandomState(8)
    X = rng.poisson(rng.uniform(0.05, 3.0, size=(700, 1)) * rng.uniform(0.5, 1.5, size=(1, 900))).astype(np.float64)
    X[rng.rand(700, 900) < 0.001] = np.nan
    host, dev = X.copy(), X.copy()
    n_expr = int(np.median(np.nansum(X >= 2, axis=0)))
    monkeypatch.setattr(sc, "DEVICE_FILTER_MIN_ELEMENTS", 1 << 62)
    ref_cells = sc.cell_filter(host, num_expr_genes=n_expr, non_zero_threshold=2)
    ref_genes = sc.gene_filter(host, perc_consensus_genes=0.8, non_zero_threshold=2)
    monkeypatch.setattr(sc, "DEVICE_FILTER_MIN_ELEMENTS", 1)
    l0 = kernels.launches
    got_cells = sc.cell_filter(dev, num_expr_genes=n_expr, non_zero_threshold=2)
    assert kernels.launches > l0 and not np.isnan(dev).any()
    got_genes = sc.gene_filter(dev, perc_consensus_genes=0.8, non_zero_threshold=2)
    np.testing.assert_array_equal(got_cells, ref_cells)
    np.testing.assert_array_equal(got_genes, ref_genes)
    np.testing.assert_array_equal(dev, host)
    assert 0 < ref_cells.size < 900 and 0 < ref_genes.size < 700
