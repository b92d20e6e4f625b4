"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (read-only at
/root/reference) under oracle/refshim.py.  Run once in the build container:

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so the outputs are committed (tests/golden/*.npz) together with
this script.  Library versions used: scikit-learn 1.9.0, scipy 1.18.1, numpy 2.3.5 (the reference
pins none of them).

Fixtures
  default_sim.npz   the "default simulated data" (cmd_generate_data.py defaults: 1000 genes, 2000
                    cells, cluster_spec [1,2,3,[4,5],[6,[7,8]]], split mode 4, 1000 source / 100
                    target cells) drawn with np.random.seed(0); random.seed(0)
  nmf_source.npz    NmfClustering (k=8) and NmfClustering_initW (k=7) on the source data with the CLI
                    parameters alpha=10, l1=.75, max_iter=4000, rel_err=1e-3 (cmd_source.py:176-193),
                    plus sklearn solver='mu' from the same init
  transfer.npz      DaNmfClustering.get_mixed_data(mix=.5) and .apply(mix=.5) with np.random.seed(1)
  toys.npz          the two toy cases of the reference's tests/test_transfer.py
"""
import os
import random
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402


def main():
    ref = refshim.import_reference()
    from sklearn.decomposition._nmf import _fit_multiplicative_update, _initialize_nmf

    # ---- default simulated data ---------------------------------------------------------------
    np.random.seed(0)
    random.seed(0)
    data, labels = ref.simulation.generate_toy_data(
        num_genes=1000, num_cells=2000, cluster_spec=[1, 2, 3, [4, 5], [6, [7, 8]]],
        dirichlet_parameter_cluster_size=10, gamma_shape=2, gamma_rate=2, nb_dispersion=0.1,
        min_prop_genes_de=0.1, max_prop_genes_de=0.4, mean_de_logfc=1, sd_de_logfc=0.5)
    ds, dt, ls, lt = ref.simulation.split_source_target(
        data, labels, target_ncells=100, source_ncells=1000, source_clusters=None, noise_target=False,
        noise_sd=0.5, mode=4, common="[0,3,5]")
    assert ds.max() < 65536 and dt.max() < 65536 and ds.min() >= 0
    gene_ids = np.arange(1000)
    np.savez_compressed(os.path.join(HERE, "default_sim.npz"), src=ds.astype(np.uint16),
                        trg=dt.astype(np.uint16), src_labels=ls.astype(np.int32),
                        trg_labels=lt.astype(np.int32))

    # ---- source NMF -----------------------------------------------------------------------------
    kw = dict(alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)
    nmf = ref.nmf_clustering.NmfClustering(ds, gene_ids, 8, ls)
    nmf.apply(k=8, **kw)
    W0, H0 = _initialize_nmf(ds.astype(np.float64), 8, init="nndsvdar", random_state=0)
    initw = ref.nmf_clustering.NmfClustering_initW(ds, gene_ids, 7, ls)
    initw.apply(k=7, **kw)
    Wmu, Hmu, nmu = _fit_multiplicative_update(ds.astype(np.float64), W0.copy(), H0.copy(), 'frobenius', 200,
                                               1e-4, 7.5, 7.5, 2.5, 2.5, True, 0)
    # iteration counts of the two CD fits (not exposed by the reference classes): re-run through sklearn
    from sklearn import decomposition as decomp
    m = decomp.NMF(alpha=10., init='nndsvdar', l1_ratio=0.75, max_iter=4000, n_components=8, random_state=0,
                   shuffle=True, solver='cd', tol=1e-3, verbose=0)
    m.fit_transform(ds.astype(np.float64))
    np.savez_compressed(
        os.path.join(HERE, "nmf_source.npz"),
        W0=W0, H0=H0, nmf_W=nmf.dictionary, nmf_H=nmf.data_matrix, nmf_labels=nmf.cluster_labels,
        nmf_n_iter=np.int64(m.n_iter_),
        initw_dictionary=initw.dictionary, initw_data_matrix=initw.data_matrix,
        initw_labels=initw.cluster_labels, mu_W=Wmu, mu_H=Hmu, mu_n_iter=np.int64(nmu))

    # ---- target transfer ------------------------------------------------------------------------
    np.random.seed(1)
    draw = np.random.randn(7, 100)
    np.random.seed(1)
    da = ref.nmf_clustering.DaNmfClustering(initw, dt.copy(), gene_ids, 7)
    mixed, new_trg, trg = da.get_mixed_data(mix=0.5)
    W, H, H2 = da.intermediate_model
    rej = {("rej_%d" % i): v for i, (name, v) in enumerate(da.reject)}
    rej_names = np.array([name for name, _ in da.reject])
    labels_t = da.cluster_labels.copy()
    np.random.seed(1)
    da2 = ref.nmf_clustering.DaNmfClustering(initw, dt.copy(), gene_ids, 7)
    da2.apply(k=7, mix=0.5, alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)
    np.savez_compressed(
        os.path.join(HERE, "transfer.npz"), randn=draw, W=W, H=H, H2=H2, labels=labels_t, mixed=mixed,
        new_trg=new_trg, rej_names=rej_names, apply_W=da2.dictionary, apply_H=da2.data_matrix,
        apply_labels=da2.cluster_labels, **rej)

    # ---- toy cases of tests/test_transfer.py ------------------------------------------------------
    sc = ref.sc3_clustering_impl
    src1 = np.array([[1, 2, 0, 0, 0], [0, 0, 10, 11, 0], [0, 0, 0, 0, 0]], dtype=float)
    ids1 = np.array(['lbl0', 'lbl1', 'lbl2'])
    t1 = ref.nmf_clustering.NmfClustering(src1, ids1, 2, None)   # 4th positional arg required (stale test)
    t1.add_cell_filter(partial(sc.cell_filter, num_expr_genes=1, non_zero_threshold=1))
    t1.add_gene_filter(partial(sc.gene_filter, perc_consensus_genes=0.96, non_zero_threshold=1))
    t1.apply()
    src2 = np.array([[1, 2, 0, 0, 0], [0, 0, 10, 11, 0], [1, 2, 1, 2, 0], [0, 0, 0, 0, 0]], dtype=float)
    trg2 = np.array([[2, 4, 6, 6, 0], [2, 1, 2, 2, 0], [1.1, 2.3, 1.2, 2.1, 0], [0, 0, 0, 0, 0]], dtype=float)
    ids2 = np.array(['lbl0', 'lbl1', 'lbl2', 'lbl3'])
    t2 = ref.nmf_clustering.NmfClustering(src2, ids2, 2, None)
    t2.add_cell_filter(partial(sc.cell_filter, num_expr_genes=1, non_zero_threshold=1))
    t2.add_gene_filter(partial(sc.gene_filter, perc_consensus_genes=0.96, non_zero_threshold=1))
    t2.apply()
    trg_ids = np.array(['lbl0', 'lbl1', 'lbl20', 'lbl3'])
    np.random.seed(2)
    draw2 = np.random.randn(2, 4)
    np.random.seed(2)
    t3 = ref.nmf_clustering.DaNmfClustering(t2, trg2, trg_ids, 2)
    t3.add_cell_filter(partial(sc.cell_filter, num_expr_genes=1, non_zero_threshold=1))
    t3.add_gene_filter(partial(sc.gene_filter, perc_consensus_genes=0.96, non_zero_threshold=1))
    mixed3, new3, trg3 = t3.get_mixed_data(mix=0.25)
    np.savez_compressed(
        os.path.join(HERE, "toys.npz"),
        src1=src1, t1_pp=t1.pp_data, t1_genes=t1.remain_gene_inds, t1_cells=t1.remain_cell_inds,
        t1_W=t1.dictionary, t1_H=t1.data_matrix, t1_labels=t1.cluster_labels,
        src2=src2, trg2=trg2, t2_pp=t2.pp_data, t2_genes=t2.remain_gene_inds, t2_cells=t2.remain_cell_inds,
        t2_W=t2.dictionary, t2_H=t2.data_matrix, t2_labels=t2.cluster_labels,
        t3_randn=draw2, t3_gene_ids=t3.gene_ids, t3_W=t3.intermediate_model[0], t3_H=t3.intermediate_model[1],
        t3_H2=t3.intermediate_model[2], t3_mixed=mixed3, t3_new=new3, t3_trg=trg3, t3_labels=t3.cluster_labels)
    # ---- SC3 on the mixed default target data (cmd_target.py:220-242 wiring) ----------------------
    n = mixed.shape[1]
    k = 7
    pc = [max(1, int(np.floor(0.04 * n))), int(np.ceil(0.07 * n))]
    out = {"pc_range": np.array(pc), "k": np.int64(k), "seed": np.int64(3)}
    for m in ("euclidean", "pearson", "spearman"):
        out["D_" + m] = sc.distances(mixed.copy(), np.arange(1000), metric=m)
        for t in ("pca", "spectral"):
            _, vecs = sc.transformations(out["D_" + m].copy(), components=pc[1], method=t)
            out["V_%s_%s" % (m, t)] = vecs

    def run_sc3(dists, transfs, seed):
        labs = []

        def km(X, k):
            lab = sc.intermediate_kmeans_clustering(X, k=k)
            labs.append(np.asarray(lab).copy())
            return lab
        np.random.seed(seed)
        s3 = ref.sc3_clustering.SC3Clustering(mixed, np.arange(1000), pc_range=pc, sub_sample=True,
                                              consensus_mode=0)
        for d in dists:
            s3.add_distance_calculation(partial(sc.distances, metric=d))
        for t in transfs:
            s3.add_dimred_calculation(partial(sc.transformations, components=pc[1], method=t))
        s3.add_intermediate_clustering(partial(km, k=k))
        s3.set_build_consensus_matrix(sc.build_consensus_matrix)
        s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
        s3.apply()
        return s3.cluster_labels, np.array(labs), s3.dists

    l1, km1, cd1 = run_sc3(["euclidean"], ["pca"], 3)
    l2, km2, cd2 = run_sc3(["euclidean", "pearson", "spearman"], ["pca", "spectral"], 3)
    out.update(labels_euclid_pca=l1, kmeans_euclid_pca=km1, cdists_euclid_pca=cd1,
               labels_all=l2, kmeans_all=km2, cdists_all=cd2)
    np.savez_compressed(os.path.join(HERE, "sc3.npz"), **out)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
