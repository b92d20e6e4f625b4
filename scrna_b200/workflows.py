"""The reference CLI drivers' work loops as library calls: `source_workflow` is
scRNA/cmd_source.py:98-216 after argument parsing and TSV loading, `target_workflow` is scRNA/cmd_target.py:110-276.
Same order of operations, RNG consumption, outputs (model `.npz` under the reference's class path, `%u` TSV label
files) and quirks (the gene intersection goes through a `set`; with source labels given `--cluster-range` does not
change the source rank) -- minus plotting / t-SNE, which are not on the hot path.

What the device buys on top of the per-call kernels: the target matrix is uploaded ONCE for the whole
(cluster number x mixture) grid instead of once per cell of it, and each mixed matrix goes from the transfer kernels
to the SC3 distance kernel without leaving HBM.
"""
from functools import partial

import numpy as np

from . import sc3_clustering_impl as sc
from .nmf_clustering import DaNmfClustering, NmfClustering, NmfClustering_initW
from .persist import load_source_model, save_source_model
from .sc3_clustering import SC3Clustering
from .utils import unsupervised_acc_kta


def _prefilter(data, gene_ids, labels, use_cell_filter, use_gene_filter, transform, min_expr_genes,
               non_zero_threshold, perc_consensus_genes):
    """cmd_source.py:108-127 / cmd_target.py:104-132: filters and log2 BEFORE the classes see the data."""
    cell_inds = np.arange(data.shape[1])
    if use_cell_filter:
        cell_inds = sc.cell_filter(data, num_expr_genes=min_expr_genes, non_zero_threshold=non_zero_threshold)
        data = data[:, cell_inds]
        if labels is not None:
            labels = labels[cell_inds]
    if use_gene_filter:
        gene_inds = sc.gene_filter(data, perc_consensus_genes=perc_consensus_genes, non_zero_threshold=non_zero_threshold)
        data = data[gene_inds, :]
        gene_ids = gene_ids[gene_inds]
    if transform:
        data = sc.data_transformation_log2(data)
    return data, gene_ids, labels, cell_inds


def _intersect_rows(data, gene_ids, other_gene_ids):
    """cmd_source.py:150-155: rows re-ordered by the iteration order of a set intersection (gene_ids themselves are
    NOT permuted -- the reference's quirk, kept)."""
    gene_intersection = list(set(x for x in other_gene_ids).intersection(set(x for x in gene_ids)))
    index = {g: i for i, g in reversed(list(enumerate(list(gene_ids))))}       # list.index: first occurrence
    return data[[index[x] for x in gene_intersection], ]


_PASS_CELLS = partial(sc.cell_filter, num_expr_genes=0, non_zero_threshold=-1)
_PASS_GENES = partial(sc.gene_filter, perc_consensus_genes=1, non_zero_threshold=-1)


def source_workflow(data, gene_ids, labels, gene_ids_trg, cluster_range=(8,), fout=None, nmf_alpha=10.0, nmf_l1=0.75,
                    nmf_max_iter=4000, nmf_rel_err=1e-3, use_cell_filter=False, use_gene_filter=False, transform=False,
                    min_expr_genes=2000, non_zero_threshold=2.0, perc_consensus_genes=0.94, args=None, comm=None):
    """Source clustering for every k of `cluster_range`.  Returns {k: (model, kta, ari)}; with `fout` also writes
    `<fout>_c<k>.npz`, `.labels.tsv`, `.true_labels.tsv` like the reference."""
    from sklearn import metrics
    data = np.array(data, dtype=np.float64)
    data, gene_ids, labels, _ = _prefilter(data, np.asarray(gene_ids), labels, use_cell_filter, use_gene_filter, transform,
                                           min_expr_genes, non_zero_threshold, perc_consensus_genes)
    data = _intersect_rows(data, gene_ids, np.asarray(gene_ids_trg))
    out = {}
    for k in cluster_range:
        if labels is None:   # generate the source labels by NMF clustering (cmd_source.py:175-183)
            gen = NmfClustering(data, gene_ids, num_cluster=k, labels=[])
            gen.add_cell_filter(_PASS_CELLS)
            gen.add_gene_filter(_PASS_GENES)
            gen.set_data_transformation(sc.no_data_transformation)
            gen.apply(k=k, alpha=nmf_alpha, l1=nmf_l1, max_iter=nmf_max_iter, rel_err=nmf_rel_err, comm=comm)
            labels = gen.cluster_labels
        src_labels = np.array(labels, dtype=int)
        k_now = np.unique(src_labels).size
        nmf = NmfClustering_initW(data, gene_ids, labels=labels, num_cluster=k_now)
        nmf.add_cell_filter(_PASS_CELLS)
        nmf.add_gene_filter(_PASS_GENES)
        nmf.set_data_transformation(sc.no_data_transformation)
        nmf.apply(k=k_now, alpha=nmf_alpha, l1=nmf_l1, max_iter=nmf_max_iter, rel_err=nmf_rel_err, comm=comm)
        kta = unsupervised_acc_kta(nmf.pp_data, nmf.cluster_labels, kernel='linear')
        ari = metrics.adjusted_rand_score(np.asarray(labels)[nmf.remain_cell_inds], nmf.cluster_labels)
        if fout is not None:
            save_source_model('{0}_c{1}.npz'.format(fout, k), nmf, args)
            np.savetxt('{0}_c{1}.labels.tsv'.format(fout, k), (nmf.cluster_labels, nmf.remain_cell_inds), fmt='%u',
                       delimiter='\t')
            np.savetxt('{0}_c{1}.true_labels.tsv'.format(fout, k),
                       (nmf.cluster_labels, np.asarray(labels)[nmf.remain_cell_inds]), fmt='%u', delimiter='\t')
        out[k] = (nmf, kta, ari)
    return out


def target_workflow(src_model, data, gene_ids, gene_ids_src, cluster_range=(8,),
                    mixtures=(0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0), labels=None, fout=None,
                    sc3_dists=('euclidean',), sc3_transf=('pca',), use_cell_filter=False, use_gene_filter=False,
                    transform=False, min_expr_genes=2000, non_zero_threshold=2.0, perc_consensus_genes=0.94):
    """Transfer + SC3 for every (k, mixture) (cmd_target.py:179-260).  `src_model`: a fitted source object, a path
    to a model `.npz` (ours or the reference's), or a callable k -> either.  Returns a dict with `accs`
    (2 x len(k) x len(mix): KTA, ARI), `trg_labels` (cells x len(k) x len(mix)), `opt_mix_ind`, `opt_mix_ktas`,
    `opt_mix_aris`, `pred_labels` {k: labels at the KTA-optimal mixture}; with `fout` also writes the reference's
    `<fout>_c<k>.labels.transfercluster.tsv`, `.data.tsv`, `.geneids.tsv`."""
    from sklearn import metrics
    from .kernels import default_kernels
    from .nmf_solver import DeviceMatrix
    data = np.array(data, dtype=np.float64)
    data, gene_ids, labels, cell_inds = _prefilter(data, np.asarray(gene_ids), labels, use_cell_filter, use_gene_filter,
                                                   transform, min_expr_genes, non_zero_threshold, perc_consensus_genes)
    data = _intersect_rows(data, gene_ids, np.asarray(gene_ids_src))
    num_cluster, mixtures = list(cluster_range), list(mixtures)
    accs = np.zeros((2, len(num_cluster), len(mixtures)))
    opt_mix_ind = np.zeros(len(num_cluster))
    opt_mix_aris = np.zeros(len(num_cluster))
    opt_mix_ktas = np.zeros(len(num_cluster))
    trg_labels = np.zeros((data.shape[1], len(num_cluster), len(mixtures)))
    pred = {}
    target_dm = {}          # the uploaded target, keyed by the matched-gene selection: ONE upload for the whole grid
    for i, k in enumerate(num_cluster):
        for j, mix in enumerate(mixtures):
            src = src_model(k) if callable(src_model) else src_model
            src_nmf = load_source_model(src) if isinstance(src, str) else src
            da_nmf = DaNmfClustering(src_nmf, data.copy(), gene_ids.copy(), k)
            da_nmf.add_cell_filter(_PASS_CELLS)
            da_nmf.add_gene_filter(_PASS_GENES)
            da_nmf.set_data_transformation(sc.no_data_transformation)
            mixed_data, _, _ = da_nmf.get_mixed_data(mix=mix, calc_transferability=False, _target_cache=target_dm)
            num_cells = mixed_data.shape[1]
            max_pca_comp = int(np.ceil(num_cells * 0.07))
            min_pca_comp = int(np.max([1, np.floor(num_cells * 0.04)]))
            sc3_mix = SC3Clustering(mixed_data, pc_range=[min_pca_comp, max_pca_comp], sub_sample=True, consensus_mode=0)
            for ds in sc3_dists:
                sc3_mix.add_distance_calculation(partial(sc.distances, metric=ds))
            for ts in sc3_transf:
                sc3_mix.add_dimred_calculation(partial(sc.transformations, components=max_pca_comp, method=ts))
            sc3_mix.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
            sc3_mix.set_build_consensus_matrix(sc.build_consensus_matrix)
            sc3_mix.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
            sc3_mix.apply()
            trg_labels[:, i, j] = sc3_mix.cluster_labels
            accs[0, i, j] = unsupervised_acc_kta(mixed_data.copy(), sc3_mix.cluster_labels.copy(), kernel='linear')
            if labels is not None:
                accs[1, i, j] = metrics.adjusted_rand_score(np.asarray(labels)[da_nmf.remain_cell_inds],
                                                            sc3_mix.cluster_labels)
        opt_mix_ind[i] = np.argmax(accs[0, i, :])
        opt_mix_ktas[i] = accs[0, i, int(opt_mix_ind[i])]
        opt_mix_aris[i] = accs[1, i, int(opt_mix_ind[i])]
        pred[k] = trg_labels[:, i, int(opt_mix_ind[i])].astype(np.int64)
        if fout is not None:
            np.savetxt('{0}_c{1}.labels.transfercluster.tsv'.format(fout, k), (pred[k], cell_inds), fmt='%u',
                       delimiter='\t')
            np.savetxt('{0}_c{1}.data.tsv'.format(fout, k), data, fmt='%u', delimiter='\t')
            np.savetxt('{0}_c{1}.geneids.tsv'.format(fout, k), gene_ids, fmt='%s', delimiter='\t')
    return dict(accs=accs, trg_labels=trg_labels, opt_mix_ind=opt_mix_ind, opt_mix_ktas=opt_mix_ktas,
                opt_mix_aris=opt_mix_aris, pred_labels=pred, mixtures=mixtures, cluster_range=num_cluster)
