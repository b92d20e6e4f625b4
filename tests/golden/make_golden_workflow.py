"""Golden for the CLI-level work loop: the body of scRNA/cmd_target.py:179-260 -- per
(k, mixture): DaNmfClustering.get_mixed_data -> SC3Clustering -> unsupervised_acc_kta / ARI -- executed with the
UNMODIFIED reference classes (oracle/refshim.py) on the default simulated data, source model from
scRNA/cmd_source.py:191-195's NmfClustering_initW call.

    python tests/golden/make_golden_workflow.py        ->  tests/golden/workflow.npz
"""
import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402


def main():
    ref = refshim.import_reference()
    from sklearn import metrics
    sc = ref.sc3_clustering_impl
    d = np.load(os.path.join(HERE, "default_sim.npz"))
    ds, dt = d["src"].astype(np.float64), d["trg"].astype(np.float64)
    ls, lt = d["src_labels"], d["trg_labels"]
    gene_ids = np.arange(1000)
    cf = partial(sc.cell_filter, num_expr_genes=0, non_zero_threshold=-1)
    gf = partial(sc.gene_filter, perc_consensus_genes=1, non_zero_threshold=-1)
    k = 7
    nmf = ref.nmf_clustering.NmfClustering_initW(ds, gene_ids, labels=ls, num_cluster=k)
    nmf.add_cell_filter(cf); nmf.add_gene_filter(gf); nmf.set_data_transformation(sc.no_data_transformation)
    nmf.apply(k=k, alpha=10.0, l1=0.75, max_iter=4000, rel_err=1e-3)
    src_kta = ref.utils.unsupervised_acc_kta(nmf.pp_data, nmf.cluster_labels, kernel='linear')
    mixtures = [0.0, 0.3, 0.7, 1.0]
    accs = np.zeros((2, 1, len(mixtures)))
    trg_labels = np.zeros((dt.shape[1], 1, len(mixtures)))
    np.random.seed(5)
    for j, mix in enumerate(mixtures):
        nmf.cell_filter_list = list(); nmf.gene_filter_list = list()
        nmf.add_cell_filter(lambda x: np.arange(x.shape[1]).tolist())
        nmf.add_gene_filter(lambda x: np.arange(x.shape[0]).tolist())
        nmf.set_data_transformation(lambda x: x)
        da = ref.nmf_clustering.DaNmfClustering(nmf, dt.copy(), gene_ids.copy(), k)
        da.add_cell_filter(cf); da.add_gene_filter(gf); da.set_data_transformation(sc.no_data_transformation)
        mixed, _, _ = da.get_mixed_data(mix=mix, calc_transferability=False)
        n = mixed.shape[1]
        hi, lo = int(np.ceil(n * 0.07)), int(np.max([1, np.floor(n * 0.04)]))
        s3 = ref.sc3_clustering.SC3Clustering(mixed, pc_range=[lo, hi], sub_sample=True, consensus_mode=0)
        s3.add_distance_calculation(partial(sc.distances, metric='euclidean'))
        s3.add_dimred_calculation(partial(sc.transformations, components=hi, method='pca'))
        s3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
        s3.set_build_consensus_matrix(sc.build_consensus_matrix)
        s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
        s3.apply()
        trg_labels[:, 0, j] = s3.cluster_labels
        accs[0, 0, j] = ref.utils.unsupervised_acc_kta(mixed.copy(), s3.cluster_labels.copy(), kernel='linear')
        accs[1, 0, j] = metrics.adjusted_rand_score(lt[da.remain_cell_inds], s3.cluster_labels)
    np.savez_compressed(os.path.join(HERE, "workflow.npz"), mixtures=np.array(mixtures), accs=accs,
                        trg_labels=trg_labels.astype(np.int32), src_kta=np.float64(src_kta),
                        src_labels_out=nmf.cluster_labels.astype(np.int32), seed=np.int64(5))
    print("accs", accs)


if __name__ == "__main__":
    main()
