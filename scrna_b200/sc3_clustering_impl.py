"""SC3 building blocks with the reference's names and signatures (scRNA/sc3_clustering_impl.py).

Host-side pieces (O(g*n) once): `cell_filter` (:32-42), `gene_filter` (:45-58),
`data_transformation_log2` (:61-67), `no_data_transformation` (:70-76).
Device pieces are added below as the SC3 kernels land (distances, transformations, k-means ensemble,
consensus matrix, consensus clustering).
"""
import numpy as np


def cell_filter(data, num_expr_genes=2000, non_zero_threshold=2):
    """Indices of cells expressing at least `num_expr_genes` genes at >= non_zero_threshold.
    NaNs are zeroed IN PLACE in the caller's matrix, as the reference does (:38-39)."""
    data[np.isnan(data)] = 0
    res = np.sum(data >= non_zero_threshold, axis=0)
    return np.where(np.isfinite(res) & (res >= num_expr_genes))[0]


def gene_filter(data, perc_consensus_genes=0.94, non_zero_threshold=2):
    """Indices of genes that are neither rare nor ubiquitous (:45-58): at least n*(1-p) cells at
    >= threshold and at most n*p cells above zero.  NaNs zeroed in place."""
    data[np.isnan(data)] = 0
    num_cells = data.shape[1]
    res_l = np.sum(data >= non_zero_threshold, axis=1)
    res_h = np.sum(data > 0, axis=1)
    lower_bound = float(num_cells) * (1. - perc_consensus_genes)
    upper_bound = float(num_cells) * perc_consensus_genes
    return np.where((res_l >= lower_bound) & (res_h <= upper_bound))[0]


def data_transformation_log2(data):
    return np.log2(data + 1.)


def no_data_transformation(data):
    return data
