"""Pipeline shell shared by the clustering classes (host side, NumPy).

Mirrors the public surface of the reference's AbstractClustering (scRNA/abstract_clustering.py:5-102):
constructor (fp64 copy of the genes x cells matrix, :22-34), `add_cell_filter` / `add_gene_filter` /
`set_data_transformation` (:36-49), `pre_processing()` (:51-83) and the result attributes `pp_data`,
`remain_gene_inds`, `remain_cell_inds`, `cluster_labels`.  Pre-processing is O(g*n) once and stays
on the host; the quirk that gene filters see the un-cell-filtered data
(:69) is kept.
"""
import numpy as np


class AbstractClustering(object):
    """Attributes (same names as the reference): data, gene_ids, num_transcripts, num_cells, the filter lists and
    the transformation, and after pre_processing(): pp_data, remain_gene_inds, remain_cell_inds; subclasses fill
    cluster_labels."""

    def __init__(self, data, gene_ids=None):
        matrix = np.array(data, dtype=np.float64)          # always a private fp64 copy of genes x cells
        n_genes, n_cells = matrix.shape
        self.__dict__.update(
            data=matrix, num_transcripts=n_genes, num_cells=n_cells,
            gene_ids=np.arange(n_genes) if gene_ids is None else gene_ids,
            cell_filter_list=[], gene_filter_list=[], data_transf=_identity,
            pp_data=None, remain_gene_inds=None, remain_cell_inds=None,
            cluster_labels=np.zeros((n_cells, 1)))

    def set_data_transformation(self, data_transf):
        self.data_transf = data_transf

    def add_cell_filter(self, cell_filter):
        if self.cell_filter_list is None:
            self.cell_filter_list = []
        self.cell_filter_list.append(cell_filter)

    def add_gene_filter(self, gene_filter):
        if self.gene_filter_list is None:
            self.gene_filter_list = []
        self.gene_filter_list.append(gene_filter)

    def pre_processing(self):
        self.pp_data, self.remain_gene_inds, self.remain_cell_inds = self.pre_processing_impl(self.data)
        return self.pp_data

    def pre_processing_impl(self, data):
        n_genes, n_cells = data.shape
        keep_cells = np.arange(n_cells)
        for flt in (self.cell_filter_list or []):
            keep_cells = np.intersect1d(keep_cells, flt(data))
        keep_genes = np.arange(n_genes)
        for flt in (self.gene_filter_list or []):
            keep_genes = np.intersect1d(keep_genes, flt(data))   # evaluated on ALL cells, as the reference
        # data[:, keep_cells][keep_genes, :] as the reference, minus the fancy-index passes that select everything
        # (two gather copies of a 3.2 GB matrix at 20 000 x 20 000); the result is always a fresh array
        a = data if keep_cells.size == n_cells else data[:, keep_cells]
        if keep_genes.size != n_genes:
            sliced = a[keep_genes, :]
        elif a is not data:
            sliced = a
        elif self.data_transf is _identity:
            # nothing filtered, nothing transformed: pp_data is a READ-ONLY view of the private copy the constructor
            # made (same values as the reference's fresh array; a write raises instead of touching `data`) -- the
            # copy is 3.2 GB of freshly faulted pages (2-3 s) at 20 000 x 20 000
            sliced = data.view()
            sliced.flags.writeable = False
            return sliced, keep_genes, keep_cells
        else:
            sliced = a.copy()          # a user transformation may work in place
        return self.data_transf(sliced), keep_genes, keep_cells

    def apply(self):
        raise NotImplementedError

    def __getstate__(self):
        # lambdas / partials do not pickle (the reference nulls the filter lists before np.savez,
        # cmd_source.py:208-211); device handles never enter the state either.
        state = dict(self.__dict__)
        for key in ("cell_filter_list", "gene_filter_list", "data_transf"):
            state[key] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        if self.cell_filter_list is None:
            self.cell_filter_list = []
        if self.gene_filter_list is None:
            self.gene_filter_list = []
        if self.data_transf is None:
            self.data_transf = _identity

    def __str__(self):
        if self.cluster_labels is None or self.pp_data is None:
            return 'Empty cluster pipeline.'
        lines = ['Cluster Pipeline ({0} processed datapoints, {1} processed features):'.format(
            self.pp_data.shape[1], self.pp_data.shape[0])]
        labels = np.asarray(self.cluster_labels).ravel()
        for lbl in np.unique(labels):
            members = np.where(labels == lbl)[0]
            lines.append('({0})[{1}]'.format(lbl, ','.join(str(i) for i in members)))
        return '\n'.join(lines) + '\n'


def _identity(x):
    return x
