"""SC3Clustering with the reference's plug-in API (scRNA/sc3_clustering.py:8-110).

The five registration methods take plain callables exactly like the reference; wiring them to
`scrna_b200.sc3_clustering_impl` (as scRNA/cmd_target.py:228-242 does with functools.partial) runs every
stage on the GPU.  `apply()` follows the reference's loop order, RNG consumption (np.random.permutation for
the dimension sub-sample, sc3_clustering.py:85) and consensus_mode semantics.  When the registered consensus
builder is this package's own, the co-association counts are accumulated on the device as integers and the
n x n matrix is formed once (identical values: the reference sums exact 0/1 matrices in fp64).
"""
import numpy as np

from .abstract_clustering import AbstractClustering
from . import sc3_clustering_impl as _impl


class SC3Clustering(AbstractClustering):
    dists_list = None
    dimred_list = None
    intermediate_clustering_list = None
    build_consensus_matrix = None
    consensus_clustering = None
    dists = None
    pc_range = None
    sub_sample = None
    consensus_mode = None

    def __init__(self, data, gene_ids=None, pc_range=[4, 10], sub_sample=True, consensus_mode=0, comm=None):
        """`comm` (opt-in, not in the reference): a TorchDistComm of the ranks that all call apply() on the same
        data; the k-means fits of the ensemble -- independent units -- are then dealt round-robin over the ranks
        and their label vectors all-reduced.  Needs clustering callables that advertise
        `supports_skip` (this package's intermediate_kmeans_clustering does)."""
        super(SC3Clustering, self).__init__(data, gene_ids=gene_ids)
        self.comm = comm if (comm is not None and comm.world > 1) else None
        self.dists_list = []
        self.dimred_list = []
        self.intermediate_clustering_list = []
        self.consensus_clustering = lambda X: np.zeros(X.shape[0])
        self.pc_range = pc_range
        self.sub_sample = sub_sample
        self.consensus_mode = consensus_mode

    def set_consensus_clustering(self, consensus_clustering):
        self.consensus_clustering = consensus_clustering

    def set_build_consensus_matrix(self, build_consensus_matrix):
        self.build_consensus_matrix = build_consensus_matrix

    def add_distance_calculation(self, dist_calculation):
        self.dists_list.append(dist_calculation)

    def add_dimred_calculation(self, dimred_computation):
        self.dimred_list.append(dimred_computation)

    def add_intermediate_clustering(self, intermediate_clustering):
        self.intermediate_clustering_list.append(intermediate_clustering)

    def _uses_device_consensus(self):
        f = self.build_consensus_matrix
        f = getattr(f, "func", f)
        return f is _impl.build_consensus_matrix and self.consensus_mode == 0

    def apply(self):
        assert self.pc_range[0] > 0
        assert self.pc_range[1] < self.num_cells
        X = self.pre_processing()

        # Between this package's OWN stage functions the n x n intermediates stay in HBM (DeviceResult) and the
        # matrix the reference's dimred returns first -- and its only caller, this loop, discards -- is not formed;
        # user-supplied callables get and return NumPy arrays exactly as in the reference.
        own_dimred = bool(self.dimred_list) and all(_impl.is_own(t, _impl.transformations) for t in self.dimred_list)
        dists = []
        for d in self.dists_list:
            if own_dimred and _impl.is_own(d, _impl.distances):
                # with a comm: row blocks of the distance matrix / of the Gram / of the back-transformed vectors
                # are computed by different ranks and all-gathered (every rank ends with identical matrices)
                dists.append(d(X, self.gene_ids[self.remain_gene_inds], _device=True, _comm=self.comm))
            else:
                dists.append(d(X, self.gene_ids[self.remain_gene_inds]))
        transf = []
        for d in dists:
            for t in self.dimred_list:
                transf.append(t(d, _device=True, _comm=self.comm) if own_dimred else t(d))

        range_inds = list(range(self.pc_range[0], self.pc_range[1] + 1))
        if self.sub_sample and len(range_inds) > 15:
            range_inds = np.random.permutation(range_inds)[:15]

        n = self.remain_cell_inds.size
        # ---- the ensemble: one labeling per (clustering, transformation, d), in the reference's loop order ----
        fits = []       # (cluster index, transformation index, labels or None when another rank owns the fit)
        fit_idx = 0
        for ci, cluster in enumerate(self.intermediate_clustering_list):
            shardable = self.comm is not None and getattr(getattr(cluster, "func", cluster), "supports_skip", False)
            for t in range(len(transf)):
                _, deigv = transf[t]
                for d in range_inds:
                    Xd = deigv[:, 0:d].reshape((deigv.shape[0], d))
                    if shardable and fit_idx % self.comm.world != self.comm.rank:
                        cluster(Xd, _skip=True)          # stay on the RNG stream; the owner computes the fit
                        fits.append((ci, t, None))
                    else:
                        fits.append((ci, t, np.array(cluster(Xd))))
                    fit_idx += 1
        if self.comm is not None and any(l is None for _, _, l in fits):
            self._gather_labels(fits, n)

        cnt = 0.
        device_acc = self._uses_device_consensus()
        if device_acc:
            ops = _impl._ops()
            M = ops.coassoc_new(n)
        else:
            consensus2 = np.zeros((n, n))
        pos = 0
        for ci in range(len(self.intermediate_clustering_list)):
            for t in range(len(transf)):
                labels = []
                for _ in range_inds:
                    labels.append(fits[pos][2])
                    pos += 1
                    if self.consensus_mode == 0:
                        if device_acc:
                            ops.coassoc_add(M, np.array(labels[-1]))
                        else:
                            consensus2 += self.build_consensus_matrix(np.array(labels[-1]))
                        cnt += 1.
                if self.consensus_mode == 1:
                    consensus2 += self.build_consensus_matrix(np.array(labels))
                    cnt += 1.
        if device_acc:
            C = ops.consensus_from_counts(M, cnt)
            if _impl.is_own(self.consensus_clustering, _impl.consensus_clustering):
                consensus2 = _impl.DeviceResult(C)
            else:
                consensus2 = _impl._remember(C.cpu().numpy(), C)
        else:
            consensus2 /= cnt
        if self.comm is not None and _impl.is_own(self.consensus_clustering, _impl.consensus_clustering):
            self.cluster_labels, self.dists = self.consensus_clustering(consensus2, _comm=self.comm)
        else:
            self.cluster_labels, self.dists = self.consensus_clustering(consensus2)

    def _gather_labels(self, fits, n):
        """All-reduce (sum) of an (n_fits x cells) label table in which every rank filled the rows it owns."""
        import torch
        table = np.zeros((len(fits), n), dtype=np.int64)
        for i, (_, _, l) in enumerate(fits):
            if l is not None:
                table[i] = np.asarray(l).reshape(-1)
        use_cuda = torch.cuda.is_available() and self.comm.dist.get_backend(self.comm.group) == "nccl"
        tt = torch.from_numpy(table)
        if use_cuda:
            tt = tt.cuda()
        self.comm.all_reduce(tt)
        table = tt.cpu().numpy()
        for i, (ci, t, l) in enumerate(fits):
            if l is None:
                fits[i] = (ci, t, table[i].copy())
