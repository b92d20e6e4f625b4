"""scrna_b200 -- B200-native (sm_100a) implementation of the scRNA clustering hot path.

Drop-in for the hot path of nicococo/scRNA: `NmfClustering`, `NmfClustering_initW`, `DaNmfClustering`
(scRNA/nmf_clustering.py) and `SC3Clustering` + `sc3_clustering_impl` (scRNA/sc3_clustering*.py), with
the arithmetic in hand-written CUDA behind a C ABI (include/scrna_b200.h).  Importing the package does
not need a GPU; constructing a solver does, and fails loudly otherwise.
"""
from .abstract_clustering import AbstractClustering
from .nmf_clustering import NmfClustering, NmfClustering_initW, NmfClustering_fixW, DaNmfClustering
from .sc3_clustering import SC3Clustering

__all__ = ["AbstractClustering", "NmfClustering", "NmfClustering_initW", "NmfClustering_fixW",
           "DaNmfClustering", "SC3Clustering"]
