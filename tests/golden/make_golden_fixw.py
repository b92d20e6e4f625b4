"""Golden for NmfClustering_fixW (scRNA/nmf_clustering.py:92-130) from the UNMODIFIED reference under
oracle/refshim.py:   python tests/golden/make_golden_fixw.py  ->  tests/golden/fixw.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402


def main():
    ref = refshim.import_reference()
    d = np.load(os.path.join(HERE, "default_sim.npz"))
    src, labels = d["src"].astype(np.float64), d["src_labels"]
    nmf = ref.nmf_clustering.NmfClustering_fixW(src, np.arange(src.shape[0]), 7, labels)
    nmf.apply(k=7, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3)
    np.savez_compressed(os.path.join(HERE, "fixw.npz"), dictionary=np.asarray(nmf.dictionary, dtype=np.float64),
                        data_matrix=np.asarray(nmf.data_matrix, dtype=np.float64),
                        cluster_labels=np.asarray(nmf.cluster_labels))
    print(np.asarray(nmf.dictionary).shape, np.asarray(nmf.data_matrix).shape, np.bincount(nmf.cluster_labels))


if __name__ == "__main__":
    main()
