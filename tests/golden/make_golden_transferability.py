"""Golden transferability scores from the UNMODIFIED reference (scRNA/utils.py:155-184) under oracle/refshim.py:
    python tests/golden/make_golden_transferability.py   ->  tests/golden/transferability.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402


def main():
    ref = refshim.import_reference()
    d = np.load(os.path.join(HERE, "default_sim.npz"))
    g = np.load(os.path.join(HERE, "nmf_source.npz"))
    W = g["initw_dictionary"].astype(np.float64)             # source dictionary (1000 genes x 7)
    trg = d["trg"].astype(np.float64)
    np.random.seed(3)
    _, H, _, _ = ref.utils.get_transferred_data_matrix(W, trg)
    np.random.seed(4)
    score, percs, p_value = ref.utils.get_transferability_score(W, H, trg, reps=8, max_iter=100)
    np.savez_compressed(os.path.join(HERE, "transferability.npz"), H=H, score=score, percs=percs, p_value=p_value)
    print(score, percs, p_value)


if __name__ == "__main__":
    main()
