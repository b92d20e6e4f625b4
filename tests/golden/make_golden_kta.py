"""Golden KTA scores from the UNMODIFIED reference (scRNA/utils.py:134-152) under oracle/refshim.py:
    python tests/golden/make_golden_kta.py   ->  tests/golden/kta.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402


def main():
    ref = refshim.import_reference()
    d = np.load(os.path.join(HERE, "default_sim.npz"))
    out = {}
    cases = {
        "trg_true": (d["trg"].astype(np.float64), d["trg_labels"].astype(np.int64)),
        "src_300": (d["src"][:, :300].astype(np.float64), d["src_labels"][:300].astype(np.int64)),
        "trg_log": (np.log2(d["trg"].astype(np.float64) + 1.0), d["trg_labels"].astype(np.int64)),
    }
    rng = np.random.RandomState(0)
    cases["trg_random_labels"] = (cases["trg_true"][0], rng.randint(0, 5, size=cases["trg_true"][0].shape[1]))
    cases["one_cluster"] = (cases["trg_true"][0][:, :40], np.zeros(40, dtype=np.int64))
    for name, (X, lab) in cases.items():
        out[name + "_X"] = X.astype(np.float32) if name != "trg_log" else X
        out[name + "_labels"] = lab
        for c in (True, False):
            for nrm in (True, False):
                out["%s_kta_c%d_n%d" % (name, c, nrm)] = ref.utils.unsupervised_acc_kta(
                    X, lab, kernel='linear', param=1.0, center=c, normalize=nrm)
    np.savez_compressed(os.path.join(HERE, "kta.npz"), **out)
    print({k: float(v) for k, v in out.items() if "_kta_" in k})


if __name__ == "__main__":
    main()
