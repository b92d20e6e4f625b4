"""The KTA restatement (oracle/eval_oracle.py) against the reference's own scores (tests/golden/kta.npz, written by
tests/golden/make_golden_kta.py from scRNA/utils.py:134-152 under the compatibility shim)."""
import numpy as np
import pytest

from oracle import eval_oracle as eo

CASES = ["trg_true", "src_300", "trg_log", "trg_random_labels"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("center", [True, False])
@pytest.mark.parametrize("normalize", [True, False])
def test_kta_oracle_matches_reference(golden, name, center, normalize):
    g = golden("kta")
    ref = float(g["%s_kta_c%d_n%d" % (name, center, normalize)])
    got = eo.unsupervised_acc_kta(g[name + "_X"].astype(np.float64), g[name + "_labels"], center, normalize)
    assert abs(got - ref) < 1e-13


def test_kta_uncentred_single_cluster_matches_reference(golden):
    g = golden("kta")
    for nrm in (True, False):
        ref = float(g["one_cluster_kta_c0_n%d" % nrm])
        got = eo.unsupervised_acc_kta(g["one_cluster_X"].astype(np.float64), g["one_cluster_labels"], False, nrm)
        assert abs(got - ref) < 1e-13
