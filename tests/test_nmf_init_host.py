"""Host logic of the device NNDSVDar init (scrna_b200/nmf_init.py) against scikit-learn's `_initialize_nmf`,
with exact numpy products standing in for the two device sweeps."""
import numpy as np
import pytest

from scrna_b200 import nmf_init


class _Products:
    def __init__(self, X):
        self.Xh = X

        class _Shape:
            pass
        self.X = _Shape()
        self.X.g, self.X.n = X.shape

    def x_times(self, Q):
        return self.Xh @ Q

    def xt_times(self, Q):
        return self.Xh.T @ Q


@pytest.mark.parametrize("shape", [(300, 500), (500, 300), (60, 40), (40, 60)])
@pytest.mark.parametrize("k", [3, 8])
def test_matches_sklearn_initialize_nmf(shape, k):
    from sklearn.decomposition._nmf import _initialize_nmf
    rng = np.random.RandomState(shape[0] + k)
    X = rng.poisson(2.0, size=shape).astype(np.float64)
    U, S, Vt = nmf_init.randomized_svd_device(_Products(X), k, seed=0)
    W, H = nmf_init.nndsvdar_from_svd(U, S, Vt, X.mean(), seed=0)
    Wr, Hr = _initialize_nmf(X, k, init="nndsvdar", random_state=0)
    np.testing.assert_allclose(W, Wr, rtol=0, atol=1e-12)
    np.testing.assert_allclose(H, Hr, rtol=0, atol=1e-12)
