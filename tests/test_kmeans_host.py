"""Host logic of the k-means stage (no GPU): the first k-means++ centre is chosen on the host from the uniform draw
the reference's RNG stream provides -- it has to be the index NumPy's own RandomState.choice returns."""
import numpy as np
import pytest

from scrna_b200.sc3_device import first_centres, kmeans_draw_count, same_clustering, uniform_choice_cdf


@pytest.mark.parametrize("n", [1, 2, 7, 100, 1000, 4097, 20000])
def test_first_centre_is_numpy_choice(n):
    cdf = uniform_choice_cdf(n)
    p = np.full(n, 1.0 / n)
    for seed in range(25):
        u = np.random.RandomState(seed).random_sample()
        ref = np.random.RandomState(seed).choice(n, p=p)        # consumes exactly that uniform
        assert int(first_centres(cdf, np.array([u]))[0]) == int(ref), (n, seed)
    # the end points of the cdf
    assert int(first_centres(cdf, np.array([0.0]))[0]) == 0
    assert int(first_centres(cdf, np.array([np.nextafter(1.0, 0.0)]))[0]) == n - 1


def test_draw_count_matches_sklearn_greedy_kmeanspp():
    # 1 uniform for the first centre, then (k - 1) x (2 + int(log k)) candidate draws
    for k in (1, 2, 5, 8, 20, 50):
        per, trials = kmeans_draw_count(k)
        assert trials == 2 + int(np.log(k)) and per == 1 + (k - 1) * trials


def test_same_clustering_is_a_relabelling_test():
    a = np.array([0, 0, 1, 2, 2, 1])
    assert same_clustering(a, np.array([2, 2, 0, 1, 1, 0]), 3)
    assert not same_clustering(a, np.array([2, 2, 0, 1, 1, 1]), 3)
    assert not same_clustering(a, np.array([0, 1, 1, 2, 2, 1]), 3)
