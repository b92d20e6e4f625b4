"""Parity of the CUDA target-transfer path (K5-K8) against the oracle and the reference goldens."""
import numpy as np
import pytest

from conftest import rel_fro
from oracle import nmf_oracle as no

pytestmark = pytest.mark.gpu


class _Src:
    """Stand-in for the unpickled source model the CLI hands to DaNmfClustering (cmd_target.py:188-196)."""
    def __init__(self, dictionary, gene_ids, num_cluster):
        self.dictionary = dictionary
        self.gene_ids = gene_ids
        self.remain_gene_inds = np.arange(gene_ids.size)
        self.num_cluster = num_cluster


def test_get_transferred_data_matrix_vs_golden(kernels, default_sim, golden):
    from scrna_b200.utils import get_transferred_data_matrix
    g = golden("transfer")
    W = golden("nmf_source")["initw_dictionary"]
    np.random.seed(1)   # same global-RNG draw as the reference (utils.py:189)
    Wout, H, H2, err = get_transferred_data_matrix(W, default_sim["trg"])
    assert Wout is not None and H.shape == (7, 100)
    assert rel_fro(H, g["H"]) < 1e-3
    np.testing.assert_array_equal(H2, g["H2"])
    Href, H2ref, err_ref, n_ref = no.transfer_mu(W, default_sim["trg"], no.transfer_h_init(g["randn"]))
    assert abs(err - err_ref) / err_ref < 1e-5


def test_get_mixed_data_vs_golden(kernels, default_sim, golden):
    from scrna_b200 import DaNmfClustering
    g = golden("transfer")
    src = _Src(golden("nmf_source")["initw_dictionary"], np.arange(1000), 7)
    da = DaNmfClustering(src, default_sim["trg"].copy(), np.arange(1000), 7)
    np.random.seed(1)
    mixed, new_trg, trg = da.get_mixed_data(mix=0.5)
    np.testing.assert_array_equal(da.cluster_labels, g["labels"])
    np.testing.assert_array_equal(trg, default_sim["trg"])
    # W.H2 is an exact gather of fp64 dictionary columns; the mix uses exact (integer) counts
    np.testing.assert_allclose(new_trg, g["new_trg"], rtol=0, atol=0)
    np.testing.assert_allclose(mixed, g["mixed"], rtol=1e-15, atol=0)
    assert [n for n, _ in da.reject] == list(g["rej_names"])
    for i, (name, v) in enumerate(da.reject):
        ref = g["rej_%d" % i]
        assert np.linalg.norm(v - ref) <= 2e-3 * max(np.linalg.norm(ref), 1e-12), name


def test_get_mixed_data_with_dense_H_on_the_device(kernels, default_sim, golden):
    """get_mixed_data(use_H2=False) (nmf_clustering.py:180-184): mix . (W.H) + (1 - mix) . X from the device kernel
    (scrna_mix_dense), against the same expression in NumPy with the session's own W and H."""
    from scrna_b200 import DaNmfClustering
    src = _Src(golden("nmf_source")["initw_dictionary"], np.arange(1000), 7)
    da = DaNmfClustering(src, default_sim["trg"].copy(), np.arange(1000), 7)
    np.random.seed(1)
    launches0 = kernels.launches
    mixed, new_trg, trg = da.get_mixed_data(mix=0.3, use_H2=False)
    assert kernels.launches > launches0
    W, H = da.intermediate_model[0], da.intermediate_model[1]
    ref_new = W.dot(H)
    np.testing.assert_allclose(new_trg, ref_new, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(mixed, 0.3 * ref_new + 0.7 * trg, rtol=1e-12, atol=1e-12)


def test_danmf_apply_vs_golden(kernels, default_sim, golden):
    from scrna_b200 import DaNmfClustering
    g = golden("transfer")
    src = _Src(golden("nmf_source")["initw_dictionary"], np.arange(1000), 7)
    da = DaNmfClustering(src, default_sim["trg"].copy(), np.arange(1000), 7)
    np.random.seed(1)
    da.apply(k=7, mix=0.5, alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)
    assert rel_fro(da.dictionary, g["apply_W"]) < 1e-3
    assert rel_fro(da.data_matrix, g["apply_H"]) < 1e-3
    np.testing.assert_array_equal(da.cluster_labels, g["apply_labels"])


def test_gene_id_mismatch_toy(kernels, golden):
    """Second toy case of the reference's tests/test_transfer.py: one target gene id does not match."""
    from functools import partial
    from scrna_b200 import DaNmfClustering
    from scrna_b200.sc3_clustering_impl import cell_filter, gene_filter
    g = golden("toys")
    ids = np.array(['lbl0', 'lbl1', 'lbl2', 'lbl3'])
    src = _Src(g["t2_W"], ids[g["t2_genes"]], 2)
    src.gene_ids = ids
    src.remain_gene_inds = g["t2_genes"]
    trg_ids = np.array(['lbl0', 'lbl1', 'lbl20', 'lbl3'])
    da = DaNmfClustering(src, g["trg2"], trg_ids, 2)
    da.add_cell_filter(partial(cell_filter, num_expr_genes=1, non_zero_threshold=1))
    da.add_gene_filter(partial(gene_filter, perc_consensus_genes=0.96, non_zero_threshold=1))
    np.random.seed(2)
    mixed, new, trg = da.get_mixed_data(mix=0.25)
    np.testing.assert_array_equal(da.gene_ids, g["t3_gene_ids"])
    np.testing.assert_array_equal(trg, g["t3_trg"])
    assert rel_fro(da.intermediate_model[1], g["t3_H"]) < 1e-3
    np.testing.assert_array_equal(da.cluster_labels, g["t3_labels"])
    np.testing.assert_allclose(mixed, g["t3_mixed"], rtol=1e-6)


@pytest.mark.parametrize("g_,n,k", [(500, 333, 6), (64, 1000, 3), (1200, 50, 12)])
def test_transfer_vs_oracle_seeded(kernels, g_, n, k):
    from scrna_b200.transfer import TransferSession
    rng = np.random.RandomState(g_ + n)
    W = np.abs(rng.randn(g_, k)) * 3
    Ht = np.abs(rng.randn(k, n))
    X = rng.poisson(W @ Ht).astype(np.float64)
    H0 = no.transfer_h_init(rng.randn(k, n))
    Href, H2ref, err_ref, n_ref = no.transfer_mu(W, X, H0)
    sess = TransferSession(W, X, kernels=kernels)
    n_it, err = sess.solve(H0)
    assert n_it == n_ref
    assert abs(err - err_ref) / err_ref < 1e-5
    assert rel_fro(sess.H_host(), Href) < 1e-3
    np.testing.assert_array_equal(sess.labels(), np.argmax(Href, axis=0))
    st = sess.reject_stats()
    np.testing.assert_allclose(st[0], X.sum(0), rtol=1e-12)
    R1 = X - W @ sess.H_host()
    np.testing.assert_allclose(st[1], np.abs(R1).sum(0), rtol=1e-9)
    np.testing.assert_allclose(st[2], (R1 ** 2).sum(0), rtol=1e-9)


def test_zero_dictionary_column_raises(kernels):
    from scrna_b200.utils import get_transferred_data_matrix
    rng = np.random.RandomState(0)
    W = np.abs(rng.randn(40, 3))
    W[:, 1] = 0.0
    X = rng.poisson(2.0, size=(40, 20)).astype(np.float64)
    with pytest.raises(Exception, match='DA target: division by zero.'):
        get_transferred_data_matrix(W, X)


def test_transferability_score_vs_reference_golden(kernels, default_sim, golden):
    """Row f4: utils.get_transferability_score (8 permuted-dictionary solves, the un-permuted solve, the target's own
    NMF) against the reference's output for the same RNG seed (tests/golden/transferability.npz)."""
    from scrna_b200.utils import get_transferability_score
    g, gs = golden("transferability"), golden("nmf_source")
    W = gs["initw_dictionary"].astype(np.float64)
    np.random.seed(4)
    score, percs, p_value = get_transferability_score(W, g["H"], default_sim["trg"], reps=8, max_iter=100)
    assert abs(score - float(g["score"])) < 1e-4
    np.testing.assert_allclose(percs, g["percs"], rtol=0, atol=1e-4)
    assert p_value == float(g["p_value"])


def test_get_mixed_data_with_transferability(kernels, default_sim, golden):
    from scrna_b200 import NmfClustering_initW, DaNmfClustering
    src = NmfClustering_initW(default_sim["src"], default_sim["gene_ids"], 7, default_sim["src_labels"])
    src.apply(k=7, alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)
    np.random.seed(1)
    da = DaNmfClustering(src, default_sim["trg"].copy(), default_sim["gene_ids"], 7)
    da.get_mixed_data(mix=0.5, calc_transferability=True, max_iter=20)
    names = [n for n, _ in da.reject]
    assert names[-3:] == ['Transfer_Percentiles', 'Transferability', 'Transferability p-value']
    assert 0.0 <= da.transferability_score <= 1.0 and da.transferability_percs.shape == (4,)
