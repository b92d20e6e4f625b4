"""World-size-2 test of the multi-GPU host logic on CPU (gloo): cell sharding, all-reduce payload,
stop rule agreement.  Compute is a test-only stand-in (tests/fake_kernels.py); the CUDA kernels
themselves are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_kernels import FakeKernels
        from oracle import nmf_oracle as no
        from scrna_b200.nmf_solver import DeviceMatrix, NmfSolver, TorchDistComm, shard_bounds
        rng = np.random.RandomState(7)
        g, n, k = 60, 101, 4
        mu = rng.gamma(2.0, 1.0, size=(g, 1)) * (1 + (rng.rand(g, 3) < 0.3) @ (rng.rand(3, n) < 0.4) * 3.0)
        X = rng.poisson(mu).astype(np.float64)
        W0, H0 = no.nndsvdar_init(X, k)
        c0, c1 = shard_bounds(n, world, rank)
        kn = FakeKernels()
        dm = DeviceMatrix.from_host(X[:, c0:c1], kn, device=torch.device("cpu"))
        slv = NmfSolver(dm, k, kernels=kn, comm=TorchDistComm())
        solver, first = case
        if solver == "cd":
            if first == "G":
                Wr, Hr, nr, _ = no.nmf_cd(X, W0, H0, 0.5, 0.5, 0.1, 0.1, tol=1e-3, max_iter=80)
                res = slv.fit(W0, H0[:, c0:c1], solver="cd", l1_G=0.5, l2_G=0.1, l1_C=0.5, l2_C=0.1, tol=1e-3,
                              max_iter=80, first="G", poll=3)
                G, C = res.G, res.C
            else:  # factorise X^T: sklearn W = C^T (updated first), H = G^T
                Wt, Ht, nr, _ = no.nmf_cd(X.T, H0.T.copy(), W0.T.copy(), 0.5, 0.5, 0.1, 0.1, tol=1e-3, max_iter=80)
                Wr, Hr = Ht.T, Wt.T
                res = slv.fit(W0, H0[:, c0:c1], solver="cd", l1_G=0.5, l2_G=0.1, l1_C=0.5, l2_C=0.1, tol=1e-3,
                              max_iter=80, first="C", poll=3)
                G, C = res.G, res.C
        else:
            Wr, Hr, nr, _ = no.nmf_mu(X, W0, H0, 0.5, 0.5, 0.1, 0.1, tol=1e-4, max_iter=60)
            res = slv.fit(W0, H0[:, c0:c1], solver="mu", l1_G=0.5, l2_G=0.1, l1_C=0.5, l2_C=0.1, tol=1e-4,
                          max_iter=60)
            G, C = res.G, res.C
        ok = (res.n_iter == nr
              and np.linalg.norm(G - Wr) / np.linalg.norm(Wr) < 1e-4
              and np.linalg.norm(C - Hr[:, c0:c1]) / np.linalg.norm(Hr[:, c0:c1]) < 1e-4)
        ret[rank] = (bool(ok), int(res.n_iter), int(nr))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [("cd", "G"), ("cd", "C"), ("mu", "G")])
def test_sharded_fit_world2(case):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, n_iter, n_ref = ret[r]
        assert ok, (r, n_iter, n_ref)
    assert ret[0][1] == ret[1][1]


def test_shard_bounds_cover_and_balance():
    from scrna_b200.nmf_solver import shard_bounds
    for n in (1, 7, 100, 100000, 100003):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scrna_b200.nmf_solver import TorchDistComm, shard_bounds
        n = 37
        full = torch.arange(n * n, dtype=torch.float64).reshape(n, n)
        mine = torch.full((n, n), -1.0, dtype=torch.float64)
        bounds = [shard_bounds(n, world, r) for r in range(world)]
        r0, r1 = bounds[rank]
        mine[r0:r1] = full[r0:r1]            # this rank's row block of the distance matrix
        TorchDistComm().all_gather_rows(mine, bounds)
        ret[rank] = bool(torch.equal(mine, full))
    finally:
        dist.destroy_process_group()


def test_sc3_distance_row_blocks_all_gather_world2():
    """SC3 distances shard by row blocks: the all-gather that assembles D on every rank."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret[r] for r in range(world)] == [True, True]


def _sc3_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.cluster.hierarchy as spc
        from scipy.spatial.distance import pdist, squareform
        from scrna_b200.nmf_solver import TorchDistComm
        from scrna_b200.sc3_clustering import SC3Clustering
        rng = np.random.RandomState(3)
        X = np.concatenate([rng.randn(30, 20) + 4 * rng.randn(30, 1) for _ in range(3)], axis=1)   # 30 genes x 60 cells

        def dist_fn(data, gene_ids):
            return squareform(pdist(data.T))

        def dimred(D):
            return None, np.linalg.svd(D - D.mean(axis=0))[2].T[:, :6]

        def cluster(V, _skip=False):
            # stand-in k-means (test infrastructure): consumes the global RNG like a real fit would
            u = np.random.random_sample(3)
            if _skip:
                return None
            return (V[:, 0] > np.quantile(V[:, 0], 0.3 + 0.4 * u[0])).astype(int) + (V[:, -1] > 0).astype(int)
        cluster.supports_skip = True

        def consensus(labels):
            l = np.asarray(labels).reshape(-1)
            return (l[:, None] == l[None, :]).astype(float)

        def final(C):
            Z = spc.complete(pdist(C))
            return spc.cut_tree(Z, n_clusters=3).reshape(-1), squareform(pdist(C))

        def run(comm):
            np.random.seed(11)
            sc3 = SC3Clustering(X, np.arange(30), pc_range=[2, 6], sub_sample=True, consensus_mode=0, comm=comm)
            sc3.add_distance_calculation(dist_fn)
            sc3.add_dimred_calculation(dimred)
            sc3.add_intermediate_clustering(cluster)
            sc3.set_build_consensus_matrix(consensus)
            sc3.set_consensus_clustering(final)
            sc3.apply()
            return sc3.cluster_labels, sc3.dists, np.random.random_sample()

        ref_labels, ref_dists, ref_next = run(None)
        got_labels, got_dists, got_next = run(TorchDistComm())
        ret[rank] = bool(np.array_equal(ref_labels, got_labels) and np.array_equal(ref_dists, got_dists)
                         and ref_next == got_next)
    finally:
        dist.destroy_process_group()


def test_sc3_kmeans_ensemble_sharded_world2():
    """The k-means fits of the SC3 ensemble dealt over two ranks: same consensus, same labels and the same position
    on the global RNG stream as the single-process run (skipped fits still consume their draws)."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sc3_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret[r] for r in range(world)] == [True, True]


def _class_worker(rank, world, port, case, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_kernels import FakeKernels
        from oracle import nmf_oracle as no
        import scrna_b200.nmf_clustering as nc
        from scrna_b200.nmf_solver import DeviceMatrix, NmfSolver, TorchDistComm, shard_bounds
        kn = FakeKernels()
        nc._kernels = lambda: kn                       # the classes' compute provider (test stand-in)
        comm = TorchDistComm()
        rng = np.random.RandomState(5)
        g, n, k = 50, 91, 4
        lab = rng.randint(0, k, size=n)
        lab[:k] = np.arange(k)
        mu = rng.gamma(2.0, 1.0, size=(g, k))[:, lab] * (0.5 + rng.rand(1, n))
        X = rng.poisson(mu).astype(np.float64) + 1.0 * (rng.rand(g, n) < 0.05)
        ids = np.arange(g)
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        if case == "nmf":
            Wr, Hr, lr, nr = no.nmf_clustering_apply(X, k, alpha=2.0, l1=0.5, max_iter=60, rel_err=1e-3)
            obj = nc.NmfClustering(X, ids, k)
            obj.apply(k=k, alpha=2.0, l1=0.5, max_iter=60, rel_err=1e-3, comm=comm)
            ok = (obj.n_iter == nr and rel(obj.dictionary, Wr) < 1e-4 and rel(obj.data_matrix, Hr) < 1e-4
                  and np.array_equal(obj.cluster_labels, lr) and obj.data_matrix.shape == (k, n))
        elif case == "nmf_sharded_device_init":
            Wr, Hr, lr, nr = no.nmf_clustering_apply(X, k, alpha=2.0, l1=0.5, max_iter=60, rel_err=1e-3)
            c0, c1 = shard_bounds(n, world, rank)
            obj = nc.NmfClustering(X[:, c0:c1], ids, k)          # this rank only ever sees its cells
            obj.apply(k=k, alpha=2.0, l1=0.5, max_iter=60, rel_err=1e-3, comm=comm, cells="sharded", init="device")
            ok = (abs(obj.n_iter - nr) <= 1 and rel(obj.dictionary, Wr) < 1e-3
                  and rel(obj.data_matrix, Hr[:, c0:c1]) < 1e-3 and np.array_equal(obj.cluster_labels, lr[c0:c1]))
        elif case == "initw":
            Dr, Mr, lr, nr = no.nmf_clustering_initw_apply(X, lab, k, alpha=2.0, l1=0.5, max_iter=60, rel_err=1e-3)
            obj = nc.NmfClustering_initW(X, ids, k, lab)
            obj.apply(k=k, alpha=2.0, l1=0.5, max_iter=60, rel_err=1e-3, comm=comm)
            ok = (tuple(obj.n_iter) == tuple(nr) and rel(obj.dictionary, Dr) < 1e-4 and rel(obj.data_matrix, Mr) < 1e-4
                  and np.array_equal(obj.cluster_labels, lr))
        else:   # fixed dictionary, cell factor alone updated, cells sharded (the violation was never
            #     all-reduced on this path and the fit stopped after one iteration)
            W = np.abs(rng.randn(g, k)) + 0.1
            Ht, _, nr, _ = no.nmf_cd(X.T, np.zeros((n, k)), W.T.copy(), 0, 0, 0, 0, tol=1e-4, max_iter=80, update_H=False)
            c0, c1 = shard_bounds(n, world, rank)
            slv = NmfSolver(DeviceMatrix.from_host(X[:, c0:c1], kn), k, kernels=kn, comm=comm)
            res = slv.fit(W, np.zeros((k, c1 - c0)), solver="cd", tol=1e-4, max_iter=80, update_G=False, update_C=True,
                          first="C", seed=0, shuffle=True)
            ok = res.n_iter == nr and nr > 2 and rel(res.C, Ht.T[:, c0:c1]) < 1e-6
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["nmf", "nmf_sharded_device_init", "initw", "fixed_dictionary_cd"])
def test_class_api_cells_sharded_world2(case):
    """NmfClustering / NmfClustering_initW .apply(comm=) with the cells sharded over two ranks (replicated and
    pre-sharded input, host and device NNDSVDar init) against the single-process oracle, and the fixed-dictionary
    sharded CD solve."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_class_worker, args=(world, port, case, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == [True, True]
