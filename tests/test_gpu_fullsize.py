"""Size-independent properties at BASELINE.json's full NMF size (configs[2]: 20 000 genes x 100 000 cells),
where the oracle cannot run: linearity of the sweeps, agreement of the two sweep implementations, run-to-run
determinism, monotone decrease of the regularised objective under multiplicative updates, zero padding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GENES, CELLS = 20000, 100000


@pytest.fixture(scope="module")
def big(kernels):
    import bench
    torch = kernels.torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~50 GB of free HBM")
    from scrna_b200.nmf_solver import DeviceMatrix
    Xd, n = bench.synth_counts(torch, torch.device("cuda", torch.cuda.current_device()), GENES, 0, CELLS)
    return DeviceMatrix(Xd, n)


def _factors(k, seed=0):
    rng = np.random.RandomState(seed)
    return 0.3 * np.abs(rng.standard_normal((GENES, k))), 0.3 * np.abs(rng.standard_normal((k, CELLS)))


@pytest.mark.parametrize("path", ["f16", "tf32"])
@pytest.mark.parametrize("k", [8, 32])
def test_full_size_sweeps_linear_consistent_deterministic(kernels, big, k, path):
    from scrna_b200.nmf_solver import NmfSolver
    torch = kernels.torch
    tc = NmfSolver(big, k, kernels=kernels, tensor_cores=path)
    fp = NmfSolver(big, k, kernels=kernels, tensor_cores=False, XT=tc.XT)
    assert (tc.half if path == "f16" else tc.tc) and not fp.tc and not fp.half
    G1, _ = _factors(k, 1)
    G2, _ = _factors(k, 2)
    a = tc.xt_times(G1)
    # the tensor-core and the FP32-pipe sweep are independent implementations of the same product
    b = fp.xt_times(G1)
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 2e-5
    # linearity: X^T (G1 + G2) = X^T G1 + X^T G2
    c = tc.xt_times(G2)
    s = tc.xt_times(G1 + G2)
    assert np.linalg.norm(s - (a + c)) / np.linalg.norm(s) < 2e-5
    # fixed-order reductions: bitwise reproducible
    assert np.array_equal(tc.xt_times(G1), a)
    # a column sample against fp64 on the host (X is integer counts: exact in fp32)
    cols = np.arange(0, CELLS, 997)[:64]
    Xs = big.data[:, torch.from_numpy(cols).to(big.data.device)].double().cpu().numpy()
    ref = Xs.T @ G1
    assert np.linalg.norm(a[cols] - ref) / np.linalg.norm(ref) < 1e-5
    # the other orientation (X . C^T through the transposed copy)
    _, C1 = _factors(k, 3)
    xa, xb = tc.x_times(C1.T), fp.x_times(C1.T)
    assert np.linalg.norm(xa - xb) / np.linalg.norm(xb) < 2e-5
    # padding components / cells never leak
    assert float(tc.NB[k:].abs().max().cpu()) == 0.0 if tc.kp > k else True


def test_full_size_mu_objective_decreases(kernels, big):
    """Lee-Seung multiplicative updates never increase the (regularised Frobenius) objective: checked on the
    whole 20k x 100k problem over 6 iterations through the public solver API."""
    from scrna_b200.nmf_solver import NmfSolver
    k, l1, l2 = 8, 7.5, 2.5
    slv = NmfSolver(big, k, kernels=kernels)
    G0, C0 = _factors(k, 5)
    slv.prepare(G0, C0, solver="mu", l1_G=l1, l2_G=l2, l1_C=l1, l2_C=l2, tol=0.0, max_iter=10 ** 6)

    def objective():
        fro = slv.frobenius_error()
        G = slv.G64[:k].double()
        C = slv.C64[:k].double()
        reg = l1 * (float(G.sum().cpu()) + float(C.sum().cpu())) + 0.5 * l2 * (float((G * G).sum().cpu()) +
                                                                         float((C * C).sum().cpu()))
        return 0.5 * fro * fro + reg

    prev = objective()
    for _ in range(6):
        slv.step()
        cur = objective()
        assert cur <= prev * (1 + 1e-9), (cur, prev)
        prev = cur
    assert np.isfinite(prev)
