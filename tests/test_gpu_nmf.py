"""Parity of the CUDA NMF path (through the C ABI) against the oracle and the reference goldens.

Tolerances are BASELINE.json's: relative Frobenius reconstruction error 1e-4, per-factor relative
error 1e-3, identical cluster labels.  The sweeps accumulate in fp32, so kernel-level products are
compared at 2e-5 relative to the fp64 product.
"""
import numpy as np
import pytest

from conftest import rel_fro
from oracle import nmf_oracle as no

pytestmark = pytest.mark.gpu

CLI = dict(alpha=10., l1=0.75, max_iter=4000, rel_err=1e-3)
TOL_FACTOR = 1e-3
TOL_RECON = 1e-4


def _recon_err(X, W, H):
    return np.linalg.norm(X - W @ H) / np.linalg.norm(X)


def _solver(kernels, X, k, transpose_copy=True, tensor_cores="auto"):
    from scrna_b200.nmf_solver import DeviceMatrix, NmfSolver
    return NmfSolver(DeviceMatrix.from_host(X, kernels), k, kernels=kernels, transpose_copy=transpose_copy,
                     tensor_cores=tensor_cores)


@pytest.fixture(params=["simt", "tcgen05", "f16"])
def sweep_path(request, monkeypatch):
    """Runs a test once on the FP32-pipe sweeps, once on the 3xTF32 tcgen05 sweep and once on the 16-bit-storage
    tcgen05 sweep (SCRNA_B200_TC override of the solver's automatic choice; the classes construct their solver
    internally; 'f16' falls back to 3xTF32 for matrices that are not fp16-exact, e.g. the mixed target data)."""
    monkeypatch.setenv("SCRNA_B200_TC", {"simt": "0", "tcgen05": "tf32", "f16": "f16"}[request.param])
    return request.param


@pytest.mark.parametrize("parts", [2, 3])
@pytest.mark.parametrize("g,n,k", [(1000, 1000, 8), (333, 517, 7), (64, 4100, 5), (2500, 130, 16), (129, 257, 33),
                                   (40, 36, 3), (700, 900, 50), (5, 7, 2), (3000, 2000, 64), (1037, 1500, 17),
                                   (2100, 1300, 24), (900, 700, 40)])
def test_f16_sweeps_match_fp64_products(kernels, g, n, k, parts, monkeypatch):
    """The 16-bit sweep (fp16 storage of X and X^T staged by swizzled TMA boxes straight into tcgen05.mma kind::f16,
    two- or three-part scaled fp16 split of the factor) on both resident copies, against fp64 products; counts up to
    fp16's exact-integer limit 2048 and factors with a wide dynamic range."""
    torch = kernels.torch
    monkeypatch.setenv("SCRNA_B200_H_PARTS", str(parts))
    rng = np.random.RandomState(g + 3 * n + k)
    X = rng.poisson(2.0, size=(g, n)).astype(np.float64)
    X[rng.rand(g, n) < 0.01] = 2048.0
    X[rng.rand(g, n) < 0.01] = 1777.0
    G = (np.abs(rng.randn(g, k)) + 0.01) * np.exp(3.0 * rng.randn(1, k))
    C = (np.abs(rng.randn(k, n)) + 0.01) * np.exp(2.0 * rng.randn(k, n))
    slv = _solver(kernels, X, k, tensor_cores="f16")
    assert slv.half and not slv.tc and slv.kp % 8 == 0 and slv.Xs.is_half and slv.XT.is_half and slv.h_parts == parts
    np.testing.assert_array_equal(slv.Xs.dense().float().cpu().numpy(), X.astype(np.float32))
    np.testing.assert_array_equal(slv.XT.dense().float().cpu().numpy(), X.T.astype(np.float32))
    # padding rows / columns of the panel-major storage are zero
    assert float(slv.XT.data.double().sum().cpu()) == float(X.sum()) == float(slv.Xs.data.double().sum().cpu())
    slv._load_factors(G, C)
    slv._g_num_from_R = True
    slv._products_for_G(None, with_gram=True)
    torch.cuda.synchronize()
    xht = slv.R_xht.view(slv.kp, slv.ldg)[:k, :g].cpu().numpy().T
    assert rel_fro(xht, X @ C.T) < 4e-6     # the tensor core truncates its fp32 accumulator (see nmf_sweep_tc.cu)
    assert np.all(slv.R_xht.view(slv.kp, slv.ldg)[k:, :g].cpu().numpy() == 0)
    slv._sweep_X(None)
    kernels.reduce_cols(slv.PB, slv.SB, slv.kp, slv.ldn, slv.X.ncols, slv.NB, None)
    torch.cuda.synchronize()
    wtx = slv.NB[:k, :n].cpu().numpy()
    ref = G.T @ X
    assert rel_fro(wtx, ref) < 4e-6
    assert np.all(slv.NB[k:, :n].cpu().numpy() == 0)
    # elementwise and per component: no tile of the mapping is misplaced, no component loses accuracy to its scale
    for t in range(k):
        assert np.max(np.abs(wtx[t] - ref[t])) < 1e-5 * np.abs(ref[t]).max(), t
    # residual over the fp16 storage
    Gs, Cs = slv.Gs32.cpu().numpy().astype(np.float64)[:, :k], slv.Ct32.cpu().numpy().astype(np.float64)[:k, :n]
    assert abs(slv.frobenius_error() - np.linalg.norm(X - Gs @ Cs)) / np.linalg.norm(X - Gs @ Cs) < 2e-5


def test_f16_not_chosen_for_inexact_data(kernels):
    rng = np.random.RandomState(0)
    X = rng.poisson(2.0, size=(200, 300)).astype(np.float64)
    assert _solver(kernels, X, 6).half
    X[3, 4] = 2049.0                                      # not an fp16 value
    assert not _solver(kernels, X, 6).half
    X[3, 4] = 0.1
    assert not _solver(kernels, X, 6).half
    with pytest.raises(ValueError):
        _solver(kernels, X, 6, tensor_cores="f16")


def test_update_writes_f16_operand_shadow(kernels):
    """The epilogue's fp16 operand shadow (three parts here) equals the one scrna_nmf_h_shadow builds from the master, the scales are
    local to 128-row groups (max|entry| * 2^e in [2^13, 2^14)), and the parts reconstruct every entry to max(2^-33 of
    the entry, 2^-38 of the group's largest entry): 33 significant bits, floored by fp16's subnormal spacing."""
    import scrna_b200._lib as L
    torch = kernels.torch
    rng = np.random.RandomState(5)
    rows, k, kq, ld = 1001, 13, 16, 1024
    dev = torch.device("cuda")
    F = np.zeros((kq, ld))
    F[:k, :rows] = np.abs(rng.randn(k, rows)) * np.exp(4.0 * rng.randn(k, 1)) * np.exp(3.0 * rng.randn(1, rows))
    F64 = torch.from_numpy(F).to(dev)
    N = np.zeros((kq, ld)); N[:k, :rows] = np.abs(rng.randn(k, rows)) * 10
    A = np.abs(rng.randn(40, k)); Gm = np.zeros((kq, kq)); Gm[:k, :k] = A.T @ A
    n_el = kernels.h_shadow_elems(rows, kq, 3)
    nb = kernels.h_nb(kq, 3)
    groups = kernels.h_groups(rows)
    sh_a = torch.zeros(n_el, dtype=torch.float16, device=dev)
    sh_b = torch.full((n_el,), -1.0, dtype=torch.float16, device=dev)
    inv_a = torch.zeros(groups * kq, dtype=torch.float32, device=dev)
    inv_b = torch.zeros(groups * kq, dtype=torch.float32, device=dev)
    S32 = torch.zeros((rows, kq), dtype=torch.float32, device=dev)
    nblk = kernels.update_blocks(rows)
    viol = torch.zeros(nblk, dtype=torch.float64, device=dev)
    gpart = torch.zeros(nblk * kq * kq, dtype=torch.float64, device=dev)
    gout = torch.zeros(kq * kq, dtype=torch.float64, device=dev)
    ctrl = torch.zeros(16, dtype=torch.int32, device=dev)
    kernels.update(L.SOLVER_MU, F64, ld, rows, k, kq, torch.from_numpy(N).to(dev), torch.from_numpy(Gm).to(dev), 0.5, 0.1,
                   None, 1, 0, S32, None, viol, ctrl, gram_partials=gpart, gram_out=gout, Fh=sh_a, inv_scale=inv_a, kq=kq,
                   parts=3)
    torch.cuda.synchronize()
    got = F64.cpu().numpy()[:k, :rows]
    np.testing.assert_allclose(gout.cpu().numpy().reshape(kq, kq)[:k, :k], got @ got.T, rtol=1e-12)
    kernels.h_shadow(F64, ld, rows, k, kq, 3, gout, 1, kq, None, sh_b, inv_b)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(sh_a.cpu().numpy(), sh_b.cpu().numpy())
    np.testing.assert_array_equal(inv_a.cpu().numpy(), inv_b.cpu().numpy())
    sh = sh_b.cpu().numpy().astype(np.float64).reshape(-1, nb, 8)             # [(r/8)][n][r%8]
    inv = inv_b.cpu().numpy().astype(np.float64).reshape(groups, kq)
    assert np.all(np.isfinite(sh)) and np.abs(sh).max() < 2.0 ** 14
    parts = [sh[:, p * kq:(p + 1) * kq, :].transpose(1, 0, 2).reshape(kq, -1)[:k, :rows] for p in range(3)]
    grp = np.arange(rows) // 128
    rec = (parts[0] + parts[1] + parts[2]) * inv[grp, :k].T
    gmax = np.stack([np.abs(got[:, grp == q]).max(axis=1) for q in range(groups)], axis=1)[:, grp]   # k x rows
    assert np.max(np.abs(rec - got) / gmax) <= 2.0 ** -31
    assert np.all(np.abs(parts[0]).reshape(k, -1) * 0 == 0)
    lead = np.stack([np.abs(parts[0][:, grp == q]).max(axis=1) for q in range(groups)], axis=1)
    assert np.all(lead >= 2.0 ** 13) and np.all(lead <= 2.0 ** 14)            # every group uses fp16's top binades
    assert np.all(sh[:, 3 * kq:, :] == 0) and np.all(sh.reshape(-1, nb, 8)[:, k:kq, :] == 0)


@pytest.mark.parametrize("exact_x", [True, False])
@pytest.mark.parametrize("g,n,k", [(1000, 1000, 8), (333, 517, 7), (64, 4100, 5), (2500, 130, 16), (129, 257, 33),
                                   (40, 36, 3), (700, 900, 50), (5, 7, 2), (3000, 2000, 64), (1037, 1500, 17)])
def test_tcgen05_sweeps_match_fp64_products(kernels, g, n, k, exact_x):
    """The tensor-core sweep (3xTF32 split, TMEM accumulators drained every 512 rows) on both resident copies
    of X, against fp64 products; counts (TF32-exact, single X pass) and wide-dynamic-range reals."""
    torch = kernels.torch
    rng = np.random.RandomState(g + 3 * n + k)
    if exact_x:
        X = rng.poisson(2.0, size=(g, n)).astype(np.float64)
    else:
        X = rng.rand(g, n) * np.exp(2.0 * rng.randn(g, n))
    G = np.abs(rng.randn(g, k)) + 0.01
    C = np.abs(rng.randn(k, n)) + 0.01
    slv = _solver(kernels, X, k, tensor_cores="tf32")
    assert slv.tc and slv.kp % 16 == 0 and (slv.tc_flags == 32) == exact_x
    slv._load_factors(G, C)
    slv._g_num_from_R = True
    slv._products_for_G(None, with_gram=True)
    torch.cuda.synchronize()
    X32 = X.astype(np.float32).astype(np.float64)
    xht = slv.R_xht.view(slv.kp, slv.ldg)[:k, :g].cpu().numpy().T
    assert rel_fro(xht, X32 @ C.T) < 1e-5
    assert np.all(slv.R_xht.view(slv.kp, slv.ldg)[k:, :g].cpu().numpy() == 0)
    kernels.wtx_partial_tc(slv.X.data, slv.X.ldx, g, slv.X.ncols, slv.Gtc, slv.kp, slv.PB, slv.ldn, slv.SB, None,
                           flags=slv.tc_flags)
    kernels.reduce_cols(slv.PB, slv.SB, slv.kp, slv.ldn, slv.X.ncols, slv.NB, None)
    torch.cuda.synchronize()
    wtx = slv.NB[:k, :n].cpu().numpy()
    assert rel_fro(wtx, G.T @ X32) < 1e-5
    assert np.all(slv.NB[k:, :n].cpu().numpy() == 0)
    # elementwise: no column/row of the tile mapping is misplaced
    ref = G.T @ X32
    assert np.max(np.abs(wtx - ref) / (np.abs(ref) + 1e-3 * np.abs(ref).max())) < 1e-4


def test_update_writes_tcgen05_operand_shadow(kernels):
    """The epilogue's [hi | lo] operand shadow equals the one scrna_nmf_tc_shadow builds from the master."""
    import scrna_b200._lib as L
    torch = kernels.torch
    rng = np.random.RandomState(5)
    rows, k, kp, ld = 1001, 13, 16, 1024
    dev = torch.device("cuda")
    F = np.zeros((kp, ld)); F[:k, :rows] = np.abs(rng.randn(k, rows))
    F64 = torch.from_numpy(F).to(dev)
    N = np.zeros((kp, ld)); N[:k, :rows] = np.abs(rng.randn(k, rows)) * 10
    A = np.abs(rng.randn(40, k)); Gm = np.zeros((kp, kp)); Gm[:k, :k] = A.T @ A
    n_el = kernels.tc_shadow_elems(rows, kp)
    tc_a = torch.zeros(n_el, dtype=torch.float32, device=dev)
    tc_b = torch.full((n_el,), -1.0, dtype=torch.float32, device=dev)
    S32 = torch.zeros((rows, kp), dtype=torch.float32, device=dev)
    viol = torch.zeros(kernels.update_blocks(rows), dtype=torch.float64, device=dev)
    ctrl = torch.zeros(16, dtype=torch.int32, device=dev)
    kernels.update(L.SOLVER_MU, F64, ld, rows, k, kp, torch.from_numpy(N).to(dev), torch.from_numpy(Gm).to(dev), 0.5, 0.1,
                   None, 1, 0, S32, None, viol, ctrl, TC32=tc_a)
    kernels.tc_shadow(F64, ld, rows, k, kp, tc_b)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(tc_a.cpu().numpy(), tc_b.cpu().numpy())
    sh = tc_b.cpu().numpy().reshape(-1, 2 * kp, 4)          # [(r/4)][n][r%4]
    got = F64.cpu().numpy()[:k, :rows]
    hi = sh[:, :kp, :].transpose(1, 0, 2).reshape(kp, -1)[:k, :rows].astype(np.float64)
    lo = sh[:, kp:, :].transpose(1, 0, 2).reshape(kp, -1)[:k, :rows].astype(np.float64)
    assert np.max(np.abs(hi + lo - got) / got.max()) < 2.0 ** -20
    assert np.all((hi.astype(np.float32).view(np.uint32) & 0x1fff) == 0)


@pytest.mark.parametrize("transpose_copy", [True, False])
@pytest.mark.parametrize("g,n,k", [(1000, 1000, 8), (333, 517, 7), (64, 4100, 5), (2500, 130, 16),
                                   (129, 257, 33), (40, 36, 3), (700, 900, 50), (5, 7, 2)])
def test_sweeps_match_fp64_products(kernels, g, n, k, transpose_copy):
    torch = kernels.torch
    rng = np.random.RandomState(g + n + k)
    X = rng.poisson(2.0, size=(g, n)).astype(np.float64)
    G = np.abs(rng.randn(g, k))
    C = np.abs(rng.randn(k, n))
    slv = _solver(kernels, X, k, transpose_copy, tensor_cores=False)
    slv._load_factors(G, C)
    if transpose_copy:
        np.testing.assert_array_equal(slv.XT.data[:n, :g].cpu().numpy(), X.T.astype(np.float32))
        assert float(slv.XT.data[:, g:].abs().sum().cpu()) == 0.0
    slv._g_num_from_R = True
    slv._products_for_G(None, with_gram=True)
    torch.cuda.synchronize()
    xht = slv.R_xht.view(slv.kp, slv.ldg)[:k, :g].cpu().numpy().T
    cct = slv.R_cct.view(slv.kp, slv.kp)[:k, :k].cpu().numpy()
    assert rel_fro(xht, X @ C.T) < 2e-5
    assert rel_fro(cct, C @ C.T) < 1e-12
    assert np.all(slv.R_xht.view(slv.kp, slv.ldg)[k:, :g].cpu().numpy() == 0)
    kernels.gram(slv.G64, 1, slv.ldg, g, k, slv.kp, slv.gram_part, slv.gtg, None)
    kernels.wtx_partial(slv.X.data, slv.X.ldx, g, slv.X.ncols, slv.Gs32, slv.kp, slv.PB, slv.ldn, slv.SB, None)
    kernels.reduce_cols(slv.PB, slv.SB, slv.kp, slv.ldn, slv.X.ncols, slv.NB, None)
    torch.cuda.synchronize()
    wtx = slv.NB[:k, :n].cpu().numpy()
    gtg = slv.gtg.view(slv.kp, slv.kp)[:k, :k].cpu().numpy()
    assert rel_fro(wtx, G.T @ X) < 2e-5
    assert rel_fro(gtg, G.T @ G) < 1e-12
    # residuals
    kernels.residual(slv.X.data, slv.X.ldx, g, slv.X.ncols, slv.Gs32, slv.Ct32, slv.ldn, slv.kp, 1, slv.res_part,
                     slv.res_out)
    l1 = float(slv.res_out.cpu()[0])
    assert abs(l1 - np.abs(X - G @ C).sum()) / np.abs(X - G @ C).sum() < 2e-5
    assert abs(slv.frobenius_error() - np.linalg.norm(X - G @ C)) / np.linalg.norm(X - G @ C) < 2e-5


@pytest.mark.parametrize("fused_parts", [False, True])
@pytest.mark.parametrize("solver,rank", [("cd", (7, 8)), ("mu", (7, 8)), ("mu", (50, 64)), ("cd", (50, 64)), ("mu", (33, 48))])
def test_single_update_matches_oracle(kernels, solver, fused_parts, rank):
    """One half step with exact fp64 inputs: the epilogue alone, compared at fp64 rounding level (row-per-thread
    kernel for CD and MU up to rank 32, tiled kernel for MU above)."""
    import scrna_b200._lib as L
    torch = kernels.torch
    rng = np.random.RandomState(11)
    rows, (k, kp), ld = 777, rank, 800
    W = np.abs(rng.randn(rows, k))
    W[rng.rand(rows, k) < 0.25] = 0.0
    A = np.abs(rng.randn(50, k))
    HHt = A.T @ A
    XHt = np.abs(rng.randn(rows, k)) * 30
    perm = rng.permutation(k).astype(np.int32)
    l1, l2 = 7.5, 2.5
    dev = torch.device("cuda")
    F = np.zeros((kp, ld)); F[:k, :rows] = W.T
    N = np.zeros((kp, ld)); N[:k, :rows] = XHt.T
    F64 = torch.from_numpy(F).to(dev)
    S32 = torch.full((rows, kp), -1.0, dtype=torch.float32, device=dev)
    T32 = torch.zeros((kp, ld), dtype=torch.float32, device=dev)
    num = torch.from_numpy(N).to(dev)
    Gm = np.zeros((kp, kp)); Gm[:k, :k] = HHt
    gram = torch.from_numpy(Gm).to(dev)
    pp = np.zeros((1, kp), dtype=np.int32); pp[0, :k] = perm
    perms = torch.from_numpy(pp).to(dev)
    viol = torch.zeros(kernels.update_blocks(rows), dtype=torch.float64, device=dev)
    ctrl = torch.zeros(16, dtype=torch.int32, device=dev)
    code = L.SOLVER_CD if solver == "cd" else L.SOLVER_MU
    gpart = torch.zeros(kernels.update_blocks(rows) * kp * kp, dtype=torch.float64, device=dev)
    gout = torch.zeros(kp * kp, dtype=torch.float64, device=dev)
    if fused_parts:
        # the same numerator presented as 3 fp32 split partials (exactly representable pieces)
        N32 = np.round(N * 0.25 * 64) / 64
        parts = np.stack([N32, N32, N - 2 * N32]).astype(np.float32)
        N = parts.astype(np.float64).sum(axis=0)
        XHt = N[:k, :rows].T.copy()
        kernels.update(code, F64, ld, rows, k, kp, None, gram, l1, l2, perms, 1, 0, S32, T32, viol, ctrl,
                       num_parts=torch.from_numpy(parts).to(dev), nsplit=3, gram_partials=gpart, gram_out=gout)
    else:
        kernels.update(code, F64, ld, rows, k, kp, num, gram, l1, l2, perms, 1, 0, S32, T32, viol, ctrl,
                       gram_partials=gpart, gram_out=gout)
    torch.cuda.synchronize()
    got = F64.cpu().numpy()[:k, :rows].T
    np.testing.assert_allclose(gout.view(kp, kp)[:k, :k].cpu().numpy(), got.T @ got, rtol=1e-12)
    Wref = W.copy()
    if solver == "cd":
        H2 = HHt.copy(); H2.flat[:: k + 1] += l2
        v = no.cd_sweep(Wref, H2, XHt - l1, perm)
        assert abs(float(viol.sum().cpu()) - v) <= 1e-11 * v
        np.testing.assert_array_equal(got == 0, Wref == 0)
    else:
        den = Wref @ HHt + l1 + l2 * Wref
        den[den == 0] = no.EPSILON
        Wref = Wref * (XHt / den)
    np.testing.assert_allclose(got, Wref, rtol=1e-12, atol=1e-300)
    s32 = S32.cpu().numpy()
    np.testing.assert_array_equal(s32[:, :k], got.astype(np.float32))
    assert np.all(s32[:, k:] == 0)
    np.testing.assert_array_equal(T32.cpu().numpy()[:k, :rows].T, got.astype(np.float32))


def test_nmf_clustering_default_data_vs_reference_golden(kernels, default_sim, golden, sweep_path):
    from scrna_b200 import NmfClustering
    g = golden("nmf_source")
    nmf = NmfClustering(default_sim["src"], default_sim["gene_ids"], 8, default_sim["src_labels"])
    nmf.apply(k=8, **CLI)
    X = default_sim["src"]
    assert nmf.n_iter == int(g["nmf_n_iter"])
    assert rel_fro(nmf.dictionary, g["nmf_W"]) < TOL_FACTOR
    assert rel_fro(nmf.data_matrix, g["nmf_H"]) < TOL_FACTOR
    e_ref = _recon_err(X, g["nmf_W"], g["nmf_H"])
    e_gpu = _recon_err(X, nmf.dictionary, nmf.data_matrix)
    assert abs(e_gpu - e_ref) / e_ref < TOL_RECON
    np.testing.assert_array_equal(nmf.cluster_labels, g["nmf_labels"])


def test_initw_default_data_vs_reference_golden(kernels, default_sim, golden, sweep_path):
    from scrna_b200 import NmfClustering_initW
    g = golden("nmf_source")
    nmf = NmfClustering_initW(default_sim["src"], default_sim["gene_ids"], 7, default_sim["src_labels"])
    nmf.apply(k=7, **CLI)
    assert nmf.n_iter == (2, 163)
    assert rel_fro(nmf.dictionary, g["initw_dictionary"]) < TOL_FACTOR
    assert rel_fro(nmf.data_matrix, g["initw_data_matrix"]) < TOL_FACTOR
    np.testing.assert_array_equal(nmf.cluster_labels, g["initw_labels"])


def test_mu_solver_vs_sklearn_golden(kernels, default_sim, golden, sweep_path):
    g = golden("nmf_source")
    slv = _solver(kernels, default_sim["src"], 8)
    res = slv.fit(g["W0"], g["H0"], solver="mu", l1_G=7.5, l2_G=2.5, l1_C=7.5, l2_C=2.5, tol=1e-4, max_iter=200)
    assert res.n_iter == int(g["mu_n_iter"])
    assert rel_fro(res.G, g["mu_W"]) < TOL_FACTOR
    assert rel_fro(res.C, g["mu_H"]) < TOL_FACTOR
    X = default_sim["src"]
    e_ref = _recon_err(X, g["mu_W"], g["mu_H"])
    assert abs(_recon_err(X, res.G, res.C) - e_ref) / e_ref < TOL_RECON


@pytest.mark.parametrize("g,n,k,solver", [(300, 220, 5, "cd"), (257, 1030, 12, "cd"), (300, 220, 5, "mu"),
                                          (120, 333, 20, "mu"), (90, 75, 3, "cd")])
def test_fit_vs_oracle_seeded(kernels, g, n, k, solver, sweep_path):
    rng = np.random.RandomState(g * 7 + n)
    mu = rng.gamma(2.0, 1.0, size=(g, 1)) * (1 + (rng.rand(g, 6) < 0.2) @ (rng.rand(6, n) < 0.3) * 3.0)
    X = rng.poisson(mu).astype(np.float64)
    W0, H0 = no.nndsvdar_init(X, k)
    l1, l2 = 0.75, 0.25
    slv = _solver(kernels, X, k)
    if solver == "cd":
        Wr, Hr, nr, _ = no.nmf_cd(X, W0, H0, l1, l1, l2, l2, tol=1e-3, max_iter=150)
        res = slv.fit(W0, H0, solver="cd", l1_G=l1, l2_G=l2, l1_C=l1, l2_C=l2, tol=1e-3, max_iter=150)
    else:
        Wr, Hr, nr, _ = no.nmf_mu(X, W0, H0, l1, l1, l2, l2, tol=1e-4, max_iter=150)
        res = slv.fit(W0, H0, solver="mu", l1_G=l1, l2_G=l2, l1_C=l1, l2_C=l2, tol=1e-4, max_iter=150)
    assert res.n_iter == nr
    assert rel_fro(res.G, Wr) < TOL_FACTOR
    assert rel_fro(res.C, Hr) < TOL_FACTOR
    np.testing.assert_array_equal(np.argmax(res.C, axis=0), np.argmax(Hr, axis=0))


def test_toy_cases_of_reference_tests(kernels, golden):
    """tests/test_transfer.py of the reference: toy matrices, SC3 filters, pre-processing indices."""
    from functools import partial
    from scrna_b200 import NmfClustering
    from scrna_b200.sc3_clustering_impl import cell_filter, gene_filter
    g = golden("toys")
    ids = np.array(['lbl0', 'lbl1', 'lbl2'])
    nmf = NmfClustering(g["src1"], ids, 2)
    nmf.add_cell_filter(partial(cell_filter, num_expr_genes=1, non_zero_threshold=1))
    nmf.add_gene_filter(partial(gene_filter, perc_consensus_genes=0.96, non_zero_threshold=1))
    nmf.apply()
    np.testing.assert_array_equal(nmf.pp_data, g["src1"][:2, :4])
    np.testing.assert_array_equal(nmf.remain_gene_inds, np.arange(2))
    np.testing.assert_array_equal(nmf.remain_cell_inds, np.arange(4))
    assert rel_fro(nmf.dictionary @ nmf.data_matrix, g["t1_W"] @ g["t1_H"]) < 1e-3
    np.testing.assert_array_equal(nmf.cluster_labels, g["t1_labels"])


def test_early_exit_after_convergence_is_a_noop(kernels, default_sim, golden):
    """Iterations enqueued past convergence must not change the factors (device-side done flag)."""
    g = golden("nmf_source")
    X = default_sim["src"][:200, :300]
    W0, H0 = no.nndsvdar_init(X, 4)
    slv = _solver(kernels, X, 4)
    r1 = slv.fit(W0, H0, solver="cd", tol=1e-2, max_iter=100, poll=1)
    r2 = slv.fit(W0, H0, solver="cd", tol=1e-2, max_iter=100, poll=16)
    assert r1.n_iter == r2.n_iter
    np.testing.assert_array_equal(r1.G, r2.G)
    np.testing.assert_array_equal(r1.C, r2.C)


@pytest.mark.parametrize("shape_k", [((1000, 1000), 8), ((600, 1500), 6), ((1500, 400), 12)])
def test_device_nndsvdar_init_matches_sklearn(kernels, default_sim, shape_k, sweep_path):
    """Row a4/f1: randomized SVD + NNDSVDar with the 16 passes over X on the device; same RNG stream."""
    from sklearn.decomposition._nmf import _initialize_nmf
    from scrna_b200.nmf_init import nndsvdar_init_device
    from scrna_b200.nmf_solver import DeviceMatrix
    (g, n), k = shape_k
    if (g, n) == (1000, 1000):
        X = default_sim["src"]
    else:
        rng = np.random.RandomState(g + n)
        mu = rng.gamma(2.0, 1.0, size=(g, 1)) * (1 + (rng.rand(g, 6) < 0.2) @ (rng.rand(6, n) < 0.3) * 3.0)
        X = rng.poisson(mu).astype(np.float64)
    W0, H0, XT = nndsvdar_init_device(DeviceMatrix.from_host(X, kernels), k, kernels, seed=0)
    Wr, Hr = _initialize_nmf(X, k, init="nndsvdar", random_state=0)
    assert XT is not None and W0.shape == Wr.shape and H0.shape == Hr.shape
    # the zero pattern decides where the random fill lands: it must agree for the fills to line up
    big_w, big_h = Wr > 1e-3 * Wr.max(), Hr > 1e-3 * Hr.max()
    assert rel_fro(W0[big_w], Wr[big_w]) < 1e-4 and rel_fro(H0[big_h], Hr[big_h]) < 1e-4
    assert rel_fro(W0, Wr) < 1e-3 and rel_fro(H0, Hr) < 1e-3


def test_nmf_clustering_with_device_init(kernels, default_sim, golden):
    """NmfClustering.apply(init='device') lands on the reference's factorization (same labels, n_iter +-1)."""
    from scrna_b200 import NmfClustering
    g = golden("nmf_source")
    nmf = NmfClustering(default_sim["src"], default_sim["gene_ids"], 8, default_sim["src_labels"])
    nmf.apply(k=8, init='device', **CLI)
    assert abs(nmf.n_iter - int(g["nmf_n_iter"])) <= 1
    assert rel_fro(nmf.dictionary, g["nmf_W"]) < TOL_FACTOR
    assert rel_fro(nmf.data_matrix, g["nmf_H"]) < TOL_FACTOR
    np.testing.assert_array_equal(nmf.cluster_labels, g["nmf_labels"])


def test_fixw_default_data_vs_reference_golden(kernels, default_sim, golden, sweep_path):
    """NmfClustering_fixW (scRNA/nmf_clustering.py:92-130): two fixed-factor CD solves, against the reference's
    own output (tests/golden/fixw.npz, written by tests/golden/make_golden_fixw.py)."""
    from scrna_b200 import NmfClustering_fixW
    g = golden("fixw")
    nmf = NmfClustering_fixW(default_sim["src"], default_sim["gene_ids"], 7, default_sim["src_labels"])
    nmf.apply(k=7, alpha=1.0, l1=0.75, max_iter=100, rel_err=1e-3)
    assert rel_fro(nmf.dictionary, g["dictionary"]) < TOL_FACTOR
    np.testing.assert_array_equal(np.asarray(nmf.data_matrix, dtype=np.float64), g["data_matrix"])
    np.testing.assert_array_equal(nmf.cluster_labels, g["cluster_labels"])


def test_nan_factors_raise_like_the_reference(kernels, default_sim):
    """nmf_clustering.py:33-38: NaNs in a factor raise Exception('W contains NaNs ...') / ('H contains NaNs ...').
    (With the CD solver a NaN entry is squashed to 0 by max(w - grad/hess, 0) -- in sklearn's sweep as in ours --
    so the multiplicative update carries the NaN to the check.)"""
    from scrna_b200 import NmfClustering
    X = default_sim["src"][:60, :50]
    W0 = np.abs(np.random.RandomState(0).randn(60, 3))
    H0 = np.abs(np.random.RandomState(1).randn(3, 50))
    W0[5, 1] = np.nan
    nmf = NmfClustering(X, np.arange(60), 3)
    with pytest.raises(Exception, match="contains NaNs"):
        nmf.apply(k=3, max_iter=5, init=(W0, H0), solver='mu')


def test_reject_ratio_fails_like_the_reference(kernels, default_sim, golden):
    """nmf_clustering.py:196-197 slices with a float: the reference raises TypeError for reject_ratio > 0, so do we;
    reject_ratio >= 1 trips the assertion (:190)."""
    from scrna_b200 import NmfClustering_initW, DaNmfClustering
    src = NmfClustering_initW(default_sim["src"][:200, :150], np.arange(200), 7, default_sim["src_labels"][:150])
    k = len(np.unique(default_sim["src_labels"][:150]))
    src.apply(k=k, alpha=10., l1=0.75, max_iter=50, rel_err=1e-3)
    da = DaNmfClustering(src, default_sim["trg"][:200].copy(), np.arange(200), k)
    np.random.seed(0)
    with pytest.raises(TypeError):
        da.get_mixed_data(mix=0.5, reject_ratio=0.2)
    with pytest.raises(AssertionError):
        da.get_mixed_data(mix=0.5, reject_ratio=1.0)
