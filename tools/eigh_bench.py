"""Timing + property check of the tridiagonalisation solver at SC3 sizes (dev tool)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scrna_b200.sc3_device import Sc3Ops
ops = Sc3Ops()
out = []
for n in [int(a) for a in sys.argv[1:]] or [4096]:
    g, k = min(n, 4000), 8
    gen = torch.Generator(device="cuda"); gen.manual_seed(n)
    centers = torch._standard_gamma(torch.full((g, k), 2.0, device="cuda"), generator=gen) * 2.0
    lab = torch.randint(0, k, (n,), device="cuda", generator=gen)
    X = torch.poisson(centers[:, lab] * torch.pow(2.0, 0.3 * torch.randn(n, device="cuda", generator=gen))[None, :], generator=gen).double()
    D = ops.distances(X, "euclidean")
    M = ops.transform_matrix(D, "pca")
    c = int(np.ceil(0.07 * n))
    ops.stage_seconds = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    V, vals = ops.right_singular_vectors(M, c, method="eigh")
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = dict(ops.stage_seconds); ops.stage_seconds = None
    eye = torch.eye(c, dtype=torch.float64, device="cuda")
    orth = float((V.T @ V - eye).abs().max().cpu())
    s2 = torch.from_numpy(vals[:c] ** 2).to("cuda")
    res = float((M.T @ (M @ V) - V * s2).norm().cpu()) / vals[0] ** 2
    rec = dict(n=n, c=c, seconds=dt, stages=st, orthogonality=orth, residual=res, sigma_first=float(vals[0]), sigma_last=float(vals[c - 1]))
    if n <= 8192:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Vj, vj = ops.right_singular_vectors(M, c, method="jacobi")
        torch.cuda.synchronize(); rec["jacobi_seconds"] = time.perf_counter() - t0
        rec["sigma_max_diff_vs_jacobi"] = float(np.abs(vj[:c] - vals[:c]).max())
        sgn = torch.sign((V * Vj).sum(0))
        rec["vector_max_diff_vs_jacobi"] = float((V * sgn - Vj).abs().max().cpu())
    print(json.dumps(rec), flush=True); out.append(rec)
json.dump(out, open("gpurun_out/eigh_bench.json", "w"), indent=1)
