"""Times the block one-sided Jacobi (K13, sc3_svd_block.cu) on an SC3-like matrix: a few large directions over a
flat noise bulk, built on the device.  Reports sweeps, seconds and the quality of the kept vectors."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scrna_b200.sc3_device import Sc3Ops


def main():
    ops = Sc3Ops()
    out = []
    for n in [int(a) for a in sys.argv[1:]]:
        g = torch.Generator(device=ops.dev)
        g.manual_seed(n)
        lab = torch.randint(0, 8, (n,), device=ops.dev, generator=g)
        cen = torch.randn((8, 64), dtype=torch.float64, device=ops.dev, generator=g) * 3.0
        pts = cen[lab] + torch.randn((n, 64), dtype=torch.float64, device=ops.dev, generator=g)
        D = torch.cdist(pts, pts)                      # test-input construction only (not a product path)
        M = ops.transform_matrix(D, "pca")
        del D
        c = int(np.ceil(0.07 * n))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        V, vals = ops.right_singular_vectors(M, c)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        MV = M @ V                                      # checks only
        res = (M.T @ MV - V * torch.from_numpy(vals[:c] ** 2).to(ops.dev)).norm().item() / vals[0] ** 2
        orth = (V.T @ V - torch.eye(c, dtype=torch.float64, device=ops.dev)).abs().max().item()
        rec = dict(n=n, components=c, sweeps=ops.jacobi_sweeps, seconds=dt, s_per_sweep=dt / ops.jacobi_sweeps,
                   tfma_per_s=4.0 * n ** 3 / (dt / ops.jacobi_sweeps) / 1e12, residual=res, orthogonality=orth,
                   sigma_1=float(vals[0]), sigma_c=float(vals[c - 1]))
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del M, V, MV
        torch.cuda.empty_cache()
    json.dump(out, open("gpurun_out/svd_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
