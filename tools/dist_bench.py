"""SC3 distance stage at BASELINE configs[3] size: Euclidean / Pearson / Spearman on 20 000 cells x 20 000 genes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scrna_b200.sc3_device import Sc3Ops
ops = Sc3Ops()
n = g = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
gen = torch.Generator(device=ops.dev); gen.manual_seed(3)
lam = torch.rand((g, 1), device=ops.dev, generator=gen) * 6
X = torch.log2(torch.poisson(lam.expand(g, n), generator=gen).double() + 1.0)
ops.distances(X[:, :64].contiguous(), "euclidean")
out = {"cells": n, "genes": g}
for m in ("euclidean", "pearson", "spearman"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    D = ops.distances(X, m)
    torch.cuda.synchronize(); out[m + "_s"] = time.perf_counter() - t0
    out[m + "_diag_max"] = float(D.diagonal().abs().max().cpu())
    out[m + "_sym_err"] = float((D[:2000, :2000] - D[:2000, :2000].T).abs().max().cpu())
    del D
print(json.dumps(out))
json.dump(out, open("gpurun_out/dist_bench.json", "w"), indent=1)
