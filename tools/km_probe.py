import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scrna_b200.sc3_device import Sc3Ops, kmeans_draw_count
ops = Sc3Ops()
n, k = 20000, 8
g = torch.Generator(device=ops.dev); g.manual_seed(1)
V = torch.randn((n, 1401), dtype=torch.float64, device=ops.dev, generator=g) / np.sqrt(n)
V[:, :8] += (torch.randint(0, 8, (n, 1), device=ops.dev, generator=g) == torch.arange(8, device=ops.dev)[None, :]).double() * 0.02
per, _ = kmeans_draw_count(k)
for d in (800, 1400):
    u = np.random.RandomState(d).random_sample(10 * per)
    ops.kmeans(V, d, k, u, n_init=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ops.kmeans(V, d, k, u)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps(dict(d=d, seconds=dt, lloyd_iters=ops.km_iters, ms_per_iter=1e3 * dt / max(ops.km_iters, 1))), flush=True)
# per-kernel timing of one Lloyd iteration at d=1400
import scrna_b200._lib as L
