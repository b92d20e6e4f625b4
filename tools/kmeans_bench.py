"""One k-means fit of the SC3 ensemble at full size (n cells x d leading singular vectors): wall seconds of
Sc3Ops.kmeans, Lloyd iterations, and the device time of one gated Lloyd iteration / one k-means++ init measured with
CUDA events on the launching stream."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scrna_b200 import _lib
from scrna_b200.sc3_device import Sc3Ops, kmeans_draw_count

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1100
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ops = Sc3Ops()
t = torch
rng = np.random.RandomState(0)
lab = rng.randint(0, k, size=n)
V = rng.randn(n, d) / np.sqrt(n)
V[:, :k] += 0.02 * np.eye(k)[lab]
Vd = t.from_numpy(V).cuda()
per, trials = kmeans_draw_count(k)
draws = rng.rand(10 * per)
out = {"cells": n, "dims": d, "k": k}
for rep in range(2):
    a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    a.record()
    labels = ops.kmeans(Vd, d, k, draws)
    b.record(); t.cuda.synchronize()
    out["fit_s_run%d" % rep] = a.elapsed_time(b) / 1e3
    out["lloyd_iterations"] = ops.km_iters
from sklearn.metrics import adjusted_rand_score
out["ari"] = float(adjusted_rand_score(lab, labels.cpu().numpy() if hasattr(labels, "cpu") else labels))

# per-kernel device times
X = Vd.contiguous()
C = X[:k].clone().contiguous(); C2 = C.clone()
labels = t.zeros(n, dtype=t.int32, device="cuda"); lold = labels.clone()
counts = t.zeros(k, dtype=t.int32, device="cuda")
celld = t.empty(n, dtype=t.float64, device="cuda")
work = t.empty(_lib.call("scrna_lloyd_sums_work_elems", k, d), dtype=t.float64, device="cuda")
state = t.zeros(8, dtype=t.float64, device="cuda")
s = t.cuda.current_stream().cuda_stream


def timed(name, fn, reps=20):
    for _ in range(3):
        fn()
    a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); t.cuda.synchronize()
    out[name + "_us"] = round(a.elapsed_time(b) * 1e3 / reps, 2)


p = lambda x: x.data_ptr()
timed("assign", lambda: _lib.call("scrna_lloyd_assign", p(X), d, n, d, p(C), d, k, p(labels), s))
timed("sums", lambda: _lib.call("scrna_lloyd_sums", p(X), d, n, d, p(labels), k, p(C2), d, p(counts), p(work), s))
timed("finish", lambda: _lib.call("scrna_lloyd_finish", p(C2), p(C), d, k, d, p(counts), p(labels), p(lold), n, p(state), s))
timed("iteration", lambda: (state.zero_(), _lib.call("scrna_lloyd_iteration", p(X), d, n, d, p(C), p(C2), d, k, p(labels),
                                                    p(lold), p(counts), p(celld), p(work), p(state), 1000, s)))
state[6] = 1.0
timed("gated_iteration", lambda: _lib.call("scrna_lloyd_iteration", p(X), d, n, d, p(C), p(C2), d, k, p(labels), p(lold),
                                           p(counts), p(celld), p(work), p(state), 1000, s))
# one k-means++ initialisation (4k+1 launches)
xsq = (X * X).sum(1).contiguous()
U = t.from_numpy(draws).cuda()
closest = t.empty(n, dtype=t.float64, device="cuda"); cand_dist = t.empty(trials * n, dtype=t.float64, device="cuda")
cums = t.empty(n, dtype=t.float64, device="cuda"); pots = t.empty(8, dtype=t.float64, device="cuda")
cand = t.zeros(8, dtype=t.int32, device="cuda"); idx = t.zeros(k, dtype=t.int32, device="cuda")
state.zero_()
timed("kmeans_pp_init", lambda: _lib.call("scrna_kmeans_pp_from", p(X), d, n, d, p(xsq), k, 17, p(U) + 8, trials,
                                          p(closest), p(cand_dist), p(cums), p(pots), p(cand), p(idx), p(state), p(C), d, s))
timed("inertia", lambda: _lib.call("scrna_km_inertia", p(X), d, n, d, p(C), d, p(labels), p(celld), p(state), s))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/kmeans_bench_%d_%d.json" % (n, d), "w"))
