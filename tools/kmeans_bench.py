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
got = {}
for mode in (True, False):
    for rep in range(2):
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        labels = ops.kmeans(Vd, d, k, draws, batched=mode)
        b.record(); t.cuda.synchronize()
        out["fit_s_%s_run%d" % ("batched" if mode else "serial", rep)] = a.elapsed_time(b) / 1e3
        out["lloyd_iterations_%s" % ("batched" if mode else "serial")] = ops.km_iters
    got[mode] = np.asarray(labels)
out["labels_equal_batched_vs_serial"] = bool((got[True] == got[False]).all())
from sklearn.metrics import adjusted_rand_score
out["ari"] = float(adjusted_rand_score(lab, labels.cpu().numpy() if hasattr(labels, "cpu") else labels))

# per-kernel device times
X = t.zeros(n, ((d + 3) // 4) * 4, dtype=t.float64, device="cuda"); X[:, :d] = Vd
ld = X.stride(0)
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
timed("assign", lambda: _lib.call("scrna_lloyd_assign", p(X), ld, n, d, p(C), ld, k, p(labels), s))
timed("sums", lambda: _lib.call("scrna_lloyd_sums", p(X), ld, n, d, p(labels), k, p(C2), ld, p(counts), p(work), s))
timed("finish", lambda: _lib.call("scrna_lloyd_finish", p(C2), p(C), ld, k, d, p(counts), p(labels), p(lold), n, p(state), s))
timed("iteration", lambda: (state.zero_(), _lib.call("scrna_lloyd_iteration", p(X), ld, n, d, p(C), p(C2), ld, k, p(labels),
                                                    p(lold), p(counts), p(celld), p(work), p(state), 1000, s)))
state[6] = 1.0
timed("gated_iteration", lambda: _lib.call("scrna_lloyd_iteration", p(X), ld, n, d, p(C), p(C2), ld, k, p(labels), p(lold),
                                           p(counts), p(celld), p(work), p(state), 1000, s))
# k-means++ of all ten restarts (4k+1 launches) and one batched Lloyd iteration
R = 10
ld = X.stride(0)
xsq = (X * X).sum(1).contiguous()
U = t.from_numpy(draws).cuda()
closest = t.empty(R, n, dtype=t.float64, device="cuda"); cand_dist = t.empty(R, trials * n, dtype=t.float64, device="cuda")
cums = t.empty(R, n, dtype=t.float64, device="cuda"); pots = t.empty(R, 8, dtype=t.float64, device="cuda")
cand = t.zeros(R, 8, dtype=t.int32, device="cuda"); idx = t.zeros(R, k, dtype=t.int32, device="cuda")
stR = t.zeros(R, 8, dtype=t.float64, device="cuda")
firsts = t.arange(17, 17 + R, dtype=t.int32, device="cuda")
CA = t.zeros(R, k, ld, dtype=t.float64, device="cuda"); CB = t.zeros_like(CA)
if ld % 2 == 0:
    timed("kmeans_pp_all_restarts", lambda: _lib.call("scrna_kmeans_pp_batched", p(X), ld, n, d, p(xsq), k, R, p(firsts), p(U),
                                                      trials, p(closest), p(cand_dist), p(cums), p(pots), p(cand), p(idx),
                                                      p(stR), p(CA), ld, s))
    labR = t.zeros(R, n, dtype=t.int32, device="cuda"); labRo = t.zeros_like(labR)
    cntR = t.zeros(R, k, dtype=t.int32, device="cuda"); celR = t.empty(R, n, dtype=t.float64, device="cuda")
    bwork = t.empty(_lib.call("scrna_lloyd_batched_work_elems", k, d, R), dtype=t.float64, device="cuda")
    timed("batched_iteration_all_restarts", lambda: (stR.zero_(), _lib.call(
        "scrna_lloyd_iteration_batched", p(X), ld, n, d, p(CA), p(CB), ld, k, R, p(labR), p(labRo), p(cntR), p(celR),
        p(bwork), p(stR), 1000, s)))
timed("inertia", lambda: _lib.call("scrna_km_inertia", p(X), ld, n, d, p(C), ld, p(labels), p(celld), p(state), s))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/kmeans_bench_%d_%d.json" % (n, d), "w"))
