"""SC3Clustering.apply through the reference's plug-in API (scRNA/cmd_target.py:220-242 wiring) on a synthetic
target: seconds end to end including the NumPy round trips the API contract implies."""
import json, os, sys, time
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scrna_b200 import SC3Clustering
from scrna_b200 import sc3_clustering_impl as sc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
g, k = n, 8
rng = np.random.RandomState(n)
centers = rng.gamma(2.0, 2.0, size=(g, k))
lab = rng.randint(0, k, size=n)
X = rng.poisson(centers[:, lab] * 2.0 ** (0.3 * rng.randn(1, n))).astype(np.float64)
out = {"cells": n, "genes": g}
for rep in range(2):
    np.random.seed(0)
    max_pca, min_pca = int(np.ceil(n * 0.07)), int(max(1, np.floor(n * 0.04)))
    sc3 = SC3Clustering(X, np.arange(g), pc_range=[min_pca, max_pca], sub_sample=True, consensus_mode=0)
    sc3.add_distance_calculation(partial(sc.distances, metric='euclidean'))
    sc3.add_dimred_calculation(partial(sc.transformations, components=max_pca, method='pca'))
    sc3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
    sc3.set_build_consensus_matrix(sc.build_consensus_matrix)
    sc3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
    ops = sc._ops()
    ops.stage_seconds = {} if os.environ.get("STAGES", "1") == "1" else None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc3.apply()
    torch.cuda.synchronize(); out["apply_s_run%d" % rep] = time.perf_counter() - t0
    if ops.stage_seconds is not None:
        out["stages_run%d" % rep] = {a: round(b, 4) for a, b in ops.stage_seconds.items()}
    ops.stage_seconds = None
from sklearn.metrics import adjusted_rand_score
out["ari_vs_simulated"] = float(adjusted_rand_score(lab, sc3.cluster_labels))
print(json.dumps(out))
json.dump(out, open("gpurun_out/sc3_api_bench_%d.json" % n, "w"))
