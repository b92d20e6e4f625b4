"""SC3 stage timings on the GPU next to the reference's scipy/sklearn calls on the host cores
(n in {1000, 2000, 4000}; the 20k-cell CPU run is hours and is not attempted).
Writes gpurun_out/sc3_bench.json."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root (this file lives in tests/: it uses oracle/ as the CPU checker)

import numpy as np
import torch

from scrna_b200.sc3_device import Sc3Ops, kmeans_draw_count


def gpu_time(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def cpu_time(fn):
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs="+", default=[1000, 2000])
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--single", action="store_true",
                    help="one distance (euclidean) and one transformation (pca), each computed once: the "
                         "reference CLI's default SC3 configuration; for the 20k-cell case")
    ap.add_argument("--out", default="gpurun_out/sc3_bench.json")
    args = ap.parse_args()
    ops = Sc3Ops()
    res = []
    for n in args.cells:
        rng = np.random.RandomState(n)
        g, k = args.genes, args.k
        centers = rng.gamma(2.0, 2.0, size=(g, k))
        lab = rng.randint(0, k, size=n)
        X = rng.poisson(centers[:, lab] * 2.0 ** (0.3 * rng.randn(1, n))).astype(np.float64)
        pc = [max(1, int(np.floor(0.04 * n))), int(np.ceil(0.07 * n))]
        row = {"cells": n, "genes": g, "k": k, "pc_range": pc, "gpu": {}, "cpu": {}}
        ops.distances(X[:, :64], "euclidean")  # warm-up
        Xd = ops.to_device(X)
        D = {}
        for m in (("euclidean",) if args.single else ("euclidean", "pearson", "spearman")):
            row["gpu"]["dist_" + m], D[m] = gpu_time(lambda: ops.distances(Xd, m))
        Vp = None
        for t in (("pca",) if args.single else ("pca", "spectral")):
            def f():
                M = ops.transform_matrix(D["euclidean"], t)
                return ops.right_singular_vectors(M, pc[1])
            row["gpu"]["transf_" + t], (V, vals) = gpu_time(f)
            row["gpu"]["jacobi_sweeps_" + t] = ops.jacobi_sweeps
            if t == "pca":
                Vp = V
        per, _ = kmeans_draw_count(k)
        dims = list(range(pc[0], pc[1] + 1))
        if len(dims) > 15:
            dims = list(np.random.RandomState(0).permutation(dims)[:15])
        draws = [np.random.RandomState(d).random_sample(10 * per) for d in dims]
        del Xd

        def km():
            return [ops.kmeans(Vp, int(d), k, u) for d, u in zip(dims, draws)]
        row["gpu"]["kmeans_15dims"], labs = gpu_time(km)

        def cons():
            M = ops.coassoc_new(n)
            for l in labs:
                ops.coassoc_add(M, l)
            C = ops.consensus_from_counts(M, float(len(labs)))
            return ops.consensus_clustering(C, k)
        row["gpu"]["consensus_hclust"], (glabels, _, _) = gpu_time(cons)
        row["gpu"]["total_1dist_1transf"] = (row["gpu"]["dist_euclidean"] + row["gpu"]["transf_pca"] +
                                             row["gpu"]["kmeans_15dims"] + row["gpu"]["consensus_hclust"])
        from sklearn.metrics import adjusted_rand_score
        row["ari_vs_simulated_labels"] = float(adjusted_rand_score(lab, glabels))
        if not args.no_cpu:
            # CPU comparison leg (measurement tooling, like bench.py's cpu_baseline): the reference's scipy / sklearn
            # calls as restated in oracle/ -- imported here only, never by the product
            from oracle import sc3_oracle as so
            Dc = {}
            for m in ("euclidean", "pearson", "spearman"):
                row["cpu"]["dist_" + m], Dc[m] = cpu_time(lambda: so.distances(X, m))
            for t in ("pca", "spectral"):
                row["cpu"]["transf_" + t], tr = cpu_time(lambda: so.transformations(Dc["euclidean"], pc[1], t))
            Vc = so.transformations(Dc["euclidean"], pc[1], "pca")[1]
            row["cpu"]["kmeans_15dims"], clabs = cpu_time(lambda: [so.kmeans(Vc[:, :int(d)], k, u) for d, u in zip(dims, draws)])

            def ccons():
                C = np.zeros((n, n))
                for l in clabs:
                    C += so.build_consensus_matrix(l)
                C /= len(clabs)
                return so.consensus_clustering(C, k)
            row["cpu"]["consensus_hclust"], (clabels, _, _) = cpu_time(ccons)
            row["cpu"]["total_1dist_1transf"] = (row["cpu"]["dist_euclidean"] + row["cpu"]["transf_pca"] +
                                                 row["cpu"]["kmeans_15dims"] + row["cpu"]["consensus_hclust"])
            row["cpu"]["cores"] = os.cpu_count()
            row["labels_equal_cpu_gpu"] = bool(np.array_equal(clabels, glabels))
            row["kmeans_labelings_equal"] = int(sum(np.array_equal(a, b) for a, b in zip(labs, clabs)))
            row["euclid_bit_exact"] = bool(np.array_equal(D["euclidean"].cpu().numpy(), Dc["euclidean"]))
        print(json.dumps(row), flush=True)
        res.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
