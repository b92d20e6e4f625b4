"""SC3Clustering.apply over the GPUs of one node (torchrun --nproc-per-node N tools/sc3_multi_gpu.py CELLS [check]):
the distance matrix, the Gram of the eigensolver, the back-transformed vectors and the consensus row distances are
computed in row blocks by different ranks and all-gathered; the k-means fits are dealt round-robin over the ranks.
With `check`, rank 0 also runs the single-GPU path on the same data and the labels must be identical."""
import json, os, sys, time
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from scrna_b200 import SC3Clustering
from scrna_b200 import sc3_clustering_impl as sc
from scrna_b200.nmf_solver import TorchDistComm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
check = len(sys.argv) > 2 and sys.argv[2] == "check"
g, k = n, 8
rng = np.random.RandomState(n)
centers = rng.gamma(2.0, 2.0, size=(g, k))
lab = rng.randint(0, k, size=n)
X = rng.poisson(centers[:, lab] * 2.0 ** (0.3 * rng.randn(1, n))).astype(np.float64)


def run(comm):
    np.random.seed(0)
    max_pca, min_pca = int(np.ceil(n * 0.07)), int(max(1, np.floor(n * 0.04)))
    s3 = SC3Clustering(X, np.arange(g), pc_range=[min_pca, max_pca], sub_sample=True, consensus_mode=0, comm=comm)
    s3.add_distance_calculation(partial(sc.distances, metric='euclidean'))
    s3.add_dimred_calculation(partial(sc.transformations, components=max_pca, method='pca'))
    s3.add_intermediate_clustering(partial(sc.intermediate_kmeans_clustering, k=k))
    s3.set_build_consensus_matrix(sc.build_consensus_matrix)
    s3.set_consensus_clustering(partial(sc.consensus_clustering, n_components=k))
    ops = sc._ops()
    ops.stage_seconds = {}
    torch.cuda.synchronize()
    if comm is not None:
        dist.barrier()
    t0 = time.perf_counter()
    s3.apply()
    torch.cuda.synchronize()
    if comm is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    st = {a: round(b, 4) for a, b in ops.stage_seconds.items()}
    ops.stage_seconds = None
    return s3.cluster_labels, dt, st


comm = TorchDistComm() if world > 1 else None
out = {"cells": n, "genes": g, "n_gpus": world}
for rep in range(2):
    labels, dt, st = run(comm)
    out["apply_s_run%d" % rep] = dt
    out["stages_rank0_run%d" % rep] = st
if world > 1:
    tl = torch.from_numpy(np.asarray(labels, dtype=np.int64)).cuda()
    ref = tl.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(tl, ref))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["labels_identical_on_all_ranks"] = bool(same.item())
if check and rank == 0:
    single, dt1, st1 = run(None)
    out["single_gpu_apply_s"] = dt1
    out["labels_equal_to_single_gpu"] = bool(np.array_equal(single, labels))
if rank == 0:
    from sklearn.metrics import adjusted_rand_score
    out["ari_vs_simulated"] = float(adjusted_rand_score(lab, labels))
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sc3_multi_gpu_%d_n%d.json" % (n, world), "w"))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
