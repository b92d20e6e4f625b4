"""BASELINE.json configs[4]: atlas-scale synthetic 30k genes x 1M source cells NMF at rank 50, cells sharded
over the GPUs of one box, then the dictionary transferred to a 10k-cell target (H-only MU, utils.py:187-230).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/atlas_bench.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=30000)
    ap.add_argument("--cells", type=int, default=1000000)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--target-cells", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/atlas_bench.json")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from scrna_b200.kernels import default_kernels
    from scrna_b200.nmf_solver import DeviceMatrix, NmfSolver, TorchDistComm, shard_bounds
    from scrna_b200.transfer import TransferSession
    if world > 1:
        comm = TorchDistComm()
    kn = default_kernels()
    g, n, k = args.genes, args.cells, args.k
    c0, c1 = shard_bounds(n, world, rank)
    Xd, n_loc = bench.synth_counts(torch, dev, g, c0, c1, clusters=k)
    dm = DeviceMatrix(Xd, n_loc)
    xsum = Xd.sum(dtype=torch.float64)
    if world > 1:
        dist.all_reduce(xsum)
    xmean = float(xsum.item()) / (float(g) * n)
    # NNDSVDar on the device over the cell shards (scrna_b200/nmf_init.py: what NmfClustering.apply(comm=,
    # cells='sharded') runs): the randomized-SVD passes are sweeps over the shards, the host factorisations replicated
    from scrna_b200.nmf_init import nndsvdar_init_device
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    G0, H0, XT = nndsvdar_init_device(dm, k, kn, seed=0, comm=comm, cell_range=(c0, c1, n) if world > 1 else None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    init_s = time.perf_counter() - t0
    C0 = np.ascontiguousarray(H0[:, c0:c1]) if H0.shape[1] == n and world > 1 else H0
    del H0
    slv = NmfSolver(dm, k, kernels=kn, comm=comm, XT=XT)
    slv.prepare(G0, C0, solver="mu", l1_G=bench.L1, l2_G=bench.L2, l1_C=bench.L1, l2_C=bench.L2, tol=0.0,
                max_iter=10 ** 9)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        slv.step()
    barrier()
    kn.sweep_events = []
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        slv.step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    sw = kn.sweep_events
    kn.sweep_events = None
    sweep_ms = sum(a.elapsed_time(b) for a, b, _ in sw) / len(sw)
    sweep_gbs = (sum(nb for _, _, nb in sw) / len(sw)) / (sweep_ms * 1e-3) / 1e9
    err = slv.frobenius_error()
    G = slv.G64[:k, :g].cpu().numpy().T.copy()
    out = {"workload": "NMF MU %d genes x %d cells k=%d on %d GPU(s) (BASELINE configs[4])" % (g, n, k, world),
           "iterations_per_s": args.steps / (ms / 1e3), "ms_per_iteration": ms / args.steps,
           "init": "NNDSVDar on the device, cells sharded (nndsvdar_init_device with comm)", "init_seconds": init_s,
           "sweep_kernel": ("tcgen05 kind::f16, fp16 storage, kq=%d" % slv.kp if slv.half else
                            "tcgen05 3xTF32 kp=%d%s" % (slv.kp, " (TF32-exact X)" if slv.tc_flags else "") if slv.tc
                            else "fp32"),
           "sweep_ms_per_gpu": sweep_ms, "sweep_gbs_per_gpu": sweep_gbs,
           "x_bytes_per_gpu": int(g) * dm.ldx * (2 if slv.half else 4), "frobenius_error_after": err}
    del slv, dm, Xd
    torch.cuda.empty_cache()
    if rank == 0:
        Xt, nt = bench.synth_counts(torch, dev, g, 10 ** 7, 10 ** 7 + args.target_cells, clusters=k)
        tdm = DeviceMatrix(Xt, nt)
        Hi = np.abs(np.random.RandomState(5).randn(k, nt))
        Hi[Hi < 1e-10] = 1e-10
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sess = TransferSession(G, tdm, kernels=kn)
            n_it, terr = sess.solve(Hi)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out["transfer"] = {"target_cells": nt, "iterations": int(n_it), "seconds": dt, "l1_error": float(terr),
                           "what": "TransferSession(W, X_t).solve(H0): W^T X once, H-only MU to rel. change 1e-3, "
                                   "sum|X - W.H| per iteration in one sweep"}
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
