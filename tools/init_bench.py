"""Times the device NNDSVDar init (scrna_b200/nmf_init.py) on the synthetic 20k x 100k workload and, on a
bounded sample, sklearn's host routine (the reference's path) for comparison."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from scrna_b200.kernels import default_kernels
from scrna_b200.nmf_init import nndsvdar_init_device
from scrna_b200.nmf_solver import DeviceMatrix


def main():
    genes, cells, k = 20000, 100000, 8
    kn = default_kernels()
    dev = torch.device("cuda", 0)
    Xd, n = bench.synth_counts(torch, dev, genes, 0, cells)
    dm = DeviceMatrix(Xd, n)
    out = {}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        W0, H0, XT = nndsvdar_init_device(dm, k, kn, seed=0)
        torch.cuda.synchronize()
        out["device_init_s_run%d" % rep] = time.perf_counter() - t0
        del XT
    out["launches"] = kn.launches
    # host routine on 1/10 of the cells
    sample = 10000
    Xs = Xd[:, :sample].to(torch.float64).cpu().numpy()
    from sklearn.decomposition._nmf import _initialize_nmf
    t0 = time.perf_counter()
    Wr, Hr = _initialize_nmf(Xs, k, init="nndsvdar", random_state=0)
    out["host_init_s_%dcells" % sample] = time.perf_counter() - t0
    out["host_init_s_scaled_to_%d" % cells] = out["host_init_s_%dcells" % sample] * cells / sample
    out["cores"] = os.cpu_count()
    # agreement on the sample
    W1, H1, _ = nndsvdar_init_device(DeviceMatrix(Xd[:, :sample + 16].contiguous()[:, :sample + 16], sample), k, kn, seed=0)
    big = Wr > 1e-3 * Wr.max()
    out["rel_err_W_sample"] = float(np.linalg.norm(W1[big] - Wr[big]) / np.linalg.norm(Wr[big]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
