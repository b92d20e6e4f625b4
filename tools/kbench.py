"""Kernel micro-benchmark (development tool): times the two X sweeps and whole iterations with CUDA
events on the launching stream.  X (>= several GB) is far larger than L2, so no flush is needed."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from scrna_b200.kernels import default_kernels
from scrna_b200.nmf_solver import DeviceMatrix, NmfSolver
from scrna_b200 import _lib


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--k", type=int, nargs="+", default=[8])
    ap.add_argument("--sa", type=int, nargs="*", default=[])
    ap.add_argument("--sb", type=int, nargs="*", default=[])
    ap.add_argument("--lowmem", action="store_true")
    ap.add_argument("--variants", type=int, nargs="*", default=[2])
    ap.add_argument("--tc", default="auto", help="auto | 0 | 1 | f16 | tf32: sweep choice passed to NmfSolver")
    ap.add_argument("--real", action="store_true", help="non-integer X (general 3xTF32 path, Xl pass on)")
    ap.add_argument("--out", default="gpurun_out/kbench.json")
    args = ap.parse_args()
    kn = default_kernels()
    g, n = args.genes, args.cells
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    ldx = ((n + 31) // 32) * 32
    X = torch.zeros((g, ldx), dtype=torch.float32, device=dev)
    rows = 2000
    for r0 in range(0, g, rows):
        r1 = min(g, r0 + rows)
        lam = torch.rand((r1 - r0, 1), device=dev, generator=gen) * 4
        X[r0:r1, :n] = torch.poisson(lam.expand(r1 - r0, n), generator=gen)
        if args.real:
            X[r0:r1, :n] = torch.log2(X[r0:r1, :n] + 1.0)
    dm = DeviceMatrix(X, n)
    bytes_x = g * ldx * 4
    out = []
    for k in args.k:
        tc = {"auto": "auto", "0": False, "1": True, "f16": "f16", "tf32": "tf32"}[args.tc]
        slv = NmfSolver(dm, k, kernels=kn, tensor_cores=tc)
        rng = np.random.RandomState(0)
        G0 = np.abs(rng.randn(g, k)); C0 = np.abs(rng.randn(k, n))
        slv._load_factors(G0, C0)
        kp = slv.kp
        sa_list = args.sa or [slv.SA]
        sb_list = args.sb or [slv.SB]
        if slv.tc or slv.half:
            eb = 2 if slv.half else 4
            name = "tcgen05 f16" if slv.half else "tcgen05 tf32"
            med, best = timeit(lambda: slv._sweep_X(None))
            xbytes = (slv.Xs.rows_pad * slv.Xs.npanels * 256 * 2) if slv.half else g * ldx * 4   # bytes one sweep streams
            out.append(dict(kernel="sweep_X(%s)" % name, k=k, kp=kp, exact=slv.half or slv.tc_flags == 32, nsplit=slv.SB,
                            ms=med, best_ms=best, gbs=xbytes / med / 1e6))
            print(out[-1], flush=True)
            XT = slv.XT
            med, best = timeit(lambda: slv._sweep_XT(None))
            xtbytes = (XT.rows_pad * XT.npanels * 256 * 2) if slv.half else XT.g * XT.ldx * 4
            out.append(dict(kernel="sweep_XT(%s)" % name, k=k, kp=kp, exact=slv.half or slv.tc_flags == 32, nsplit=slv.SA,
                            ms=med, best_ms=best, gbs=xtbytes / med / 1e6))
            print(out[-1], flush=True)
            sb_list, sa_list = [], []
            if slv.half:
                bytes_x = g * ldx * 2
        for sb in sb_list:
          for var in args.variants:
            PB = torch.zeros((sb, kp, slv.ldn), dtype=torch.float32, device=dev)
            med, best = timeit(lambda: kn.wtx_partial(X, ldx, g, dm.ncols, slv.Gs32, kp, PB, slv.ldn, sb, None, variant=var))
            out.append(dict(kernel="sweep_X(K3)", k=k, variant=var, nsplit=sb, ms=med, best_ms=best, gbs=bytes_x / med / 1e6))
            print(out[-1], flush=True)
            if var > 0:
                ref = torch.zeros_like(PB)
                kn.wtx_partial(X, ldx, g, dm.ncols, slv.Gs32, kp, ref, slv.ldn, sb, None, variant=0)
                print("   variant", var, "max abs diff vs variant 0:", float((ref - PB).abs().max()), flush=True)
        XT = slv.XT
        for sa in sa_list:
            PA = torch.zeros((sa, kp, slv.ldg), dtype=torch.float32, device=dev)
            med, best = timeit(lambda: kn.wtx_partial(XT.data, XT.ldx, XT.g, slv.gcols, slv.Cs32, kp, PA, slv.ldg, sa, None))
            out.append(dict(kernel="sweep_XT(K1)", k=k, nsplit=sa, ms=med, best_ms=best, gbs=XT.g * XT.ldx * 4 / med / 1e6))
            print(out[-1], flush=True)
        if args.lowmem:
            sa0 = kn.xht_splits(g, dm.ncols, kp)
            PA0 = torch.zeros((sa0, g, kp), dtype=torch.float32, device=dev)
            med, best = timeit(lambda: kn.xht_partial(X, ldx, g, dm.ncols, slv.Ct32, slv.ldn, kp, PA0, sa0, None))
            out.append(dict(kernel="xht_partial(lowmem K1)", k=k, nsplit=sa0, ms=med, best_ms=best, gbs=bytes_x / med / 1e6))
            print(out[-1], flush=True)
        Xr = slv.Xs.data
        med, best = timeit(lambda: kn.residual(Xr, slv.Xs.ldx, g, dm.ncols, slv.Gs32, slv.Ct32, slv.ldn, kp, 1, slv.res_part, slv.res_out))
        out.append(dict(kernel="residual_l1", k=k, ms=med, best_ms=best, gbs=bytes_x / med / 1e6))
        print(out[-1], flush=True)
        # whole iterations
        for solver in ("mu", "cd"):
            code = _lib.SOLVER_MU if solver == "mu" else _lib.SOLVER_CD
            slv._load_factors(G0, C0)
            perms = np.zeros((64, kp), dtype=np.int32); perms[:, :k] = np.arange(k)
            slv.perms = torch.from_numpy(perms).to(dev)

            slv._g_num_from_R = False
            slv._gram_C(None); slv._gram_G(None)

            def it():
                slv._products_for_G(None)
                slv._update_G(code, 7.5, 2.5, 0, 0, None)
                slv._half_C(code, 7.5, 2.5, 0, 0, None)
            med, best = timeit(it, warm=3, reps=10)
            algo = 2 * bytes_x + 2 * (g + n) * k * 4
            out.append(dict(kernel="iteration_" + solver, k=k, kp=kp, tc=bool(slv.tc), ms=med, best_ms=best, it_per_s=1000.0 / med,
                            gbs=algo / med / 1e6))
            print(out[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
