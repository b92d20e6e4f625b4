#!/usr/bin/env python
"""Bring-up probe of the tcgen05 sweep (scrna_nmf_wtx_partial_tc): one configuration per process so that a
faulting descriptor variant cannot poison the next one.

    python tools/tc_probe.py KP FLAGS ROWS COLS [time]

Compares sum over splits of the partials with an fp64 reference of F^T.X and prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    kp, flags, rows, cols = (int(a) for a in sys.argv[1:5])
    do_time = len(sys.argv) > 5
    import torch
    from scrna_b200.kernels import default_kernels
    kn = default_kernels()
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    ldx = ((cols + 31) // 32) * 32
    ncols = ((cols + 3) // 4) * 4
    X = torch.zeros((rows, ldx), dtype=torch.float32, device=dev)
    if do_time:
        X[:, :cols] = torch.poisson(torch.rand((1, cols), device=dev) * 3.0 + torch.zeros((rows, 1), device=dev))
    else:
        # non-integer, wide dynamic range: exercises the hi/lo split of both operands
        X[:, :cols] = torch.rand((rows, cols), device=dev) * torch.exp(3.0 * torch.randn((rows, cols), device=dev))
    k = kp - 3
    F = torch.zeros((kp, ((rows + 31) // 32) * 32), dtype=torch.float64, device=dev)
    F[:k, :rows] = torch.rand((k, rows), dtype=torch.float64, device=dev) + 0.01
    Fs = torch.empty(kn.tc_shadow_elems(rows, kp), dtype=torch.float32, device=dev)
    kn.tc_shadow(F, F.shape[1], rows, k, kp, Fs)
    if kn.tf32_exact(X) and not (flags & 64):
        flags |= 32
    ns = kn.wtx_tc_splits(rows, ncols, kp)
    part = torch.full((ns, kp, ldx), float("nan"), dtype=torch.float32, device=dev)
    kn.wtx_partial_tc(X, ldx, rows, ncols, Fs, kp, part, ldx, ns, None, flags=flags)
    torch.cuda.synchronize()
    cc = min(cols, 8192) if do_time else cols   # timing shapes: check a column subset
    got = part[:, :, :cc].to(torch.float64).sum(0)
    ref = F[:, :rows] @ X[:, :cc].to(torch.float64)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    rel_f = ((got - ref).norm() / ref.norm()).item()
    out = {"kp": kp, "flags": flags, "rows": rows, "cols": cols, "nsplit": ns, "max_err_rel": err, "fro_rel": rel_f,
           "pad_rows_zero": bool((got[k:] == 0).all().item())}
    if do_time:
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            kn.wtx_partial_tc(X, ldx, rows, ncols, Fs, kp, part, ldx, ns, None, flags=flags)
        a.record()
        n = 10
        for _ in range(n):
            kn.wtx_partial_tc(X, ldx, rows, ncols, Fs, kp, part, ldx, ns, None, flags=flags)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        out["ms"] = ms
        out["gbs"] = rows * ncols * 4 / ms / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
