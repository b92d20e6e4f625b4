"""bench.py's synthetic generator tiles the SAME global matrix for every sharding (the 1/2/4/8-GPU runs factorise
identical data), and the shard bounds land on its cell-block boundaries."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_synthetic_matrix_is_identical_across_shardings():
    import bench
    from scrna_b200.nmf_solver import shard_bounds
    cpu = torch.device("cpu")
    genes, cells = 12, 4 * bench.CELL_BLOCK
    full, n = bench.synth_counts(torch, cpu, genes, 0, cells, rows_per_band=5)
    assert n == cells and float(full[:, :cells].sum()) > 0
    for world in (2, 4):
        parts = []
        for r in range(world):
            c0, c1 = shard_bounds(cells, world, r)
            assert c0 % bench.CELL_BLOCK == 0
            xs, ns = bench.synth_counts(torch, cpu, genes, c0, c1, rows_per_band=5)
            parts.append(xs[:, :ns])
        assert torch.equal(torch.cat(parts, dim=1), full[:, :cells])
    # a shard that starts inside a block still reproduces the global columns
    xs, ns = bench.synth_counts(torch, cpu, genes, 1000, 30000, rows_per_band=5)
    assert torch.equal(xs[:, :ns], full[:, 1000:30000])
    assert np.isfinite(full.numpy()).all()
