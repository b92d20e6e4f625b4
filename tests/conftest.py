import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def default_sim(golden):
    d = golden("default_sim")
    return dict(src=d["src"].astype(np.float64), trg=d["trg"].astype(np.float64),
                src_labels=d["src_labels"], trg_labels=d["trg_labels"], gene_ids=np.arange(1000))


@pytest.fixture(scope="session")
def kernels():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from scrna_b200.kernels import default_kernels
    return default_kernels()


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
