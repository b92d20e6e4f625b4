"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/scrna_b200.h declares (no compute calls -- there is no GPU in this tier of the suite)."""
import ctypes
import os
import re

import pytest

from scrna_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scrna_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint(?:64_t)?\s+(scrna_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    syms = declared_symbols()
    assert len(syms) >= 20
    assert "scrna_nmf_xht_partial" in syms and "scrna_nmf_wtx_partial" in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in include/scrna_b200.h but not exported: %s" % missing


def test_python_binding_covers_header():
    assert sorted(_lib.exported_symbols()) == declared_symbols()
    lib = _lib.load()
    assert lib.scrna_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scrna_b200.kernels import CudaKernels
    with pytest.raises(_lib.ScrnaError):
        CudaKernels()
    import numpy as np
    from scrna_b200 import NmfClustering
    nmf = NmfClustering(np.ones((4, 5)), np.arange(4), 2)
    with pytest.raises(_lib.ScrnaError):
        nmf.apply()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scrna_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
