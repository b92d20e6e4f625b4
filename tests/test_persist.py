"""Row f2: source-model files are interchangeable with the reference's CLI drivers -- class path
`scRNA.nmf_clustering.NmfClustering_initW`, attribute names -- in both directions.  CPU only; the reference side
runs in a subprocess (its own `scRNA` package must own the class path there)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

needs_ref = pytest.mark.skipif(not refshim.reference_available(), reason="no reference tree (mount or oracle/_ref)")


def _fitted_like(cls, g=12, n=9, k=3):
    rng = np.random.RandomState(0)
    obj = cls(rng.poisson(3.0, size=(g, n)).astype(float), np.array(["g%d" % i for i in range(g)]), k,
              rng.randint(0, k, size=n))
    obj.pre_processing()
    obj.dictionary = rng.rand(g, k)
    obj.data_matrix = rng.rand(k, n)
    obj.cluster_labels = np.argmax(obj.data_matrix, axis=0)
    return obj


REF_READS = r"""
import sys
sys.path.insert(0, %(root)r)
from oracle import refshim
refshim.install()
import numpy as np
import scRNA.nmf_clustering as rn
src = np.load(%(path)r, allow_pickle=True)['src'][()]
assert type(src) is rn.NmfClustering_initW, type(src)
src.cell_filter_list = list(); src.gene_filter_list = list()
src.add_cell_filter(lambda x: np.arange(x.shape[1]).tolist())
src.add_gene_filter(lambda x: np.arange(x.shape[0]).tolist())
src.set_data_transformation(lambda x: x)
pp = src.pre_processing()
print("OK", src.dictionary.shape, src.data_matrix.shape, pp.shape, src.num_cluster, float(src.dictionary.sum()))
"""

REF_WRITES = r"""
import sys
sys.path.insert(0, %(root)r)
from oracle import refshim
refshim.install()
import numpy as np
import scRNA.nmf_clustering as rn
rng = np.random.RandomState(1)
obj = rn.NmfClustering_initW(rng.poisson(3.0, size=(10, 7)).astype(float), np.arange(10), 2, rng.randint(0, 2, size=7))
obj.pre_processing()
obj.dictionary = rng.rand(10, 2); obj.data_matrix = rng.rand(2, 7); obj.cluster_labels = np.argmax(obj.data_matrix, 0)
obj.cell_filter_list = None; obj.gene_filter_list = None; obj.data_transf = None
np.savez(%(path)r, src=obj, args=None, allow_pickle=True)
print("OK", float(obj.dictionary.sum()))
"""


def test_alias_refuses_to_shadow_a_real_reference_import(monkeypatch):
    import types
    from scrna_b200 import persist
    if persist._alias:
        pytest.skip("alias already installed in this process")
    monkeypatch.setitem(sys.modules, "scRNA", types.ModuleType("scRNA"))
    with pytest.raises(RuntimeError):
        persist.reference_alias()


OURS_WRITES_AND_READS = r"""
import sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import numpy as np
from test_persist import _fitted_like
from scrna_b200 import NmfClustering_initW
from scrna_b200.persist import save_source_model, load_source_model
obj = _fitted_like(NmfClustering_initW)
save_source_model(%(path)r, obj)
assert type(obj) is NmfClustering_initW and obj.cell_filter_list == []      # the live object is untouched
back = load_source_model(%(path)r)
assert type(back) is NmfClustering_initW
assert np.array_equal(back.dictionary, obj.dictionary) and np.array_equal(back.pre_processing(), obj.pp_data)
print("OK", repr(float(obj.dictionary.sum())))
"""

OURS_READS = r"""
import sys
sys.path.insert(0, %(root)r)
from scrna_b200 import NmfClustering_initW
from scrna_b200.persist import load_source_model
src = load_source_model(%(path)r)
assert type(src) is NmfClustering_initW and isinstance(src, NmfClustering_initW)
assert src.dictionary.shape == (10, 2) and src.data_matrix.shape == (2, 7) and src.num_cluster == 2
assert src.pre_processing().shape == (10, 7)                                 # pass-through filters re-armed
print("OK", repr(float(src.dictionary.sum())))
"""


def _run(script, **kw):
    out = subprocess.run([sys.executable, "-c", script % dict(root=ROOT, **kw)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


@needs_ref
def test_model_written_here_is_read_by_the_unmodified_reference(tmp_path):
    # every side in its own process: the class path scRNA.* belongs to whichever package is imported there
    path = str(tmp_path / "src_c3.npz")
    ours = _run(OURS_WRITES_AND_READS, path=path)
    theirs = _run(REF_READS, path=path)
    assert theirs.startswith("OK (12, 3) (3, 9) (12, 9) 3"), theirs
    assert abs(float(theirs.split()[-1]) - float(ours.split()[-1])) < 1e-12


@needs_ref
def test_model_written_by_the_reference_loads_into_our_classes(tmp_path):
    path = str(tmp_path / "ref_c2.npz")
    theirs = _run(REF_WRITES, path=path)
    ours = _run(OURS_READS, path=path)
    assert abs(float(ours.split()[-1]) - float(theirs.split()[-1])) < 1e-12
