"""Source-model persistence compatible with the reference's CLI drivers.

`scRNA/cmd_source.py:208-212` stores the fitted source object with `np.savez(fout, src=nmf, args=arguments)` --
a pickle whose class path is `scRNA.nmf_clustering.NmfClustering_initW` -- and `scRNA/cmd_target.py:188-196`
loads it back with `np.load(src_fname)['src'][()]`.  A drop-in has to read such files and write files the
reference reads:

  * `reference_alias()` registers alias modules `scRNA`, `scRNA.nmf_clustering`, ... whose classes are thin
    subclasses of this package's classes carrying the reference's class path (only when the real reference is not
    imported in this process).  Unpickling a reference-written model then yields objects that ARE
    scrna_b200 classes (same attribute names: dictionary, data_matrix, cluster_labels, pp_data, remain_gene_inds,
    remain_cell_inds, gene_ids, num_cluster, ...).
  * `save_source_model` re-classes the object to the alias before `np.savez`, so the file names
    `scRNA.nmf_clustering.<Class>` and the unmodified reference unpickles it into its own class.
"""
import sys
import types

import numpy as np

_CLASSES = {
    "scRNA.abstract_clustering": ("abstract_clustering", ["AbstractClustering"]),
    "scRNA.nmf_clustering": ("nmf_clustering", ["NmfClustering", "NmfClustering_initW", "NmfClustering_fixW",
                                                "DaNmfClustering"]),
    "scRNA.sc3_clustering": ("sc3_clustering", ["SC3Clustering"]),
}
_alias = {}


def reference_alias():
    """Install (idempotent) the `scRNA.*` alias modules; returns {our class: alias class}."""
    if _alias:
        return _alias
    real = sys.modules.get("scRNA")
    if real is not None and not getattr(real, "__scrna_b200_alias__", False):
        raise RuntimeError("the reference package scRNA is imported in this process; its own classes own the "
                           "class path scRNA.* (load / save the model from a process without it)")
    import importlib
    pkg = types.ModuleType("scRNA")
    pkg.__scrna_b200_alias__ = True
    pkg.__path__ = []
    sys.modules["scRNA"] = pkg
    for mod_name, (ours, names) in _CLASSES.items():
        src = importlib.import_module("scrna_b200." + ours)
        mod = types.ModuleType(mod_name)
        mod.__scrna_b200_alias__ = True
        for name in names:
            base = getattr(src, name)
            cls = type(name, (base,), {"__module__": mod_name, "__doc__": base.__doc__})
            setattr(mod, name, cls)
            _alias[base] = cls
        sys.modules[mod_name] = mod
        setattr(pkg, mod_name.split(".")[1], mod)
    return _alias


def _to_alias(obj):
    alias = reference_alias()
    for base, cls in alias.items():
        if type(obj) is base:
            obj.__class__ = cls
    if getattr(obj, "src", None) is not None:
        _to_alias(obj.src)
    return obj


def _from_alias(obj):
    for base, cls in _alias.items():
        if type(obj) is cls:
            obj.__class__ = base
    if getattr(obj, "src", None) is not None:
        _from_alias(obj.src)
    return obj


def save_source_model(fname, nmf, args=None):
    """cmd_source.py:208-212: null the (unpicklable) filters and transformation, then
    `np.savez(fname, src=nmf, args=args)` with the reference's class path."""
    nmf.cell_filter_list = None
    nmf.gene_filter_list = None
    nmf.data_transf = None
    _to_alias(nmf)
    try:
        np.savez(fname, src=nmf, args=args, allow_pickle=True)
    finally:
        _from_alias(nmf)
    nmf.__setstate__({})                     # back to empty filter lists / the identity transformation


def load_source_model(fname):
    """cmd_target.py:188-196: `np.load(fname)['src'][()]`, the filters re-armed as pass-throughs (the source data
    is stored filtered and transformed).  Accepts files written by the reference or by `save_source_model`."""
    reference_alias()
    with np.load(fname, allow_pickle=True) as f:
        src = f["src"][()]
    _from_alias(src)
    src.cell_filter_list = []
    src.gene_filter_list = []
    src.add_cell_filter(lambda x: np.arange(x.shape[1]).tolist())
    src.add_gene_filter(lambda x: np.arange(x.shape[0]).tolist())
    src.set_data_transformation(lambda x: x)
    return src
