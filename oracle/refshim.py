"""Compatibility shim that lets the UNMODIFIED reference (nicococo/scRNA, mounted read-only at
/root/reference) import and run under the library versions installed in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under scrna_b200/ may import this module.  It is used by
`tests/golden/make_golden.py` (run once, in the build container, to produce the committed fixtures)
and by the `-m "not gpu"` tests that re-validate the oracle restatement against the live reference
when /root/reference happens to be present.  It can not travel to the GPU box (no /root/reference
there), which is why the goldens are committed.

Shim items:
  1. np.float / np.int / np.str aliases (removed in NumPy >= 1.24; used e.g. abstract_clustering.py:28)
  2. stub matplotlib + matplotlib.pyplot modules (scRNA/utils.py:6 imports pyplot at module scope)
  3. decomp.NMF subclass accepting `alpha=` with the era semantics
        l1_W = l1_H = alpha * l1_ratio ;  l2_W = l2_H = alpha * (1 - l1_ratio)      (unscaled)
     (nmf_clustering.py:25-29 passes alpha=, removed in sklearn 1.2)
  4. decomp.non_negative_factorization wrapper: drops alpha/l1_ratio (era default
     regularization=None => no penalty when called as a function) and converts a DataFrame H
     into a C-contiguous float64 ndarray (nmf_clustering.py:64-66)
  5. cluster.KMeans wrapper dropping precompute_distances / n_jobs (sc3_clustering_impl.py:215-216)
  6. np.load default allow_pickle=True (cmd_target.py:188-189)
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """The mounted reference tree, else the copy oracle/make_ref.py pip-installed into oracle/_ref/ (that one
    travels to the GPU box; /root/reference does not)."""
    env = os.environ.get("SCRNA_REFERENCE_ROOT")
    for root in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if os.path.isdir(os.path.join(root, "scRNA")):
            return root
    return env or "/root/reference"


REFERENCE_ROOT = _find_reference()

_installed = False


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scRNA"))


def install():
    """Patch the process so that `import scRNA` resolves to the unmodified reference."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    # (1) removed numpy aliases
    for name, typ in (("float", float), ("int", int), ("str", str), ("bool", bool)):
        if name not in np.__dict__:
            setattr(np, name, typ)

    # (2) matplotlib stubs
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        def _noop(*a, **k):
            return None

        plt.__getattr__ = lambda name: _noop
        mpl.use = _noop
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt

    import sklearn.decomposition as decomp
    import sklearn.decomposition._nmf as _nmf
    import sklearn.cluster as cluster

    # (3) NMF with era `alpha=` semantics
    _BaseNMF = _nmf.NMF

    class EraNMF(_BaseNMF):
        def __init__(self, n_components=None, init=None, solver="cd", beta_loss="frobenius",
                     tol=1e-4, max_iter=200, random_state=None, alpha=0.0, l1_ratio=0.0,
                     verbose=0, shuffle=False):
            super().__init__(n_components=n_components, init=init, solver=solver,
                             beta_loss=beta_loss, tol=tol, max_iter=max_iter,
                             random_state=random_state, alpha_W=0.0, alpha_H="same",
                             l1_ratio=l1_ratio, verbose=verbose, shuffle=shuffle)
            self.alpha = alpha

        def _compute_regularization(self, X):
            l1 = float(self.alpha) * float(self.l1_ratio)
            l2 = float(self.alpha) * (1.0 - float(self.l1_ratio))
            return l1, l1, l2, l2

        @classmethod
        def _get_param_names(cls):
            return sorted(set(_BaseNMF._get_param_names()) | {"alpha"})

    EraNMF._parameter_constraints = dict(_BaseNMF._parameter_constraints)
    EraNMF._parameter_constraints["alpha"] = "no_validation"
    EraNMF.__name__ = "NMF"
    decomp.NMF = EraNMF

    # (4) non_negative_factorization without regularisation
    _base_nnf = _nmf.non_negative_factorization

    def era_nnf(X, W=None, H=None, n_components=None, *, init=None, update_H=True, solver="cd",
                beta_loss="frobenius", tol=1e-4, max_iter=200, alpha=0.0, l1_ratio=0.0,
                regularization=None, random_state=None, verbose=0, shuffle=False):
        if H is not None and not isinstance(H, np.ndarray):
            H = np.ascontiguousarray(np.asarray(H, dtype=np.float64))
        if W is not None and not isinstance(W, np.ndarray):
            W = np.ascontiguousarray(np.asarray(W, dtype=np.float64))
        if regularization is not None:
            raise NotImplementedError("shim only covers the era default regularization=None")
        return _base_nnf(X, W=W, H=H, n_components=n_components, init=init, update_H=update_H,
                         solver=solver, beta_loss=beta_loss, tol=tol, max_iter=max_iter,
                         alpha_W=0.0, alpha_H="same", l1_ratio=0.0, random_state=random_state,
                         verbose=verbose, shuffle=shuffle)

    decomp.non_negative_factorization = era_nnf

    # (5) KMeans without removed kwargs
    _BaseKMeans = cluster.KMeans

    def era_kmeans(n_clusters=8, init="k-means++", n_init=10, max_iter=300, tol=1e-4,
                   precompute_distances="auto", verbose=0, random_state=None, copy_x=True,
                   n_jobs=None, algorithm="lloyd"):
        return _BaseKMeans(n_clusters=n_clusters, init=init, n_init=n_init, max_iter=max_iter,
                           tol=tol, verbose=verbose, random_state=random_state, copy_x=copy_x,
                           algorithm=algorithm)

    cluster.KMeans = era_kmeans

    # (6) np.load default allow_pickle=True
    _np_load = np.load

    def load(*a, **k):
        k.setdefault("allow_pickle", True)
        return _np_load(*a, **k)

    np.load = load

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def import_reference():
    """Returns the reference's hot-path modules (nmf_clustering, sc3_clustering, sc3_clustering_impl,
    utils, simulation) after installing the shim."""
    install()
    import importlib
    mods = {}
    for name in ("abstract_clustering", "utils", "nmf_clustering", "sc3_clustering",
                 "sc3_clustering_impl", "simulation"):
        mods[name] = importlib.import_module("scRNA." + name)
    return types.SimpleNamespace(**mods)
