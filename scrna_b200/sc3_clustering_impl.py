"""SC3 building blocks with the reference's names and signatures (scRNA/sc3_clustering_impl.py).

Host-side pieces (O(g*n) once): `cell_filter` (:32-42), `gene_filter` (:45-58),
`data_transformation_log2` (:61-67), `no_data_transformation` (:70-76).
Device pieces are added below as the SC3 kernels land (distances, transformations, k-means ensemble,
consensus matrix, consensus clustering).
"""
import numpy as np


DEVICE_FILTER_MIN_ELEMENTS = 1 << 24   # below this the NumPy passes cost milliseconds


def _filter_counts(data, non_zero_threshold):
    """(col_ge, row_ge, row_gt0) of the genes x cells matrix: counts of entries >= threshold per cell and per gene,
    and of entries > 0 per gene, NaNs zeroed IN PLACE in the caller's matrix as the reference does (:38-39, :51-52).
    Large matrices on a GPU box stream through the device in row bands (scrna_filter_counts_f64: one pass, integer
    counts -- identical to NumPy's); the host matrix is only touched when a NaN was actually seen."""
    big = data.size >= DEVICE_FILTER_MIN_ELEMENTS and data.dtype == np.float64 and data.ndim == 2
    if big:
        try:
            import torch
            big = torch.cuda.is_available()
        except Exception:
            big = False
    if not big:
        data[np.isnan(data)] = 0
        return (np.sum(data >= non_zero_threshold, axis=0), np.sum(data >= non_zero_threshold, axis=1),
                np.sum(data > 0, axis=1))
    from . import _lib
    ops = _ops()
    t = ops.torch
    g, n = data.shape
    i32 = t.int32
    col_ge = ops.zeros(n, dtype=i32)
    row_ge = ops.zeros(g, dtype=i32)
    row_gt0 = ops.zeros(g, dtype=i32)
    flag = ops.zeros(1, dtype=i32)
    rows_per_band = max(1, int((256 << 20) // max(1, n * 8)))
    for r0 in range(0, g, rows_per_band):
        r1 = min(g, r0 + rows_per_band)
        band = t.from_numpy(np.ascontiguousarray(data[r0:r1])).to(ops.dev)
        ops._call("scrna_filter_counts_f64", band, n, r1 - r0, n, float(non_zero_threshold), col_ge, row_ge[r0:r1],
                  row_gt0[r0:r1], flag, ops._s())
    if int(flag.cpu()[0]):
        data[np.isnan(data)] = 0
    return (col_ge.cpu().numpy().astype(np.int64), row_ge.cpu().numpy().astype(np.int64),
            row_gt0.cpu().numpy().astype(np.int64))


def cell_filter(data, num_expr_genes=2000, non_zero_threshold=2):
    """Indices of cells expressing at least `num_expr_genes` genes at >= non_zero_threshold.
    NaNs are zeroed IN PLACE in the caller's matrix, as the reference does (:38-39)."""
    res, _, _ = _filter_counts(data, non_zero_threshold)
    return np.where(np.isfinite(res) & (res >= num_expr_genes))[0]


def gene_filter(data, perc_consensus_genes=0.94, non_zero_threshold=2):
    """Indices of genes that are neither rare nor ubiquitous (:45-58): at least n*(1-p) cells at
    >= threshold and at most n*p cells above zero.  NaNs zeroed in place."""
    _, res_l, res_h = _filter_counts(data, non_zero_threshold)
    num_cells = data.shape[1]
    lower_bound = float(num_cells) * (1. - perc_consensus_genes)
    upper_bound = float(num_cells) * perc_consensus_genes
    return np.where((res_l >= lower_bound) & (res_h <= upper_bound))[0]


def data_transformation_log2(data):
    return np.log2(data + 1.)


def no_data_transformation(data):
    return data


# --------------------------------------------------------------------------------------------------
# Device stages.  Same signatures and NumPy-in / NumPy-out contract as the reference callables
# (SC3Clustering's plug-in API, scRNA/sc3_clustering.py:34-56).  A result also stays resident on the
# device, keyed by the identity of the returned array, so the next stage does not upload it again.
# --------------------------------------------------------------------------------------------------
import weakref

_resident = {}


def _remember(host_array, device_tensor):
    """Pair a stage's host result with its device twin.  The host array is made READ-ONLY: the twin is found again
    by the array's identity, so an in-place edit between two stages (`D /= D.max()`, `D[np.isnan(D)] = 0`, ...)
    would otherwise make the next stage run silently on the stale device copy.  Such an edit now raises
    `ValueError: assignment destination is read-only`; edit a copy instead (a copy has no twin and is uploaded)."""
    host_array.flags.writeable = False
    key = id(host_array)
    _resident[key] = (weakref.ref(host_array, lambda _r, k=key: _resident.pop(k, None)), device_tensor)
    return host_array


def _lookup(host_array):
    """Device twin of `host_array`, or of the array it is a leading-columns view of, else None."""
    hit = _resident.get(id(host_array))
    if hit is not None and hit[0]() is host_array:
        return hit[1]
    base = getattr(host_array, "base", None)
    if base is not None:
        hit = _resident.get(id(base))
        if hit is not None and hit[0]() is base and host_array.ndim == 2 and base.ndim == 2 \
                and host_array.shape[0] == base.shape[0] and host_array.strides == base.strides \
                and host_array.__array_interface__['data'][0] == base.__array_interface__['data'][0]:
            return hit[1][:, : host_array.shape[1]]
    return None


def _ops():
    from .sc3_device import default_ops
    return default_ops()


class DeviceResult:
    """A stage result that stays in HBM: what this package's own SC3Clustering passes between this package's own
    stage functions (`_device=True`) instead of an n x n NumPy array -- at 20 000 cells every such array is a 3.2 GB
    device-to-host copy the next stage would only upload again.  Never handed to user-supplied callables."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.shape = tuple(tensor.shape)

    def host(self):
        return self.tensor.cpu().numpy()


def _device_of(x):
    """Device tensor of a stage input: a DeviceResult, the resident twin of a remembered host array, or None."""
    if isinstance(x, DeviceResult):
        return x.tensor
    return _lookup(x)


def is_own(fn, target):
    """True when the callable (possibly a functools.partial) is this module's `target`."""
    return getattr(fn, "func", fn) is target


def distances(data, gene_ids, metric='euclidean', _device=False, _comm=None):
    """cells x cells distance matrix of a transcripts x cells matrix (scRNA/sc3_clustering_impl.py:106-132):
    'euclidean' (bit-identical to scipy pdist), 'pearson' (1 - corrcoef), 'spearman' (1 - rank corrcoef)."""
    ops = _ops()
    D = ops.distances(np.asarray(data, dtype=np.float64), metric, comm=_comm)
    if _device:
        return DeviceResult(D)
    return _remember(D.cpu().numpy(), D)


def da_nmf_distances(data, gene_ids, da_model, reject_ratio=0., metric='euclidean', mixture=0.5, use_H2=True):
    """Convex combination of the distances of the data and of its dictionary reconstruction W.H2 (or W.H), the
    latter rescaled to the former's maximum (scRNA/sc3_clustering_impl.py:79-103; unused by the reference's CLI).
    Both distance matrices come from the device kernels; the returned array is a fresh host array."""
    if mixture == 0.0:
        return distances(data, [], metric=metric)
    W, H, H2 = da_model
    dist1 = distances(data, [], metric=metric)
    if use_H2:
        dist2 = distances(W.dot(H2), [], metric=metric)
    else:
        print('Using H instead of H2 for reconstruction.')
        dist2 = distances(W.dot(H), [], metric=metric)
    scale = 1.0
    if np.max(dist2) < 1e-12:
        if mixture == 1.0:
            raise Exception('Distances are all close to zero and mixture=1.0. '
                            'Seems that source and target data do not go well together.')
        else:
            print('Warning! Max distance is close to zero.')
    else:
        scale = np.max(dist1) / np.max(dist2)
    return mixture * (dist2 * scale) + (1. - mixture) * dist1


def transformations(dm, components=5, method='pca', _device=False, _comm=None):
    """(cells x cells matrix rebuilt from the kept components, cells x components right singular vectors)
    of the PCA-scaled or 'spectral' distance matrix (scRNA/sc3_clustering_impl.py:135-203).  The first return
    value is what the reference computes and its only caller discards (sc3_clustering.py:77,95); this package's
    own SC3Clustering asks for it not to be formed (`_device=True`: (None, vectors))."""
    ops = _ops()
    dev = _device_of(dm)
    D = dev if dev is not None else ops.to_device(np.asarray(dm, dtype=np.float64))
    num_cells = D.shape[0]
    M = ops.transform_matrix(D, method)
    V, vals = ops.right_singular_vectors(M, components, comm=_comm)
    vecs = V.cpu().numpy()
    if _device:
        return None, _remember(vecs, V)
    vals = vals / np.sqrt(np.max((1, num_cells - 1)))
    if ops._last is not None:
        rebuilt = ops.rebuilt_matrix(vals[:components]).cpu().numpy()     # n x n Gram on the device
    else:
        rebuilt = vecs.dot(np.diag(vals[:components]).dot(vecs.T))         # n <= 64
    return rebuilt, _remember(vecs, V)


def kmeans_draws(k, n_init=10):
    """The uniforms one KMeans(n_init).fit consumes from the global legacy NumPy RNG, drawn here in the
    reference's order so that seeded runs match it."""
    from .sc3_device import kmeans_draw_count
    per, _ = kmeans_draw_count(k)
    return np.random.random_sample(n_init * per)


def intermediate_kmeans_clustering(X, k=5, n_init=10, max_iter=10000, init='k-means++', _skip=False):
    """cells x 1 labels of KMeans(n_clusters=k, n_init, max_iter, init='k-means++') on the cells x d matrix
    (scRNA/sc3_clustering_impl.py:206-219).

    `_skip=True` (used by SC3Clustering when the k-means ensemble is distributed over several GPUs)
    only consumes the fit's random draws from the global NumPy RNG -- so that every rank stays on the reference's
    RNG stream -- and returns None: the fit itself runs on another rank."""
    if init != 'k-means++':
        raise NotImplementedError("only init='k-means++' (the reference's default and only use) is on the device")
    draws = kmeans_draws(k, n_init)
    if _skip:
        return None
    ops = _ops()
    X = np.asarray(X)
    n, d = X.shape
    dev = _lookup(X)
    V = dev if dev is not None else ops.to_device(np.ascontiguousarray(X, dtype=np.float64))
    # V may be a column-slice view of the resident vectors: its row stride is the leading dimension
    labels = ops.kmeans(V, d, k, draws, n_init=n_init, max_iter=max_iter)
    assert labels.size == n
    return labels


intermediate_kmeans_clustering.supports_skip = True


def build_consensus_matrix(X):
    """cells x cells co-association matrix of an (n_labelings x cells) label matrix, averaged over the
    labelings (scRNA/sc3_clustering_impl.py:222-241)."""
    ops = _ops()
    X = np.asarray(X)
    if X.ndim == 1:
        X = X[np.newaxis, :]
    n, cells = X.shape
    M = ops.coassoc_new(cells)
    for i in range(n):
        ops.coassoc_add(M, X[i])
    C = ops.consensus_from_counts(M, float(n))
    return _remember(C.cpu().numpy(), C)


def consensus_clustering(consensus, n_components=5, _comm=None):
    """(labels, cells x cells Euclidean distances between consensus rows): pdist -> complete linkage ->
    cut_tree(n_components) (scRNA/sc3_clustering_impl.py:244-260)."""
    ops = _ops()
    dev = _device_of(consensus)
    C = dev if dev is not None else ops.to_device(np.asarray(consensus, dtype=np.float64))
    labels, D, _ = ops.consensus_clustering(C, n_components, comm=_comm)
    return labels, D.cpu().numpy()
