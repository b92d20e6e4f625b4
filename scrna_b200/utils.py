"""Hot-path helpers with the reference's names and signatures (scRNA/utils.py:187-278).

Only the functions on the north-star path live here: `get_transferred_data_matrix` (:187-230) and
`get_matching_gene_inds` (:233-278).  KTA kernels, TSV loaders and the transferability score
(:11-184) are evaluation / IO and out of scope.
"""
import numpy as np

from .transfer import TransferSession


def transfer_h_init(k, n_cells):
    """H init of the reference (utils.py:189-193): |N(0,1)| clamped to >= 1e-10, drawn with ONE
    np.random.randn(k, n) call on the global legacy RNG so the stream matches."""
    H = np.random.randn(k, n_cells)
    H[H < 0.] *= -1.
    H[H < 1e-10] = 1e-10
    return H


def get_transferred_data_matrix(W, trg_data, normalize_H2=False, max_iter=100, rel_err=1e-3,
                                H_init=None, _session=None):
    """Solve trg_data ~ W.H for H >= 0 with W fixed (multiplicative updates on the GPU).

    Returns (W, H, H2, err) exactly like the reference: H2 is the one-hot argmax of H and err the last
    mean absolute reconstruction error.  `H_init` (k x n_t) injects the start instead of drawing it.
    """
    W = np.asarray(W, dtype=np.float64)
    k = W.shape[1]
    sess = _session if _session is not None else TransferSession(W, trg_data)
    n_t = sess.X.n
    H0 = transfer_h_init(k, n_t) if H_init is None else np.asarray(H_init, dtype=np.float64)
    _, new_err = sess.solve(H0, max_iter=max_iter, rel_err=rel_err)
    H = sess.H_host()
    labels = sess.labels()
    H2 = np.zeros((k, n_t))
    H2[(labels, np.arange(n_t))] = 1
    if normalize_H2:
        # second loop of the reference (utils.py:213-229): the same update applied to H2
        keep = (sess.H_host(), sess.n_iter, sess.err)
        sess.solve(H2, max_iter=max_iter, rel_err=rel_err)
        H2 = sess.H_host()
        sess._set_H(keep[0])
        sess.n_iter, sess.err = keep[1], keep[2]
        sess.labels()
    return W, H, H2, new_err


def get_matching_gene_inds(src_gene_ids, trg_gene_ids):
    """Order-preserving intersection of gene ids in TARGET order; the first occurrence of a
    duplicated id wins on both sides (utils.py:233-278).  Returns (inds into trg, inds into src).
    Dictionary-based, O(G) instead of the reference's O(G^2) scans, same result."""
    src_gene_ids = np.asarray(src_gene_ids)
    trg_gene_ids = np.asarray(trg_gene_ids)
    if not np.unique(src_gene_ids).size == src_gene_ids.size:
        print(('\nWarning! (MTL gene ids) Gene ids are supposed to be unique. '
               'Only {0} of {1}  entries are unique.'.format(np.unique(src_gene_ids).shape[0],
                                                             src_gene_ids.shape[0])))
        print('Only first occurance will be used.\n')
    if not np.unique(trg_gene_ids).size == trg_gene_ids.size:
        print(('\nWarning! (Target gene ids) Gene ids are supposed to be unique. '
               'Only {0} of {1}  entries are unique.'.format(np.unique(trg_gene_ids).shape[0],
                                                             trg_gene_ids.shape[0])))
        print('Only first occurance will be used.\n')
    first_src, first_trg = {}, {}
    for i, s in enumerate(src_gene_ids.tolist()):
        first_src.setdefault(s, i)
    for i, s in enumerate(trg_gene_ids.tolist()):
        first_trg.setdefault(s, i)
    inds1, inds2 = [], []
    for s in trg_gene_ids.tolist():
        if s in first_src:
            inds1.append(first_trg[s])
            inds2.append(first_src[s])
    assert len(inds1) > 0
    return np.asarray(inds1, dtype=int), np.asarray(inds2, dtype=int)
