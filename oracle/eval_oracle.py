"""TEST INFRASTRUCTURE: CPU restatement of the reference's kernel-target-alignment score (row f3 of
SURVEY.md §8), scRNA/utils.py:66-152.  Only tests/ may import this file.

`unsupervised_acc_kta` (utils.py:134-152) builds the linear kernel Kx = X^T X of the cells and the label kernel
Ky = Y Y^T, centres both in feature space (`center_kernel`, :79-83: K - 1K - K1 + 1K1), normalises them to unit
diagonal (`normalize_kernel`, :66-76; identity when a diagonal entry is not finite or <= 1e-16) and returns
<Kx, Ky>_F / sqrt(<Kx, Kx>_F <Ky, Ky>_F) (`kta_align_general`, :86-95 -- there via three n^3 products).
The restatement uses the closed forms (row means) and element-wise Frobenius products: same arithmetic up to
rounding, pinned against the reference's own outputs in tests/golden/kta.npz."""
import numpy as np


def center_kernel(K):
    rm = K.mean(axis=1, keepdims=True)
    cm = K.mean(axis=0, keepdims=True)
    return K - rm - cm + K.mean()


def normalize_kernel(K):
    a = np.sqrt(np.diag(K))
    if np.any(np.isnan(a)) or np.any(np.isinf(a)) or np.any(np.abs(a) <= 1e-16):
        return K * np.eye(K.shape[0])
    b = 1.0 / a
    return K * np.outer(b, b)


def unsupervised_acc_kta(X, labels, center=True, normalize=True):
    X = np.asarray(X, dtype=np.float64)
    labels = np.asarray(labels)
    Y = np.zeros((labels.size, int(np.max(labels)) + 1))
    Y[np.arange(labels.size), labels] = 1.0
    Kx = X.T @ X
    Ky = Y @ Y.T
    if center:
        Kx, Ky = center_kernel(Kx), center_kernel(Ky)
    if normalize:
        Kx, Ky = normalize_kernel(Kx), normalize_kernel(Ky)
    return float((Kx * Ky).sum() / np.sqrt((Kx * Kx).sum() * (Ky * Ky).sum()))
