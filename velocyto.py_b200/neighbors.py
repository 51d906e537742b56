"""Drop-in for the smoothing half of ``velocyto/neighbors.py`` (same names and signatures).

``connectivity_to_weights`` (neighbors.py:385-390) is graph bookkeeping on ``nnz = cells*(k+1)`` entries
and stays on the host with SciPy, exactly as in the reference.  ``convolve_by_sparse_weights``
(neighbors.py:416-423) -- 68% of ``knn_imputation``'s time in the reference, a single-threaded SciPy
``coo_matmat_dense`` -- runs as the CSR row-gather kernel ``velo_dev_knn_smooth``.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def connectivity_to_weights(mknn, axis: int = 1):
    """Row-normalise a connectivity matrix into smoothing weights (velocyto/neighbors.py:385-390)."""
    if not isinstance(mknn, sparse.csr_matrix):
        mknn = sparse.csr_matrix(mknn)
    return mknn.multiply(1. / sparse.csr_matrix.sum(mknn, axis=axis))


def _csr_rows(w):
    """CSR by rows of ``w`` (cells x cells): row c lists the cells averaged into smoothed cell c."""
    w = sparse.csr_matrix(w)
    w.sum_duplicates()
    return w.indptr.astype(np.int64), w.indices.astype(np.int32), w.data.astype(np.float32), w


def convolve_by_sparse_weights(data: np.ndarray, w) -> np.ndarray:
    """``data (genes x cells) . w^T`` on the GPU (velocyto/neighbors.py:416-423).

    Like the reference's SciPy product, the result is an F-contiguous ``(genes, cells)`` float64 array
    (physically cell-major)."""
    from . import device as dev
    indptr, indices, weights, wcsr = _csr_rows(w)
    assert np.allclose(np.asarray(wcsr.sum(1)).ravel(), 1), "weight matrix need to sum to one over the columns"
    data = np.asarray(data)
    if data.shape[1] != wcsr.shape[0]:
        raise ValueError("dimension mismatch")
    S_cm = dev.CellMajor.from_gene_major(data)
    out = dev.knn_smooth(indptr, indices, weights, S_cm)
    return out.t[:, :out.G].to("cpu").numpy().astype(np.float64).T      # (G x C) view of a C-order (C x G) array
