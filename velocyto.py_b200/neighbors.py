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


# --------------------------------------------------------------------------- balanced kNN graph (host side)
# The greedy hub-limiting pass of velocyto/neighbors.py:13-183 is inherently sequential (each accepted edge changes
# the in-degree budget seen by every later node), so -- like the reference -- it runs on the host under numba.
# Restated here, not copied: one routine for the constrained and unconstrained cases.
_BALANCE_JIT = None


def _balance_jit():
    global _BALANCE_JIT
    if _BALANCE_JIT is None:
        from numba import njit

        @njit(cache=False)
        def greedy(cand, cand_d, order, groups, use_groups, maxl, k):
            n, sight = cand.shape
            nbr = np.full((n, k + 1), -1, np.int64)
            nbr_d = np.zeros((n, k + 1), np.float64)
            load = np.zeros(n, np.int64)                  # edges already pointing AT each node
            for t in range(n):
                node = order[t]
                nbr[node, 0] = node                       # column 0 is the node itself
                taken = 0
                exhausted = True
                for j in range(sight):
                    if taken >= k:
                        exhausted = False
                        break
                    m = cand[node, j]
                    if m == node:
                        continue
                    if use_groups and groups[node] != groups[m]:
                        continue                          # connectivity only inside a group
                    if load[m] >= maxl:
                        continue                          # m is already the neighbour of maxl nodes
                    taken += 1
                    nbr[node, taken] = m
                    nbr_d[node, taken] = cand_d[node, j]
                    load[m] += 1
                if exhausted:
                    while taken < k:                      # sight exhausted: pad with the node itself at distance
                        taken += 1                        # cand_d[node, 0] (0 -> dropped again by `knn > 0`)
                        nbr[node, taken] = node
                        nbr_d[node, taken] = cand_d[node, 0]
            return nbr_d, nbr, load

        _BALANCE_JIT = greedy
    return _BALANCE_JIT


def knn_balance(dsi: np.ndarray, dist: np.ndarray = None, maxl: int = 200, k: int = 60, constraint: np.ndarray = None):
    """Balance a kNN candidate list so that no node is the neighbour of more than ``maxl`` others
    (velocyto/neighbors.py:143-183).  Nodes are visited from the most to the least requested one."""
    dsi = np.ascontiguousarray(dsi, dtype=np.int64)
    assert dsi.shape[1] >= k, "sight needs to be bigger than k"
    indeg = np.bincount(dsi.ravel(), minlength=dsi.shape[0])
    order = np.argsort(indeg, kind="mergesort")[::-1].astype(np.int64)          # neighbors.py:169-170
    want_dist = dist is not None
    if dist is None:
        dist = np.ones(dsi.shape, dtype=np.float64)
        dist[:, 0] = 0
    groups = np.zeros(dsi.shape[0], np.int64) if constraint is None else np.asarray(constraint).astype(np.int64)
    nbr_d, nbr, load = _balance_jit()(dsi, np.ascontiguousarray(dist, dtype=np.float64), np.ascontiguousarray(order),
                                      groups, constraint is not None, int(maxl), int(k))
    if not want_dist:
        nbr_d = np.ones_like(nbr, np.float64)
    return nbr_d, nbr, load


class BalancedKNN:
    """Greedy balanced k-nearest-neighbour graph with the reference's scikit-learn-like API
    (velocyto/neighbors.py:186-321): ``fit`` -> ``kneighbors`` / ``kneighbors_graph``."""

    def __init__(self, k: int = 50, sight_k: int = 100, maxl: int = 200, constraint: np.ndarray = None,
                 mode: str = "distance", metric: str = "euclidean", n_jobs: int = 4, search: str = "device") -> None:
        # search="device": candidate lists from the brute-force GPU kernel (Euclidean); "host": scikit-learn, as the
        # reference does (explicit opt-in, also what the correlation metric uses)
        self.k, self.sight_k, self.maxl, self.mode, self.metric, self.n_jobs = k, sight_k, maxl, mode, metric, n_jobs
        self.search = search
        self.constraint = constraint
        self.dist_new = self.dsi_new = self.l = None
        self.bknn = None
        self._nn = None

    @property
    def n_samples(self) -> int:
        return self.data.shape[0]

    def _device_metric(self):
        """The metric as the device kernel knows it, or None when only scikit-learn implements it."""
        from . import device as dev
        if self.search != "device" or self.metric not in dev.KNN_DEVICE_METRICS:
            return None
        return "euclidean" if self.metric in ("l2", "minkowski") else self.metric

    @property
    def nn(self):
        """scikit-learn searcher (neighbors.py:239-243), built on first use: the device search never needs it."""
        if self._nn is None:
            from sklearn.neighbors import NearestNeighbors
            if self.metric == "correlation":
                self._nn = NearestNeighbors(n_neighbors=self.sight_k + 1, metric=self.metric, n_jobs=self.n_jobs,
                                            algorithm="brute")
            else:
                self._nn = NearestNeighbors(n_neighbors=self.sight_k + 1, metric=self.metric, n_jobs=self.n_jobs,
                                            leaf_size=30)
            self._nn.fit(self.fitdata)
        return self._nn

    def fit(self, data: np.ndarray, sight_k: int = None):
        self.data = self.fitdata = data
        if sight_k is not None:
            self.sight_k = sight_k
        self._nn = None
        return self

    def kneighbors(self, X: np.ndarray = None, maxl: int = None, mode: str = "distance"):
        if X is not None:
            self.data = X
        if maxl is not None:
            self.maxl = maxl
        from . import device as dev
        dev_metric = self._device_metric()
        if (dev_metric is not None and self.sight_k + 1 <= dev.KNN_MAX_K and self.data is self.fitdata
                and np.shape(self.fitdata)[1] <= 4096):
            # candidate lists (self included, ascending distance) from the brute-force device kernel
            idx, dist = dev.knn(np.ascontiguousarray(self.fitdata, dtype=np.float64), self.sight_k + 1, include_self=True,
                                metric=dev_metric)
            self.dist, self.dsi = dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
        else:
            # any other scikit-learn metric ("manhattan", ...), queries other than the fitted data, oversized k
            self.dist, self.dsi = self.nn.kneighbors(self.data, return_distance=True)
        self.dist_new, self.dsi_new, self.l = knn_balance(self.dsi, self.dist, maxl=self.maxl, k=self.k,
                                                          constraint=self.constraint)
        return self.dist_new, self.dsi_new, self.l

    def kneighbors_graph(self, X: np.ndarray = None, maxl: int = None, mode: str = "distance") -> sparse.csr_matrix:
        dist_new, dsi_new, _ = self.kneighbors(X=X, maxl=maxl, mode=mode)
        n, w = dist_new.shape
        self.bknn = sparse.csr_matrix((np.ravel(dist_new), np.ravel(dsi_new), np.arange(0, n * w + 1, w)), (n, n))
        return self.bknn
