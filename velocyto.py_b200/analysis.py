"""Drop-in for the hot methods of ``velocyto.analysis.VelocytoLoom`` (velocyto/analysis.py).

Same method names, keyword arguments, attribute names and shapes as the reference for the path

    knn_imputation -> fit_gammas -> predict_U -> calculate_velocity -> calculate_shift ->
    extrapolate_cell_at_t -> estimate_transition_prob -> calculate_embedding_shift

(the canonical order of doc/tutorial/analysis.rst:108-165).  Method bodies stay Python; all array
math runs in ``libvelo_b200.so`` on matrices that stay resident in HBM between methods as cell-major
fp32 (``device.CellMajor``).  The big ``(genes, cells)`` attributes (``Sx_sz``, ``Upred``, ``velocity``
...) are materialised as float64 NumPy arrays only when they are read, so ``to_hdf5``-style consumers
and plotting code keep working while a 100k-cell pipeline never round-trips through the host.

What stays on the host, as in the reference: the neighbour sampler and the randomised control BY DEFAULT (NumPy /
numba RNG streams must match the reference's for identical ``sampling_ixs`` / ``delta_S_rndm``, analysis.py:1552-1566,
2407-2420 -- the sampler runs as a C++ restatement of the NumPy stream, ``csrc/host_sampler.cpp``;
``random_backend="device"`` moves both to the GPU with their own streams), BalancedKNN's greedy pass (sequential by
construction) and the O(nnz) graph bookkeeping.  The kNN searches run on the device (exact, brute force).

Also mirrored (SURVEY.md 8f "next" rows): ``normalize`` (analysis.py:535-676), ``perform_PCA`` (:678-702),
``calculate_grid_arrows`` (:1735-1816), and sparse ingest -- ``VelocytoLoom(S=<scipy sparse>, U=...)`` /
``VelocytoLoom.from_csr(...)`` keep the count layers sparse on the host and upload them as CSR by cell.
Out of scope: loom / HDF5 file parsing, filtering, TSNE, plotting, Markov diffusion.
The dense ``(cells, cells)`` results (``corrcoef``, ``transition_prob`` and their ``_random`` twins) are built lazily, the
first time they are read, from the compact ``(cells, m)`` device results (``*_compact`` + ``neigh_ixs``).
"""
from __future__ import annotations

import logging
import warnings
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
from scipy import sparse

from . import _cabi
from .neighbors import connectivity_to_weights

# (genes, cells) matrices that live on the device once a hot method has produced or consumed them
_MATRIX_ATTRS = ("S", "U", "S_sz", "U_sz", "Sx", "Ux", "Sx_sz", "Ux_sz", "Upred", "velocity", "delta_S",
                 "Sx_sz_t", "Sx_t", "delta_S_rndm", "S_norm", "U_norm", "Sx_norm", "Ux_norm")


class _DeviceBacked:
    """Descriptor: a ``(genes, cells)`` float64 attribute with a device-resident cell-major twin."""

    def __init__(self, name: str):
        self.name = name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        host = obj.__dict__.setdefault("_host", {})
        if self.name in host:
            return host[self.name]
        devs = obj.__dict__.setdefault("_devs", {})
        if self.name in devs:
            # a download of device-born data: handed out READ-ONLY, because the device twin is what the later stages
            # read -- an in-place edit of this array would silently not reach it.  Assign a new array instead
            # (``vlm.delta_S = edited``): __set__ drops the device twin and the next stage uploads the new values.
            arr = devs[self.name].to_gene_major(np.float64)
            arr.flags.writeable = False
            host[self.name] = arr
            return arr
        raise AttributeError(f"'{type(obj).__name__}' object has no attribute '{self.name}'")

    def __set__(self, obj, value):
        obj.__dict__.setdefault("_host", {})[self.name] = value
        obj.__dict__.setdefault("_devs", {}).pop(self.name, None)
        obj.__dict__.setdefault("_fp32_exact", set()).discard(self.name)

    def __delete__(self, obj):
        obj.__dict__.setdefault("_host", {}).pop(self.name, None)
        obj.__dict__.setdefault("_devs", {}).pop(self.name, None)


class _LazyDense:
    """Descriptor for the reference's dense ``(cells, cells)`` float64 results (``corrcoef``, ``transition_prob`` and
    their ``_random`` twins).  The native result is the compact ``(cells, m)`` device matrix aligned with ``neigh_ixs``;
    the dense form (800 MB per matrix at 10k cells, most of a second of scatter + D2H + host allocation each) is built
    the first time it is READ, and only when cells <= ``dense_limit`` -- a pipeline that goes on to
    ``calculate_embedding_shift`` / ``calculate_grid_arrows`` never pays for it."""

    def __init__(self, name: str):
        self.name = name

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        cache = obj.__dict__.setdefault("_dense", {})
        if self.name in cache:
            return cache[self.name]
        src = obj.__dict__.get("_dense_src", {}).get(self.name)
        if src is None:
            raise AttributeError(f"'{type(obj).__name__}' object has no attribute '{self.name}'")
        compact, ix, C, limit = src
        if C > limit:
            raise AttributeError(f"{self.name}: the dense ({C} x {C}) matrix is not materialised above dense_limit={limit} "
                                 f"cells; use {self.name}_compact with neigh_ixs")
        from . import device as dev
        cache[self.name] = dev.scatter_dense(compact, ix, C).cpu().numpy()
        return cache[self.name]

    def __set__(self, obj, value):
        obj.__dict__.setdefault("_dense", {})[self.name] = value
        obj.__dict__.setdefault("_dense_src", {}).pop(self.name, None)

    def __delete__(self, obj):
        obj.__dict__.setdefault("_dense", {}).pop(self.name, None)
        obj.__dict__.setdefault("_dense_src", {}).pop(self.name, None)


_DENSE_ATTRS = ("corrcoef", "corrcoef_random", "transition_prob", "transition_prob_random")


def knn_graph_device(data: np.ndarray, k: int, mode: str = "connectivity", include_self: bool = False,
                     metric: str = "euclidean") -> sparse.csr_matrix:
    """kNN graph built by the brute-force device kernel (``velo_dev_knn``): the CSR scikit-learn's
    ``kneighbors_graph(X=None, mode=...)`` returns -- k entries per row in ascending distance."""
    from . import device as dev
    idx, dist = dev.knn(np.ascontiguousarray(data, dtype=np.float64), k, include_self, metric)
    n = idx.shape[0]
    vals = dist.cpu().numpy().ravel() if mode == "distance" else np.ones(n * k)
    return sparse.csr_matrix((vals, idx.cpu().numpy().astype(np.int32).ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))


def knn_distance_matrix(data: np.ndarray, metric: str = None, k: int = 40, mode: str = "connectivity",
                        n_jobs: int = 4) -> sparse.csr_matrix:
    """k nearest neighbours NOT including the point itself (velocyto/neighbors.py:363-376).  The reference passes
    ``metric`` to scikit-learn only when it is "correlation" (otherwise the default Euclidean search); both run on the
    device (exact, brute force).  scikit-learn remains only for k or dimensions beyond the kernel's limits."""
    from . import device as dev
    dev_metric = "correlation" if metric == "correlation" else "euclidean"
    if k <= dev.KNN_MAX_K and np.shape(data)[1] <= 4096:
        return knn_graph_device(data, k, mode, metric=dev_metric)
    from sklearn.neighbors import NearestNeighbors
    if metric == "correlation":
        nn = NearestNeighbors(n_neighbors=k, metric="correlation", algorithm="brute", n_jobs=n_jobs)
    else:
        nn = NearestNeighbors(n_neighbors=k, n_jobs=n_jobs)
    nn.fit(data)
    return nn.kneighbors_graph(X=None, mode=mode)


def connectivity_with_diagonal(knn, diag: float = 1) -> sparse.csr_matrix:
    """``(knn > 0).astype(float)`` followed by ``setdiag(diag)`` (analysis.py:1006-1009), built directly on the CSR
    arrays: SciPy's comparison operator and ``setdiag`` on a CSR matrix take seconds at 5e6 edges, this takes tens of
    milliseconds.  Zero-distance edges are dropped exactly as ``knn > 0`` drops them."""
    knn = sparse.csr_matrix(knn)
    n = knn.shape[0]
    k = knn.indptr[1] - knn.indptr[0] if n else 0
    if n and knn.indptr[-1] == n * k and np.array_equal(knn.indptr, np.arange(0, n * k + 1, k)):
        # regular graph (k entries per row: what the kNN searches return) -- work on (n, k) blocks, no per-edge row ids
        idx2, dat2 = knn.indices.reshape(n, k), knn.data.reshape(n, k)
        if (dat2 > 0).all() and (idx2 != np.arange(n, dtype=idx2.dtype)[:, None]).all():
            indices = np.empty((n, k + 1), dtype=np.int32)
            indices[:, 0] = np.arange(n, dtype=np.int32)          # diagonal first in every row
            indices[:, 1:] = idx2
            data = np.ones((n, k + 1), dtype=np.float64)
            data[:, 0] = diag
            return sparse.csr_matrix((data.ravel(), indices.ravel(), np.arange(0, n * (k + 1) + 1, k + 1)), shape=knn.shape)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(knn.indptr))
    keep = (knn.data > 0) & (knn.indices != rows)                 # the diagonal is overwritten by setdiag anyway
    counts = np.bincount(rows[keep], minlength=n) + 1             # + the diagonal entry
    indptr = np.concatenate([[0], np.cumsum(counts)])
    indices = np.empty(indptr[-1], dtype=np.int32)
    data = np.ones(indptr[-1], dtype=np.float64)
    indices[indptr[:-1]] = np.arange(n, dtype=np.int32)           # diagonal first in every row
    data[indptr[:-1]] = diag
    off = np.ones(indptr[-1], dtype=bool)
    off[indptr[:-1]] = False
    indices[off] = knn.indices[keep]                              # CSR order is preserved row by row
    return sparse.csr_matrix((data, indices, indptr), shape=knn.shape)


def smoothing_weights_from_knn(knn, diag: float = 1) -> sparse.csr_matrix:
    """``connectivity_to_weights(setdiag((knn > 0).astype(float), diag))`` (analysis.py:1006-1010) in one step.

    For the regular graph a kNN search returns (k entries per row, positive distances, no self edges) the result is known
    in closed form -- every row is ``[diag, 1, ..., 1] / (diag + k)`` over ``[c, neighbours of c]`` -- and is written
    directly into its CSR arrays (two passes over the 5e6 edges of BASELINE config 2 instead of the dozen that SciPy's
    comparison, setdiag, row sums and broadcast multiply make).  Anything else goes through the general path."""
    knn = sparse.csr_matrix(knn)
    n = knn.shape[0]
    k = int(knn.indptr[1] - knn.indptr[0]) if n else 0
    if n and k and knn.indptr[-1] == n * k and (np.diff(knn.indptr) == k).all():
        idx2 = knn.indices.reshape(n, k)
        if knn.data.min() > 0 and not (idx2 == np.arange(n, dtype=idx2.dtype)[:, None]).any():
            indices = np.empty((n, k + 1), dtype=np.int32)
            indices[:, 0] = np.arange(n, dtype=np.int32)              # diagonal first in every row
            indices[:, 1:] = idx2
            data = np.empty((n, k + 1), dtype=np.float64)
            data[:, 0] = diag / (diag + k)
            data[:, 1:] = 1.0 / (diag + k)
            return sparse.csr_matrix((data.ravel(), indices.ravel(), np.arange(0, n * (k + 1) + 1, k + 1)), shape=knn.shape)
    connectivity = connectivity_with_diagonal(knn, diag)
    rowsum = np.add.reduceat(connectivity.data, connectivity.indptr[:-1])
    return sparse.csr_matrix((connectivity.data / np.repeat(rowsum, np.diff(connectivity.indptr)), connectivity.indices,
                              connectivity.indptr), shape=connectivity.shape)


class VelocytoLoom:
    """The hot-path subset of ``velocyto.analysis.VelocytoLoom`` (velocyto/analysis.py:26-2342)."""

    def __init__(self, S: np.ndarray = None, U: np.ndarray = None, A: np.ndarray = None,
                 ca: Optional[Dict] = None, ra: Optional[Dict] = None, loom_filepath: str = None) -> None:
        if loom_filepath is not None:
            raise NotImplementedError("loom I/O is out of scope (SURVEY.md section 2): pass the S and U count matrices")
        if S is None or U is None:
            raise ValueError("S and U (genes x cells) are required")
        # SciPy sparse count matrices stay sparse on the host and go to the device as CSR by cell (device.CsrCounts);
        # the reference's loader materialises dense float64 matrices (analysis.py:56-64: 24 GB per layer at 100k x 30k)
        keep = lambda M: sparse.csc_matrix(M, dtype=np.float64) if sparse.issparse(M) else np.asarray(M, dtype=np.float64)
        self.S, self.U = keep(S), keep(U)
        if A is None:
            self.A = sparse.csc_matrix(self.S.shape) if sparse.issparse(self.S) else np.zeros_like(self.S)
        else:
            self.A = A if sparse.issparse(A) else np.asarray(A)
        self.ca, self.ra = dict(ca or {}), dict(ra or {})
        self.initial_cell_size = np.asarray(self.S.sum(0)).ravel()        # analysis.py:62-63
        self.initial_Ucell_size = np.asarray(self.U.sum(0)).ravel()

    @classmethod
    def from_csr(cls, S: Tuple, U: Tuple, n_genes: int, A: Tuple = None, ca: Optional[Dict] = None,
                 ra: Optional[Dict] = None) -> "VelocytoLoom":
        """Build the object from by-cell compressed layers, ``(data, indices, indptr)`` per layer with ``indptr`` over
        CELLS and gene ids as indices -- the arrays a 10x / AnnData / sparse-loom HDF5 file stores, read with any HDF5
        reader (h5py is not part of this package's requirements).  Nothing dense is built on the host."""
        def mat(t):
            data, indices, indptr = t
            return sparse.csc_matrix((np.asarray(data, dtype=np.float64), np.asarray(indices), np.asarray(indptr)),
                                     shape=(n_genes, len(indptr) - 1))
        return cls(S=mat(S), U=mat(U), A=None if A is None else mat(A), ca=ca, ra=ra)

    # ------------------------------------------------------------------ device residency helpers
    def _dev(self, name: str, residual: bool = False):
        """Cell-major device copy of matrix attribute ``name`` (uploaded on first use).

        ``residual=True`` (used for the expression matrix of ``estimate_transition_prob``) re-uploads a float64
        host array together with its fp32 residuals so that fp32 ties keep the sign the reference sees."""
        from . import device as dev
        devs = self.__dict__.setdefault("_devs", {})
        host = self.__dict__.setdefault("_host", {})
        need_split = (residual and name in host and not sparse.issparse(host[name])
                      and getattr(host[name], "dtype", None) == np.float64
                      and (name not in devs or devs[name].lo is None)
                      and name not in self.__dict__.setdefault("_fp32_exact", set()))
        if name not in devs or need_split:
            if name not in host:
                raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")
            if sparse.issparse(host[name]):
                # sparse host layer: 8 bytes per NON-ZERO over PCIe, densified on the device (fp32, cell-major)
                devs[name] = dev.CsrCounts.from_scipy(host[name]).to_cellmajor()
                self.__dict__["_fp32_exact"].discard(name)
                return devs[name]
            devs[name] = dev.CellMajor.from_gene_major(host[name], residual=residual)
            if residual and devs[name].lo is None:
                self.__dict__["_fp32_exact"].add(name)
        return devs[name]

    def _set_dev(self, name: str, cm) -> None:
        self.__dict__.setdefault("_devs", {})[name] = cm
        self.__dict__.setdefault("_host", {}).pop(name, None)
        self.__dict__.setdefault("_fp32_exact", set()).discard(name)

    # ------------------------------------------------------------------ normalize (analysis.py:535-676)
    def _size_log_normalize(self, src: str, size: bool, log: bool, pcount: float, cell_size, target_size,
                            guard: bool, dst_sz: str, dst_norm: str):
        """Shared body of ``_normalize_S/_U/_Sx/_Ux``: per-cell totals (``velo_dev_cell_sums``) unless a size vector
        is given, ``norm_factor = avg_size / cell_size`` in float64 on the host (C values), then one device pass writing
        ``X_sz`` and, with ``log``, ``X_norm = log2(X_sz + pcount)``.  Returns ``(cell_size, avg_size, norm_factor)``."""
        import torch
        from . import device as dev
        X = self._dev(src)
        if size:
            if cell_size is None:
                cell_size = dev.cell_sums(X).cpu().numpy()
            cell_size = np.asarray(cell_size)
            avg_size = cell_size.mean() if target_size is None else target_size
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with np.errstate(divide="ignore", invalid="ignore"):
                    norm_factor = avg_size / cell_size
            fac = torch.from_numpy(np.array(np.broadcast_to(norm_factor, (X.C,)), dtype=np.float64)).to(X.t.device)
        else:
            avg_size, norm_factor, fac = None, 1, None
        sz, nm = dev.size_normalize(X, fac, pcount, want_sz=True, want_norm=bool(log), nonfinite_to_zero=guard)
        self._set_dev(dst_sz, sz)
        if log:
            self._set_dev(dst_norm, nm)
        return cell_size, avg_size, norm_factor

    def _normalize_S(self, size: bool = True, log: bool = True, pcount: float = 1, relative_size=None,
                     target_size=None) -> None:
        """analysis.py:535-552: creates ``cell_size``, ``avg_size``, ``norm_factor``, ``S_sz`` and ``S_norm``."""
        cs, avg, nf = self._size_log_normalize("S", size, log, pcount,
                                               relative_size if type(relative_size) is np.ndarray else None,
                                               target_size, False, "S_sz", "S_norm")
        if size:
            self.cell_size, self.avg_size = cs, avg
        self.norm_factor = nf

    def _normalize_U(self, size: bool = True, log: bool = True, pcount: float = 1, use_S_size: bool = False,
                     relative_size=None, target_size=None) -> None:
        """analysis.py:554-584 (non-finite ``U_sz`` -> 0, :581)."""
        given = None
        if size:
            if use_S_size:
                given = self.cell_size if hasattr(self, "cell_size") else self._host_or_dev_sums("S")
            elif type(relative_size) is np.ndarray:
                given = relative_size
        cs, avg, nf = self._size_log_normalize("U", size, log, pcount, given, target_size, True, "U_sz", "U_norm")
        if size:
            self.Ucell_size, self.Uavg_size = cs, avg
        self.Unorm_factor = nf

    def _normalize_Sx(self, size: bool = True, log: bool = True, pcount: float = 1, relative_size=None,
                      target_size=None) -> None:
        """analysis.py:586-601 (``if relative_size:`` -- truthiness, as in the reference)."""
        cs, avg, nf = self._size_log_normalize("Sx", size, log, pcount, relative_size if relative_size else None,
                                               target_size, False, "Sx_sz", "Sx_norm")
        if size:
            self.xcell_size, self.xavg_size = cs, avg
        self.xnorm_factor = nf

    def _normalize_Ux(self, size: bool = True, log: bool = True, pcount: float = 1, use_Sx_size: bool = False,
                      relative_size=None, target_size=None) -> None:
        """analysis.py:603-633 (the ``hasattr(self, "cell_size")`` test guarding ``xcell_size`` is the reference's)."""
        given = None
        if size:
            if use_Sx_size:
                given = self.xcell_size if hasattr(self, "cell_size") else self._host_or_dev_sums("Sx")
            elif type(relative_size) is np.ndarray:
                given = relative_size
        cs, avg, nf = self._size_log_normalize("Ux", size, log, pcount, given, target_size, True, "Ux_sz", "Ux_norm")
        if size:
            self.xUcell_size, self.xUavg_size = cs, avg
        self.xUnorm_factor = nf

    def _host_or_dev_sums(self, name: str) -> np.ndarray:
        from . import device as dev
        return dev.cell_sums(self._dev(name)).cpu().numpy()

    def normalize(self, which: str = "both", size: bool = True, log: bool = True, pcount: float = 1,
                  relative_size: np.ndarray = None, use_S_size_for_U: bool = False,
                  target_size: Tuple[float, float] = (None, None)) -> None:
        """Normalization interface (analysis.py:635-676): creates ``S_sz``/``S_norm``, ``U_sz``/``U_norm`` or the
        ``Sx``/``Ux`` twins; the matrices stay on the device."""
        plan = {"both": ("S", "U"), "S": ("S",), "U": ("U",), "imputed": ("Sx", "Ux"), "Sx": ("Sx",), "Ux": ("Ux",)}
        for name in plan.get(which, ()):                      # an unknown `which` is a silent no-op in the reference too
            common = dict(size=size, log=log, pcount=pcount, relative_size=relative_size)
            if name in ("S", "Sx"):                           # spliced: first entry of target_size
                getattr(self, "_normalize_" + name)(target_size=target_size[0], **common)
            else:                                             # unspliced: may borrow the spliced cell sizes
                borrow = {"use_S_size" if name == "U" else "use_Sx_size": use_S_size_for_U}
                getattr(self, "_normalize_" + name)(target_size=target_size[1], **borrow, **common)

    # ------------------------------------------------------------------ perform_PCA (analysis.py:678-702)
    def perform_PCA(self, which: str = "S_norm", n_components: int = None, div_by_std: bool = False) -> None:
        """PCA with cells as samples; creates ``pca`` (the fitted attributes scikit-learn's object would carry) and
        ``pcs`` ``(cells, npcs)``.  Exact (eigendecomposition of the fp64 second-moment matrix accumulated on the
        device) where scikit-learn's ``svd_solver="auto"`` switches to a randomised solver for large inputs."""
        from . import device as dev
        pcs, self.pca = dev.pca(self._dev(which), n_components, div_by_std)
        self.pcs = pcs.cpu().numpy()

    # ------------------------------------------------------------------ knn_imputation (analysis.py:933-1023)
    def knn_imputation(self, k: int = None, pca_space: float = True, metric: str = "euclidean", diag: float = 1,
                       n_pca_dims: int = None, maximum: bool = False, size_norm: bool = True,
                       balanced: bool = False, b_sight: int = None, b_maxl: int = None,
                       group_constraint: Union[str, np.ndarray] = None, n_jobs: int = 8) -> None:
        """k-nn smoothing of the data matrix; creates ``knn``, ``knn_smoothing_w``, ``Sx``, ``Ux``, ``Sx_sz``, ``Ux_sz``."""
        N = self.S.shape[1]
        if k is None:
            k = int(N * 0.025)                                                          # analysis.py:983-984
        if b_sight is None and balanced:
            b_sight = np.maximum(int(k * 8), N - 1)                                     # analysis.py:985-988 (sic: maximum)
        if b_maxl is None and balanced:
            b_maxl = np.maximum(int(k * 4), N - 1)
        space = self.pcs[:, :n_pca_dims] if pca_space else self.S_norm.T                 # analysis.py:989-992
        if balanced:                                                                     # analysis.py:993-1001
            from .neighbors import BalancedKNN
            constraint = None
            if group_constraint is not None:
                constraint = (np.array(self.cluster_ix) if isinstance(group_constraint, str) and group_constraint == "clusters"
                              else np.asarray(group_constraint))
            bknn = BalancedKNN(k=k, sight_k=b_sight, maxl=b_maxl, metric=metric, constraint=constraint, mode="distance",
                               n_jobs=n_jobs)
            bknn.fit(space)
            self.knn = bknn.kneighbors_graph(mode="distance")
        else:
            if group_constraint is not None:
                raise ValueError("group_constraint is currently supported only if the argument balanced is set to True")
            self.knn = knn_distance_matrix(space, metric=metric, k=k, mode="distance", n_jobs=n_jobs)
        # (knn > 0).astype(float); setdiag(diag); connectivity_to_weights  (analysis.py:1006-1010, neighbors.py:385-390):
        # rows / rowsum on the CSR arrays -- SciPy's `multiply` by a dense column returns COO and, with the comparison
        # and setdiag, costs seconds at 5e6 edges; the values are the same
        self.knn_smoothing_w = smoothing_weights_from_knn(self.knn, diag)
        self._smooth(maximum, size_norm)

    def knn_imputation_precomputed(self, knn_smoothing_w, maximum: bool = False) -> None:
        """Smoothing with externally computed weights (analysis.py:1025-1053)."""
        self.knn_smoothing_w = knn_smoothing_w
        self._smooth(maximum, True)

    def _smooth(self, maximum: bool, size_norm: bool) -> None:
        import torch
        from . import device as dev
        dv = dev.require_cuda()
        ip, ix, wt = self._weights_to_device(self.knn_smoothing_w, dv)
        # the reference's `assert np.allclose(w_.sum(0), 1)` (neighbors.py:422), evaluated on the device
        rows = torch.repeat_interleave(torch.arange(ip.numel() - 1, device=dv), ip[1:] - ip[:-1])
        rsum = torch.zeros(ip.numel() - 1, dtype=torch.float64, device=dv).index_add_(0, rows, wt.double())
        assert bool(((rsum - 1).abs() <= 1e-5).all()), "weight matrix need to sum to one over the columns"
        src_s, src_u = ("S_sz", "U_sz") if size_norm else ("S", "U")                    # analysis.py:1011-1016
        host = self.__dict__.setdefault("_host", {})

        def smooth(name, own_max):
            # sparse counts (an extension: the reference is dense-only) stay CSR on the device and only the smoothed
            # matrix is dense -- what makes the 500k-cell configuration fit (velo_dev_knn_smooth_csr)
            if name in host and sparse.issparse(host[name]):
                return dev.knn_smooth_csr(ip, ix, wt, sparse.csr_matrix(host[name].T), maximum=own_max)
            return dev.knn_smooth(ip, ix, wt, self._dev(name), own_max)

        # maximum: np.maximum(self.S_sz, self.Sx) -- against the SIZE-NORMALISED counts whatever was smoothed
        # (analysis.py:1017-1019); fused into the kernel when that is also its input
        Sx, Ux = smooth(src_s, maximum and size_norm), smooth(src_u, maximum and size_norm)
        if maximum and not size_norm:
            torch.maximum(Sx.t, self._dev("S_sz").t, out=Sx.t)
            torch.maximum(Ux.t, self._dev("U_sz").t, out=Ux.t)
        self._set_dev("Sx", Sx)
        self._set_dev("Ux", Ux)
        # "a differently named variable for backwards compatibility" (analysis.py:1022-1023): copies in the reference;
        # here the SAME device matrix under both names -- no kernel ever writes a named matrix in place (every stage
        # allocates its output), and a host-side assignment replaces only the name it targets
        self._set_dev("Sx_sz", Sx)
        self._set_dev("Ux_sz", Ux)

    @staticmethod
    def _weights_to_device(w, dv):
        """CSR-by-rows arrays of the smoothing weights as device tensors ``(indptr int64, indices int32, data fp32)``.
        A SciPy CSR matrix with canonical (duplicate-free) rows -- what ``connectivity_to_weights`` returns -- goes
        over as it is; anything else is canonicalised by SciPy first."""
        import torch
        if not isinstance(w, sparse.csr_matrix):
            w = sparse.csr_matrix(w)                   # COO (what SciPy's multiply returns) / dense: canonicalised by SciPy
            w.sum_duplicates()
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dv).to(dt)
        return to(w.indptr, torch.int64), to(w.indices, torch.int32), to(w.data, torch.float32)

    # ------------------------------------------------------------------ fit_gammas (analysis.py:1120-1260)
    def fit_gammas(self, steady_state_bool: np.ndarray = None, use_imputed_data: bool = True, use_size_norm: bool = True,
                   fit_offset: bool = True, fixperc_q: bool = False, weighted: bool = True,
                   weights: Union[str, np.ndarray] = "maxmin_diag", limit_gamma: bool = False,
                   maxmin_perc: List[float] = [2, 98], maxmin_weighted_pow: float = 15) -> None:
        """Fit gamma using spliced and unspliced data; creates ``gammas``, ``q`` (and ``R2`` for weighted fits)."""
        from . import device as dev
        if steady_state_bool is not None:                                                # analysis.py:1159-1162
            self.steady_state = np.asarray(steady_state_bool, dtype=bool)
        else:
            self.steady_state = np.ones(self.S.shape[1], dtype=bool)
        if use_imputed_data:
            nS, nU = ("Sx_sz", "Ux_sz") if use_size_norm else ("Sx", "Ux")               # analysis.py:1164-1177
        else:
            nS, nU = ("S_sz", "U_sz") if use_size_norm else ("S", "U")
        Sd, Ud = self._dev(nS), self._dev(nU)
        mask = None if self.steady_state.all() else self.steady_state
        Wd = None
        if weighted:
            Wd = self._fit_weights(weights, nS, nU, maxmin_perc, maxmin_weighted_pow)
        import torch
        mask_t = None if mask is None else torch.from_numpy(mask.astype(np.uint8))
        # dispatch of analysis.py:1221-1257: fit_offset > fixperc_q > plain
        use_fix = fixperc_q and not fit_offset
        if fit_offset or use_fix:
            mode = dev.FIT_SLOPE_WEIGHTED_OFFSET if weighted else dev.FIT_SLOPE_OFFSET
        else:
            mode = dev.FIT_SLOPE_WEIGHTED if weighted else dev.FIT_SLOPE
        if limit_gamma and not weighted:
            logging.warning("limit_gamma not implemented with this settings")            # analysis.py:1230,1242,1253
        use_lim = limit_gamma and weighted and not use_fix
        if mask is not None and (use_fix or use_lim):
            # the reference slices tmpS[:, steady_state] before every fit (analysis.py:1223-1256), so the percentile
            # constraints of fixperc_q / limit_gamma (estimation.py:199-204, 221-224) see the selected cells only:
            # compact the cell rows on the device and fit without a mask
            sel = torch.from_numpy(np.flatnonzero(mask)).to(Sd.t.device)
            Sd, Ud = dev.CellMajor(Sd.t.index_select(0, sel), Sd.G), dev.CellMajor(Ud.t.index_select(0, sel), Ud.G)
            if Wd is not None:
                Wd = dev.CellMajor(Wd.t.index_select(0, sel), Wd.G)
            mask_t = None
        qf, up = dev.fit_constraints(Sd, Ud, use_fix, use_lim)
        lo = 1e-8 if (mode == dev.FIT_SLOPE_WEIGHTED_OFFSET or use_lim) else 0.0
        g, q, r2, _ = dev.fit_gammas(mode, Sd, Ud, Wd, mask_t, lo, 20.0, want_r2=weighted and not use_fix,
                                     hi_per_gene=up, q_fixed=qf)
        fit_offset = fit_offset or use_fix
        self._gamma_dev, self._q_dev = torch.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0), q
        self.gammas = g.cpu().numpy()
        self.q = q.cpu().numpy() if fit_offset else np.zeros_like(self.gammas)          # analysis.py:1250,1257
        if r2 is not None:
            self.R2 = r2.cpu().numpy()
        self.gammas[~np.isfinite(self.gammas)] = 0                                      # analysis.py:1260

    def _fit_weights(self, weights, nS: str, nU: str, maxmin_perc, maxmin_weighted_pow):
        """Weight matrix of the least-squares fit (analysis.py:1179-1219) as a device matrix.

        The per-gene percentiles (np.percentile, linear interpolation) are order statistics over the cell
        axis: they are found by a radix-select kernel on the device (``velo_dev_fit_weights_ex``).  Only
        user-supplied weight arrays come from the host."""
        from . import device as dev
        if isinstance(weights, np.ndarray):
            return dev.CellMajor.from_gene_major(weights)
        if weights not in dev.WEIGHT_KINDS:
            raise ValueError(f"unknown weights={weights!r}")
        diag = weights in ("maxmin_diag", "maxmin_double")
        return dev.fit_weights(weights, self._dev(nS), self._dev(nU), self._dev("Sx") if diag else None,
                               self._dev("Ux") if diag else None, maxmin_perc, maxmin_weighted_pow)

    # ------------------------------------------------------------------ predict_U .. extrapolate (analysis.py:1321-1439)
    def _gamma_q(self, which_gamma: str, which_offset: Optional[str]):
        import torch
        from . import device as dev
        dv = dev.require_cuda()
        g = torch.from_numpy(np.asarray(getattr(self, which_gamma), dtype=np.float32)).to(dv)
        q = None if which_offset is None else torch.from_numpy(np.asarray(getattr(self, which_offset), dtype=np.float32)).to(dv)
        return g, q

    def predict_U(self, which_gamma: str = "gammas", which_S: str = "Sx_sz", which_offset: str = "q") -> None:
        """``Upred = gamma * S (+ q)`` (analysis.py:1321-1346)."""
        from . import device as dev
        self.which_S_for_pred = which_S
        if which_offset is None and (hasattr(self, "q_W") or hasattr(self, "q")):
            logging.warning("Predicting U without intercept but intercept was previously fit! Set which_offset='q' or 'q_W' ")
        self._pred_gamma, self._pred_q = self._gamma_q(which_gamma, which_offset)
        Sd = self._dev(which_S)
        out = dev.velocity_chain(Sd, Sd, self._pred_gamma, self._pred_q, want=("Upred",))   # U is not read for Upred
        self._set_dev("Upred", out["Upred"])

    def calculate_velocity(self, kind: str = "residual", eps: float = None) -> None:
        """``velocity = U_measured - U_predicted`` (analysis.py:1348-1379)."""
        from . import device as dev
        if kind != "residual":
            raise NotImplementedError(f"Velocity calculation kind={kind} is not implemented")
        if self.which_S_for_pred not in ("Sx_sz", "Sx"):
            return                                                                       # analysis.py:1373 (no-op, sic)
        nS = self.which_S_for_pred
        nU = "Ux_sz" if nS == "Sx_sz" else "Ux"
        out = dev.velocity_chain(self._dev(nS), self._dev(nU), self._pred_gamma, self._pred_q, eps=eps,
                                 want=("velocity",))
        self._velocity_eps = eps
        self._set_dev("velocity", out["velocity"])

    def calculate_shift(self, assumption: str = "constant_velocity", delta_t: float = 1) -> None:
        """``delta_S`` under Model I / Model II (analysis.py:1381-1408)."""
        from . import device as dev
        if assumption not in ("constant_velocity", "constant_unspliced"):
            raise NotImplementedError(f"Assumption {assumption} is not implemented")
        if assumption == "constant_unspliced":
            g, q = self._gamma_q("gammas", "q")                                          # analysis.py:1403-1406 uses self.gammas/self.q
            out = dev.velocity_chain(self._dev("Sx_sz"), self._dev("Ux_sz"), g, q, assumption=assumption,
                                     dt_shift=delta_t, want=("delta_S",))
        else:
            # delta_S = delta_t * self.velocity (analysis.py:1399): whatever `velocity` currently is -- the device
            # matrix calculate_velocity left, or an array the user assigned / masked since (re-uploaded by _dev)
            out = {"delta_S": dev.delta_transform(self._dev("velocity"), delta_t, "linear", 0.0)}
        self._shift = (assumption, float(delta_t))
        self._set_dev("delta_S", out["delta_S"])

    def extrapolate_cell_at_t(self, delta_t: float = 1, clip: bool = True) -> None:
        """``Sx_sz_t = clip(Sx_sz + delta_t * delta_S, 0)`` (analysis.py:1410-1439)."""
        from . import device as dev
        if self.which_S_for_pred not in ("Sx_sz", "Sx"):
            return
        nS = self.which_S_for_pred
        out = dev.extrapolate(self._dev(nS), self._dev("delta_S"), delta_t, clip)
        if clip:
            self.used_delta_t = delta_t                                                  # analysis.py:1432 (set only when clip, sic)
        self._set_dev("Sx_sz_t" if nS == "Sx_sz" else "Sx_t", out)

    # ------------------------------------------------------------------ estimate_transition_prob (analysis.py:1452-1668)
    def estimate_transition_prob(self, hidim: str = "Sx_sz", embed: str = "ts", transform: str = "sqrt",
                                 ndims: int = None, n_sight: int = None, psc: float = None,
                                 knn_random: bool = True, sampled_fraction: float = 0.3,
                                 sampling_probs: Tuple[float, float] = (0.5, 0.1), max_dist_embed: float = None,
                                 n_jobs: int = 4, threads: int = None, calculate_randomized: bool = True,
                                 random_seed: int = 15071990, **kwargs) -> None:
        """Correlation of the velocity with the displacement towards each embedding neighbour.

        Creates ``corrcoef`` (+ ``corrcoef_random``), ``embedding``, ``embedding_knn``, ``sampling_ixs``,
        ``corr_calc``, ``delta_S_rndm`` as the reference does.  ``corrcoef`` is the dense ``(cells, cells)``
        float64 matrix when it fits comfortably (cells <= ``dense_limit``, default 20000) and is always
        available in compact form as ``corrcoef_compact`` / ``neigh_ixs`` (cells x m)."""
        import torch
        from . import device as dev
        dense_limit = kwargs.pop("dense_limit", 20000)
        # "reference" (default): the reference's NumPy / numba random streams on the host -- bit-equal sampling_ixs and
        # delta_S_rndm, but a Python loop of np.random.choice per cell and a numba shuffle per gene (tens of seconds at
        # 100k cells).  "device": the same distributions drawn by velo_dev_sample_neighbors / velo_dev_permute_rows_nsign.
        random_backend = kwargs.pop("random_backend", "reference")
        if random_backend not in ("reference", "device"):
            raise ValueError("random_backend must be 'reference' or 'device'")
        for stale in ("_corrcoef_random_dev", "corrcoef_random_compact", "_transition_prob_random_dev",
                      "transition_prob_random_compact"):
            self.__dict__.pop(stale, None)
        for stale in ("corrcoef_random", "transition_prob_random"):
            if stale in self.__dict__.get("_dense", {}) or stale in self.__dict__.get("_dense_src", {}):
                delattr(self, stale)
        if calculate_randomized and random_backend == "reference":
            _numba_seed(random_seed)                                                     # analysis.py:1501 (only the randomised
        self.which_hidim = hidim                                                         # control consumes numba's stream, :2413-2420)
        if "n_neighbors" in kwargs:
            n_neighbors = kwargs.pop("n_neighbors")
            if len(kwargs) > 0:
                logging.warning(f"keyword arguments were passed but could not be interpreted {kwargs}")
        else:
            n_neighbors = None
        if n_sight is None and n_neighbors is None:
            n_neighbors = int(self.S.shape[1] / 5)
        if (n_sight is not None) and (n_neighbors is not None) and n_neighbors != n_sight:
            raise ValueError("n_sight and n_neighbors are different names for the same parameter, they cannot be set differently")
        if n_sight is not None and n_neighbors is None:
            n_neighbors = n_sight
        if psc is None:                                                                  # analysis.py:1520-1526
            psc = 1. if transform in ("log", "logratio") else (1e-10 if transform == "sqrt" else 0)
        if transform not in ("log", "sqrt", "linear", "logratio"):
            raise NotImplementedError(f"transform={transform} is not a valid parameter")
        use_pcs = "pcs" in hidim                                                         # sic, analysis.py:1531
        if use_pcs and calculate_randomized:
            raise ValueError("calculate_randomized is not available with hidim='pcs' (the reference never builds the "
                             "randomised displacement in that branch, analysis.py:1531-1533)")
        if ndims is not None and not use_pcs:
            raise ValueError(f"ndims was set to {ndims} but hidim != 'pcs'. Set ndims = None for hidim='{hidim}'")
        if knn_random:
            np.random.seed(random_seed)                                                  # analysis.py:1529
        self.corr_calc = "knn_random" if knn_random else "full"
        if use_pcs:
            # principal-component space: rows = components, columns = cells (analysis.py:1532-1533)
            hi = np.array(getattr(self, hidim).T[:, :ndims], order="C")
            hi_t = np.array(getattr(self, hidim + "_t").T[:, :ndims], order="C")
            e_cm = dev.CellMajor.from_gene_major(hi, residual=True)
            dS = dev.CellMajor.from_gene_major(hi_t - hi)
            used_dt = 1.0                                                                # delta = hi_dim_t - hi_dim directly
        else:
            e_cm = self._dev(hidim, residual=True)
            dS = self._dev("delta_S")
            used_dt = float(self.used_delta_t)
        C, G = e_cm.C, e_cm.G
        tname = {"log": "log10", "sqrt": "sqrt", "linear": "linear", "logratio": "linear"}[transform]
        hi_cm = e_cm
        if transform == "logratio":
            # correlate log2(hi_dim + psc) with log2(|hi_dim_t| + psc) - log2(hi_dim + psc), linear kernel (analysis.py:1582-1590)
            e_cm = dev.logratio(hi_cm, None, 0.0, psc, 0)

        def transformed(delta_cm):
            # d = f(hi_dim_t - hi_dim), hi_dim_t = hi_dim + used_delta_t * delta_S   (analysis.py:1538, 1577/1594/1597)
            if transform == "logratio":
                return dev.logratio(hi_cm, delta_cm, used_dt, psc, 1)
            return dev.delta_transform(delta_cm, used_dt, tname, psc)

        d_cm = transformed(dS)
        d_rnd = None
        if calculate_randomized:                                                         # analysis.py:1539-1542
            if random_backend == "device":
                self._set_dev("delta_S_rndm", dev.permute_rows_nsign(dS, random_seed))
            else:
                rnd = np.copy(self.delta_S)
                _permute_rows_nsign(rnd)
                self.delta_S_rndm = rnd
            d_rnd = transformed(self._dev("delta_S_rndm"))
        embedding = getattr(self, embed)
        self.embedding = embedding
        knn_idx_dev = None
        if knn_random and random_backend == "device" and n_neighbors + 1 <= dev.KNN_MAX_K:
            knn_idx_dev, _ = dev.knn(np.ascontiguousarray(embedding, dtype=np.float64), n_neighbors + 1, False)
        elif n_neighbors + 1 <= dev.KNN_MAX_K:                                           # analysis.py:1547-1549, on the device
            self.embedding_knn = knn_graph_device(np.asarray(embedding), n_neighbors + 1, "connectivity")
        else:
            from sklearn.neighbors import NearestNeighbors       # only beyond the device kernel's k limit (1.4 s of import)
            nn = NearestNeighbors(n_neighbors=n_neighbors + 1, n_jobs=n_jobs)
            nn.fit(embedding)
            self.embedding_knn = nn.kneighbors_graph(mode="connectivity")
        if knn_random:
            p = np.linspace(sampling_probs[0], sampling_probs[1], n_neighbors + 1)
            p = p / p.sum()
            size = int(sampled_fraction * (n_neighbors + 1))
            if random_backend == "device":
                if knn_idx_dev is None:
                    knn_idx_dev = torch.from_numpy(np.ascontiguousarray(
                        self.embedding_knn.indices.reshape((-1, n_neighbors + 1)), dtype=np.int32)).to(dev.require_cuda())
                ix_dev, samp_dev = dev.sample_neighbors(knn_idx_dev.contiguous(), p, size, random_seed)
                self.sampling_ixs = samp_dev.cpu().numpy().astype(np.intp)
                neigh_ixs = ix_dev.cpu().numpy().astype(np.intp)
            else:
                neigh_ixs = self.embedding_knn.indices.reshape((-1, n_neighbors + 1))
                sampling_ixs = _sample_neighbors_numpy_stream(random_seed, neigh_ixs.shape[0], neigh_ixs.shape[1], p, size)
                self.sampling_ixs = sampling_ixs
                neigh_ixs = neigh_ixs[np.arange(neigh_ixs.shape[0])[:, None], sampling_ixs]
            nonzero = neigh_ixs.shape[0] * neigh_ixs.shape[1]
            self.embedding_knn = sparse.csr_matrix((np.ones(nonzero), neigh_ixs.ravel(),
                                                    np.arange(0, nonzero + 1, neigh_ixs.shape[1])),
                                                   shape=(neigh_ixs.shape[0], neigh_ixs.shape[0]))
            ix = ix_dev if random_backend == "device" else dev.indices_to_device(neigh_ixs, C)
            self.neigh_ixs = neigh_ixs
        else:
            ix = None
            self.neigh_ixs = None
        self._ix_dev = ix
        corr = dev.coldeltacor(e_cm, d_cm, ix, tname, psc)
        self._finish_corr("corrcoef", corr, ix, C, dense_limit, warn=knn_random)
        if calculate_randomized:
            corr_r = dev.coldeltacor(e_cm, d_rnd, ix, tname, psc)
            self._finish_corr("corrcoef_random", corr_r, ix, C, dense_limit, warn=knn_random)

    def _finish_corr(self, name: str, corr, ix, C: int, dense_limit: int, warn: bool) -> None:
        """diag -> 0, NaN -> 1 (analysis.py:1604-1612, 1666-1668) on the compact result; dense adapter for small C."""
        from . import device as dev
        # np.fill_diagonal(corrcoef, 0); knn_random mode also maps NaN -> 1 with a warning
        if dev.patch_corr(corr, ix, 0, patch_nan=warn):
            logging.warning(f"Nans encountered in {name} and corrected to 1s. If not identical cells were present "
                            "it is probably a small isolated cluster converging after imputation.")
        setattr(self, "_" + name + "_dev", corr)
        setattr(self, name + "_compact", corr.cpu().numpy())
        self._set_dense_source(name, corr, ix, C, dense_limit)

    def _set_dense_source(self, name: str, compact, ix, C: int, dense_limit: int) -> None:
        """Register the compact device result behind the lazily built dense attribute ``name`` (see _LazyDense)."""
        self.__dict__.setdefault("_dense", {}).pop(name, None)
        self.__dict__.setdefault("_dense_src", {})[name] = (compact, ix, C, dense_limit)

    # ------------------------------------------------------------------ calculate_embedding_shift (analysis.py:1670-1733)
    def calculate_embedding_shift(self, sigma_corr: float = 0.05, expression_scaling: bool = True,
                                  scaling_penalty: float = 1., dense_limit: int = 20000) -> None:
        """Transition probabilities (exponential kernel on the correlations, row-normalised over the
        embedding neighbourhood) and their projection on the embedding.

        ``transition_prob_compact`` (cells x m, aligned with ``neigh_ixs``) always; the reference's dense
        ``transition_prob`` / ``delta_embedding`` when cells <= ``dense_limit``."""
        import torch
        from . import device as dev
        if self.corr_calc not in ("full", "knn_random"):
            raise NotImplementedError(f"Weird value self.corr_calc={self.corr_calc}")
        C = self.embedding.shape[0]
        ix = self._ix_dev
        if ix is None:
            # full mode: the mask is the (n_neighbors + 1)-nn graph of the embedding (analysis.py:1634, 1697)
            knn = self.embedding_knn.tocsr()
            m = int(knn.indptr[1] - knn.indptr[0])
            nb = knn.indices.reshape(C, m)
            ix_m = dev.indices_to_device(nb, C)
            self.neigh_ixs = nb
        else:
            ix_m = ix
        names = [("corrcoef", "transition_prob")]
        if hasattr(self, "_corrcoef_random_dev"):
            names.append(("corrcoef_random", "transition_prob_random"))
        for cname, pname in names:
            corr = getattr(self, "_" + cname + "_dev")
            poisoned = None
            if ix is None:
                # "full" branch: exp(corrcoef / sigma) * embedding_knn.A over the DENSE matrix (analysis.py:1697): a NaN
                # anywhere in a row (only the diagonal was zeroed, :1666-1668) makes exp(NaN) * 0 = NaN and with it
                # the whole row of transition probabilities, neighbours or not
                poisoned = torch.isnan(corr).any(dim=1)
                corr = torch.gather(corr, 1, ix_m.to(torch.int64))
            # NaN -> 1 belongs to the knn_random branch only (analysis.py:1604-1612)
            tp = dev.transition_prob(corr.contiguous(), ix_m, sigma_corr, patch_nan=ix is not None)
            if poisoned is not None and bool(poisoned.any()):
                tp[poisoned] = float("nan")
            setattr(self, "_" + pname + "_dev", tp)
            setattr(self, pname + "_compact", tp.cpu().numpy())
            self._set_dense_source(pname, tp, ix_m, C, dense_limit)
        # delta_embedding = sum_j (P_ij - 1/k) * unit(emb_j - emb_i), neighbours only (analysis.py:1704-1712)
        for cname, pname in names:
            tp = getattr(self, "_" + pname + "_dev")
            de = dev.embedding_shift(tp, ix_m, self.embedding).cpu().numpy()
            rnd = pname != "transition_prob"
            if expression_scaling:                                                       # analysis.py:1714-1719, 1726-1731
                sc = dev.expression_scaling(tp, ix_m, self._dev(self.which_hidim),
                                            self._dev("delta_S_rndm" if rnd else "delta_S"), scaling_penalty).cpu().numpy()
                setattr(self, "scaling_rndm" if rnd else "scaling", sc)
                de = de * sc[:, None]
            setattr(self, "delta_embedding_random" if rnd else "delta_embedding", de)


    # ------------------------------------------------------------------ calculate_grid_arrows (analysis.py:1735-1816)
    def calculate_grid_arrows(self, embed: str = "embedding", smooth: float = 0.5, steps: Tuple = (40, 40),
                              n_neighbors: int = 100, n_jobs: int = 4) -> None:
        """Velocity field on a regular grid: gaussian-kernel average of ``delta_<embed>`` over the ``n_neighbors``
        cells nearest to every grid point.  Creates ``flow_embedding``, ``flow_grid``, ``flow``, ``flow_norm``,
        ``flow_norm_magnitude``, ``total_p_mass`` (and the ``_rndm`` twins when a randomised control exists)."""
        from . import device as dev
        embedding = np.asarray(getattr(self, embed))
        if not hasattr(self, f"delta_{embed}"):
            raise KeyError("This embedding does not have a delta_*")
        delta_embedding = getattr(self, f"delta_{embed}")
        has_random = "_corrcoef_random_dev" in self.__dict__ or "corrcoef_random" in self.__dict__.get("_dense", {})
        grs = []
        for dim_i in range(embedding.shape[1]):                                          # analysis.py:1776-1782
            m, M = np.min(embedding[:, dim_i]), np.max(embedding[:, dim_i])
            m = m - 0.025 * np.abs(M - m)
            M = M + 0.025 * np.abs(M - m)                                                # (uses the widened m, as the reference does)
            grs.append(np.linspace(m, M, steps[dim_i]))
        meshes_tuple = np.meshgrid(*grs)
        gridpoints_coordinates = np.vstack([i.flat for i in meshes_tuple]).T
        neighs, dists = dev.knn_query(embedding, gridpoints_coordinates, n_neighbors)    # analysis.py:1788-1790
        std = np.mean([(g[1] - g[0]) for g in grs])
        mass, UZ = dev.grid_flow(neighs, dists, delta_embedding, smooth * std)           # analysis.py:1792-1797
        self.total_p_mass = mass.cpu().numpy()
        UZ = UZ.cpu().numpy()
        magnitude = np.linalg.norm(UZ, axis=1)
        self.flow_embedding = embedding
        self.flow_grid = gridpoints_coordinates
        self.flow = UZ
        self.flow_norm = UZ / np.percentile(magnitude, 99.5)
        self.flow_norm_magnitude = np.linalg.norm(self.flow_norm, axis=1)
        if has_random:                                                                   # analysis.py:1809-1816
            _, UZ_rndm = dev.grid_flow(neighs, dists, getattr(self, f"delta_{embed}_random"), smooth * std)
            UZ_rndm = UZ_rndm.cpu().numpy()
            magnitude_rndm = np.linalg.norm(UZ, axis=1)                                  # (of UZ, as in the reference)
            self.flow_rndm = UZ_rndm
            self.flow_norm_rndm = UZ_rndm / np.percentile(magnitude_rndm, 99.5)
            self.flow_norm_magnitude_rndm = np.linalg.norm(self.flow_norm_rndm, axis=1)


for _n in _MATRIX_ATTRS:
    setattr(VelocytoLoom, _n, _DeviceBacked(_n))
for _n in _DENSE_ATTRS:
    setattr(VelocytoLoom, _n, _LazyDense(_n))


def _sample_neighbors_numpy_stream(seed: int, n_cells: int, W: int, p: np.ndarray, size: int) -> np.ndarray:
    """``np.random.seed(seed); np.stack([np.random.choice(W, size, replace=False, p=p) for _ in range(n_cells)])``
    (analysis.py:1529, 1561-1564) -- the same MT19937 stream consumed the same way, bit for bit, by the C++ restatement
    in libvelo_b200 (csrc/host_sampler.cpp) instead of a Python loop of ~1 ms per cell.  NumPy's global generator is
    left in the state the reference's loop would have left it in."""
    import ctypes
    if not 0 <= int(seed) < 2 ** 32:
        raise ValueError("Seed must be between 0 and 2**32 - 1")
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty((n_cells, size), dtype=np.int64)
    key, pos = np.empty(624, dtype=np.uint32), ctypes.c_int(0)
    _cabi.call("velo_host_sample_neighbors_numpy", int(seed), n_cells, W, p.ctypes.data, size, out.ctypes.data,
               key.ctypes.data, ctypes.addressof(pos))
    np.random.set_state(("MT19937", key, pos.value, 0, 0.0))
    return out.astype(np.intp, copy=False)


# --------------------------------------------------------------------------- host RNG helpers (analysis.py:2407-2420)
def _numba_seed(value: int) -> None:
    _jit_helpers()[0](value)


def _permute_rows_nsign(A: np.ndarray) -> None:
    _jit_helpers()[1](A)


_JIT = None


def _jit_helpers():
    """numba-compiled twins of ``numba_random_seed`` / ``permute_rows_nsign``: the randomised control must
    consume numba's MT19937 stream exactly as the reference does to be reproducible against it."""
    global _JIT
    if _JIT is None:
        from numba import jit

        @jit(nopython=True, cache=True)
        def seed(value):
            np.random.seed(value)

        @jit(nopython=True, cache=True)
        def permute(A):
            plmi = np.array([+1, -1])
            for i in range(A.shape[0]):
                np.random.shuffle(A[i, :])
                A[i, :] = A[i, :] * np.random.choice(plmi, size=A.shape[1])

        _JIT = (seed, permute)
    return _JIT
