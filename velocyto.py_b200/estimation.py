"""Drop-in for ``velocyto/estimation.py`` -- same names, arguments and return values.

``colDeltaCor{,Log10,Sqrt}{,partial}`` (velocyto/estimation.py:11-170) marshal their host
arrays straight into the host tier of ``libvelo_b200.so`` (``velo_colDeltaCor*``, the symbols a
maintainer binds in place of ``velocyto.speedboosted._colDeltaCor*``, speedboosted.pyx:542-610).
The returned array is the reference's dense ``(cells, cells)`` float64 matrix, accumulated into
zeros.  ``threads`` is accepted and ignored (the work runs on the GPU).

Extensions that do not exist in the reference (opt-in keyword only):
``compact=True`` on the partial functions returns the ``(cells, m)`` float32 correlations aligned
with ``ixs`` instead of the dense matrix -- the only form that exists at 100k cells.
"""
from __future__ import annotations

import ctypes
from typing import Any, Optional, Tuple

import numpy as np

from . import _cabi


def _as_c_f64(name: str, a: np.ndarray, coerce: bool) -> np.ndarray:
    """The reference's typed-memoryview contract (``double[:, ::1]``, speedboosted.pyx:542-610)."""
    if coerce:
        a = np.require(a, requirements="C")            # estimation.py:59,113,167 (emat of the partial wrappers)
    if not isinstance(a, np.ndarray) or a.ndim != 2:
        raise ValueError(f"{name}: Buffer has wrong number of dimensions (expected 2)")
    if a.dtype != np.float64:
        raise ValueError(f"{name}: Buffer dtype mismatch, expected 'double' but got '{a.dtype}'")
    if not a.flags.c_contiguous:
        raise ValueError(f"{name}: ndarray is not C-contiguous")
    return a


def _num_threads(threads: Optional[int]) -> int:
    import multiprocessing                              # estimation.py:26-30 (value is unused on the GPU)
    return int(multiprocessing.cpu_count() / 2) if threads is None else max(threads, multiprocessing.cpu_count())


def _full(symbol: str, emat, dmat, threads, *extra) -> np.ndarray:
    emat = _as_c_f64("emat", emat, coerce=False)
    dmat = _as_c_f64("dmat", dmat, coerce=False)
    if emat.shape != dmat.shape:
        raise ValueError("emat and dmat must have the same shape")
    out = np.zeros((emat.shape[1], emat.shape[1]))      # estimation.py:31 -- the kernel accumulates into zeros
    _cabi.call(symbol, emat.ctypes.data, dmat.ctypes.data, out.ctypes.data, emat.shape[0], emat.shape[1],
               _num_threads(threads), *extra)
    return out


def _partial(symbol: str, transform: int, emat, dmat, ixs, threads, psc, compact: bool, pass_psc: bool):
    emat = _as_c_f64("emat", emat, coerce=True)
    dmat = _as_c_f64("dmat", dmat, coerce=False)
    if emat.shape != dmat.shape:
        raise ValueError("emat and dmat must have the same shape")
    ixs = np.require(ixs, requirements="C").astype(np.intp)      # estimation.py:60
    if ixs.ndim != 2 or ixs.shape[0] != emat.shape[1]:
        raise ValueError("ixs must be (ncells, nneighbours)")
    G, C = emat.shape
    if compact:
        out = np.empty((C, ixs.shape[1]), dtype=np.float32)
        _cabi.call("velo_colDeltaCorpartial_compact", transform, emat.ctypes.data, dmat.ctypes.data, 8,
                   ixs.ctypes.data, out.ctypes.data, G, C, ixs.shape[1], float(psc))
        return out
    out = np.zeros((C, C))
    args = [emat.ctypes.data, dmat.ctypes.data, out.ctypes.data, ixs.ctypes.data, G, C, ixs.shape[1],
            _num_threads(threads)]
    if pass_psc:
        args.append(float(psc))
    _cabi.call(symbol, *args)
    return out


def colDeltaCor(emat: np.ndarray, dmat: np.ndarray, threads: int = None) -> np.ndarray:
    """Correlation between the displacement ``d[:, c]`` and ``e[:, i] - e[:, c]`` for all cell pairs
    (velocyto/estimation.py:11-33 -> speedboosted.pyx:13-87)."""
    return _full("velo_colDeltaCor", emat, dmat, threads)


def colDeltaCorpartial(emat: np.ndarray, dmat: np.ndarray, ixs: np.ndarray, threads: int = None,
                       compact: bool = False) -> np.ndarray:
    """Same on the sampled neighbourhoods ``ixs`` (velocyto/estimation.py:36-62 -> speedboosted.pyx:263-346)."""
    return _partial("velo_colDeltaCorpartial", _cabi.LINEAR, emat, dmat, ixs, threads, 0.0, compact, False)


def colDeltaCorLog10(emat: np.ndarray, dmat: np.ndarray, threads: int = None, psc: float = 1.0) -> np.ndarray:
    """velocyto/estimation.py:65-87 -> speedboosted.pyx:178-257."""
    return _full("velo_colDeltaCorLog10", emat, dmat, threads, float(psc))


def colDeltaCorLog10partial(emat: np.ndarray, dmat: np.ndarray, ixs: np.ndarray, threads: int = None,
                            psc: float = 1.0, compact: bool = False) -> np.ndarray:
    """velocyto/estimation.py:90-116 -> speedboosted.pyx:449-538."""
    return _partial("velo_colDeltaCorLog10partial", _cabi.LOG10, emat, dmat, ixs, threads, psc, compact, True)


def colDeltaCorSqrt(emat: np.ndarray, dmat: np.ndarray, threads: int = None, psc: float = 0.0) -> np.ndarray:
    """velocyto/estimation.py:119-141 -> speedboosted.pyx:93-172."""
    return _full("velo_colDeltaCorSqrt", emat, dmat, threads, float(psc))


def colDeltaCorSqrtpartial(emat: np.ndarray, dmat: np.ndarray, ixs: np.ndarray, threads: int = None,
                           psc: float = 0.0, compact: bool = False) -> np.ndarray:
    """velocyto/estimation.py:144-170 -> speedboosted.pyx:352-443 (the default path of
    ``estimate_transition_prob``: ``transform="sqrt"``, ``knn_random=True``)."""
    return _partial("velo_colDeltaCorSqrtpartial", _cabi.SQRT, emat, dmat, ixs, threads, psc, compact, True)
