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


# --------------------------------------------------------------------------- gamma fits
def _fit(mode: int, Y, X, W=None, lo: float = 0.0, hi: float = 20.0, want_r2: bool = False,
         fixperc_q: bool = False, limit_gamma: bool = False):
    """Upload (genes, cells) host matrices, run the batched fit kernel, return float32 host vectors."""
    from . import device as dev
    Y, X = np.asarray(Y), np.asarray(X)
    if Y.shape != X.shape or Y.ndim != 2:
        raise ValueError("Y and X must be (genes, cells) arrays of the same shape")
    Xd, Yd = dev.CellMajor.from_gene_major(X), dev.CellMajor.from_gene_major(Y)
    Wd = None
    if W is not None:
        W = np.asarray(W)
        if W.shape != Y.shape:
            raise ValueError("W must have the shape of Y")
        Wd = dev.CellMajor.from_gene_major(W.astype(np.float64, copy=False))
    qf, up = dev.fit_constraints(Xd, Yd, fixperc_q, limit_gamma)
    gamma, offset, r2, _ = dev.fit_gammas(mode, Xd, Yd, Wd, None, lo, hi, want_r2=want_r2, hi_per_gene=up, q_fixed=qf)
    return (gamma.cpu().numpy(), offset.cpu().numpy(), None if r2 is None else r2.cpu().numpy())


def fit_slope(Y: np.ndarray, X: np.ndarray) -> np.ndarray:
    """Per-gene slope through the origin, constrained >= 0 (velocyto/estimation.py:267-279, ``_fit1_slope``
    :173-188 = ``nnls``): ``max(0, sum xy / sum xx)``; NaN where x == 0, 0 where y == 0.  float32."""
    return _fit(0, Y, X)[0]


def fit_slope_offset(Y: np.ndarray, X: np.ndarray, fixperc_q: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Per-gene OLS slope and intercept (velocyto/estimation.py:282-297; ``leastsq`` from (0, 0), :244-264)."""
    g, q, _ = _fit(1, Y, X, fixperc_q=fixperc_q)       # fixperc_q: q = median(y[x <= p1(x)]), slope bounded to (0, 20)
    return g, q


def fit_slope_weighted(Y: np.ndarray, X: np.ndarray, W: np.ndarray, return_R2: bool = False,
                       limit_gamma: bool = False, bounds: Tuple[float, float] = (0, 20)) -> Any:
    """Per-gene weighted slope through the origin, bounded to (0, 20) (velocyto/estimation.py:300-334;
    ``minimize_scalar(bounded)``, :191-209).  As in the reference, ``bounds`` is accepted but the per-gene
    solver always runs with its default (0, 20) (estimation.py:319 does not forward it)."""
    g, _, r2 = _fit(2, Y, X, W, 1e-8 if limit_gamma else 0.0, 20.0, want_r2=return_R2, limit_gamma=limit_gamma)
    return (g, r2) if return_R2 else g


def fit_slope_weighted_offset(Y: np.ndarray, X: np.ndarray, W: np.ndarray, fixperc_q: bool = False,
                              return_R2: bool = True, limit_gamma: bool = False) -> Any:
    """Per-gene weighted slope + offset under box constraints ``m in [1e-8, 20]``, ``q in [0, 2 sum(yw)/sum(w)]``
    (velocyto/estimation.py:337-366; L-BFGS-B with numeric gradients, :212-241).  Returns the exact constrained
    optimum; SciPy's iterate differs from it by more than 1e-5 on a few percent of genes."""
    # fixperc_q takes precedence and ignores limit_gamma, as in the reference (estimation.py:220-224)
    g, q, r2 = _fit(3, Y, X, W, 1e-8, 20.0, want_r2=return_R2, fixperc_q=fixperc_q,
                    limit_gamma=limit_gamma and not fixperc_q)
    return (g, q, r2) if return_R2 else (g, q)
