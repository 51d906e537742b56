"""Device-memory containers for the cell-major fp32 layout (torch tensors as plain HBM buffers).

Layout (DESIGN.md "Data layout in HBM"): a matrix the reference holds as gene-major
``(genes, cells)`` float64 lives on the device as ``x[cell, gene]`` float32 with a row stride
``ld`` that is a multiple of 32 floats (128 B), pad columns zero.  A cell's expression profile
is then one contiguous, TMA-/128-bit-load-friendly row; the neighbour gathers of
``colDeltaCor*partial`` (speedboosted.pyx:279-282, stride = cells doubles in the reference)
become contiguous row reads.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _cabi

ROW_ALIGN = 32          # floats; 128-byte rows


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _cabi.VeloError("no CUDA device visible: velocyto.py_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def padded_ld(G: int) -> int:
    return (G + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN


class CellMajor:
    """``(cells, ld)`` float32 CUDA tensor holding a ``(genes, cells)`` matrix transposed."""

    __slots__ = ("t", "G", "lo")

    def __init__(self, tensor: torch.Tensor, G: int, lo: Optional[torch.Tensor] = None):
        assert tensor.is_cuda and tensor.dtype == torch.float32 and tensor.dim() == 2 and tensor.is_contiguous()
        assert tensor.shape[1] % 4 == 0 and tensor.shape[1] >= G
        # lo: optional fp32 residuals (e64 - e32) of fp64-origin data, same layout; lets the correlation kernel
        # resolve fp32 ties the way the fp64 reference sees them
        assert lo is None or (lo.shape == tensor.shape and lo.dtype == torch.float32 and lo.is_contiguous())
        self.t, self.G, self.lo = tensor, int(G), lo

    @property
    def C(self) -> int:
        return self.t.shape[0]

    @property
    def ld(self) -> int:
        return self.t.shape[1]

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @classmethod
    def empty(cls, C: int, G: int, device: Optional[torch.device] = None) -> "CellMajor":
        device = device or require_cuda()
        t = torch.empty((C, padded_ld(G)), dtype=torch.float32, device=device)
        if t.shape[1] != G:
            t[:, G:].zero_()                   # only the pad columns need defined values
        return cls(t, G)

    @classmethod
    def from_gene_major(cls, arr, chunk_bytes: int = 256 << 20, residual: bool = False) -> "CellMajor":
        """Upload a host ``(genes, cells)`` array (float64/float32; any strides) or a CUDA tensor.

        The host copy goes over in gene-row chunks and is transposed/converted on the device
        (``velo_dev_pack_cellmajor``); nothing is transposed on the host.  ``residual=True`` keeps the fp32
        residuals of float64 data (``.lo``; dropped again when the data turns out to be fp32-representable).
        """
        device = require_cuda()
        if isinstance(arr, torch.Tensor):
            src = arr
            G, C = src.shape
        else:
            arr = np.asarray(arr)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            G, C = arr.shape
            if arr.flags.f_contiguous and not arr.flags.c_contiguous and not (residual and arr.dtype == np.float64):
                # physically cell-major already (what scipy's sparse product returns)
                out = cls.empty(C, G, device)
                out.t[:, :G].copy_(torch.from_numpy(arr.T).to(device, non_blocking=False))
                return out
            arr = np.ascontiguousarray(arr)
            src = None
        out = cls.empty(C, G, device)
        esz = 8 if (src.dtype == torch.float64 if src is not None else arr.dtype == np.float64) else 4
        split = residual and esz == 8
        if split:
            out.lo = torch.zeros_like(out.t)
            flag = torch.zeros(1, dtype=torch.int32, device=device)
        rows_per = max(32, min(G, chunk_bytes // max(1, C * esz)))
        for g0 in range(0, G, rows_per):
            g1 = min(G, g0 + rows_per)
            if src is not None:
                blk = src[g0:g1].contiguous()
                if not blk.is_cuda:
                    blk = blk.to(device)
            else:
                blk = torch.from_numpy(arr[g0:g1]).to(device)
            if split:
                _cabi.call("velo_dev_pack_cellmajor_split", blk.data_ptr(), esz, g1 - g0, C, out.ptr, out.lo.data_ptr(),
                           flag.data_ptr(), out.ld, g0, _stream_ptr())
            else:
                _cabi.call("velo_dev_pack_cellmajor", blk.data_ptr(), esz, g1 - g0, C, out.ptr, out.ld, g0, _stream_ptr())
            del blk
        if split and int(flag.item()) == 0:
            out.lo = None                      # exactly representable in fp32: nothing to resolve
        return out

    def to_gene_major(self, dtype=np.float64) -> np.ndarray:
        """Download as the reference's ``(genes, cells)`` host array."""
        dt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        dst = torch.empty((self.G, self.C), dtype=dt, device=self.t.device)
        _cabi.call("velo_dev_unpack_genemajor", self.ptr, self.ld, self.G, self.C, dst.data_ptr(),
                   8 if dt == torch.float64 else 4, _stream_ptr())
        return dst.cpu().numpy()

    def rows(self, c0: int, nc: int) -> "CellMajor":
        return CellMajor(self.t[c0:c0 + nc], self.G, None if self.lo is None else self.lo[c0:c0 + nc])


def indices_to_device(ixs, C: int) -> torch.Tensor:
    """``(cells, m)`` neighbour indices -> contiguous int32 CUDA tensor (validated on the way)."""
    device = require_cuda()
    if isinstance(ixs, torch.Tensor):
        t = ixs.to(device)
    else:
        t = torch.from_numpy(np.ascontiguousarray(ixs)).to(device)
    if t.numel() and (int(t.min()) < 0 or int(t.max()) >= C):
        raise ValueError(f"neighbour index outside [0, {C})")
    return t.to(torch.int32).contiguous()


def cell_stats(d_cm: CellMajor) -> torch.Tensor:
    """Per-cell mean and centred sum of squares of the velocity rows (speedboosted.pyx:46-55,67-72)."""
    stats = torch.empty((d_cm.C, 2), dtype=torch.float32, device=d_cm.t.device)
    _cabi.call("velo_dev_cell_stats", d_cm.ptr, d_cm.ld, d_cm.G, d_cm.C, stats.data_ptr(), _stream_ptr())
    return stats


def coldeltacor(e_cm: CellMajor, d_cm: CellMajor, ixs: Optional[torch.Tensor], transform: str, psc: float,
                rule: Optional[int] = None, c0: int = 0, stats: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Compact correlation ``out[r, n]`` for local cells ``c0 .. c0+nc`` (``nc = d_cm.C``).

    ``e_cm`` holds ALL cells (neighbours may be anywhere); ``d_cm`` only the local rows.
    ``ixs``: ``(nc, m)`` int32 global neighbour ids, or None for the full (all-pairs) variants.
    """
    tr = _cabi.TRANSFORMS[transform]
    nc, C, G = d_cm.C, e_cm.C, e_cm.G
    assert d_cm.G == G and d_cm.ld == e_cm.ld
    if rule is None:
        rule = _cabi.RULE_FULL if ixs is None else _cabi.RULE_PARTIAL
    m = C if ixs is None else ixs.shape[1]
    if ixs is not None:
        assert ixs.is_cuda and ixs.dtype == torch.int32 and ixs.is_contiguous() and ixs.shape[0] == nc
    if stats is None:
        stats = cell_stats(d_cm)
    if out is None:
        out = torch.empty((nc, m), dtype=torch.float32, device=e_cm.t.device)
    _cabi.call("velo_dev_coldeltacor_ex", tr, rule, e_cm.ptr, 0 if e_cm.lo is None else e_cm.lo.data_ptr(), d_cm.ptr,
               e_cm.ld, stats.data_ptr(), 0 if ixs is None else ixs.data_ptr(), m, out.data_ptr(), out.stride(0),
               G, C, c0, nc, m, float(psc), _stream_ptr())
    return out


def coldeltacor_linear_tc(e_cm: CellMajor, d_cm: CellMajor, c0: int = 0, stats: Optional[torch.Tensor] = None,
                          out: Optional[torch.Tensor] = None, debug: bool = False):
    """All-pairs linear correlation on the tensor cores (K2g; ``x_colDeltaCor``, speedboosted.pyx:13-87):
    ``out[r, i]`` for local cells ``c0 .. c0+nc`` against every cell.  ``debug=True`` also returns the raw
    products ``(P, Q)``."""
    nc, C, G = d_cm.C, e_cm.C, e_cm.G
    assert d_cm.G == G and d_cm.ld == e_cm.ld
    if stats is None:
        stats = cell_stats(d_cm)
    if out is None:
        out = torch.empty((nc, C), dtype=torch.float32, device=e_cm.t.device)
    P = torch.zeros_like(out) if debug else None
    Q = torch.zeros_like(out) if debug else None
    _cabi.call("velo_dev_coldeltacor_tc", e_cm.ptr, d_cm.ptr, e_cm.ld, stats.data_ptr(), out.data_ptr(), out.stride(0),
               G, C, c0, nc, 0 if P is None else P.data_ptr(), 0 if Q is None else Q.data_ptr(), _stream_ptr())
    return (out, P, Q) if debug else out


def transition_prob(corr: torch.Tensor, ixs: Optional[torch.Tensor], sigma: float, c0: int = 0,
                    out: Optional[torch.Tensor] = None, patch_nan: bool = True) -> torch.Tensor:
    """Compact ``exp(corr/sigma)`` row-normalised (analysis.py:1697-1698) with the self->0 patch and, with
    ``patch_nan`` (the knn_random branch, analysis.py:1604-1612), NaN->1; ``patch_nan=False`` is the "full" branch
    (analysis.py:1666-1668): a NaN correlation turns its whole row NaN, as in the reference."""
    nc, m = corr.shape
    if out is None:
        out = torch.empty_like(corr)
    _cabi.call("velo_dev_transition_prob_ex", corr.data_ptr(), corr.stride(0), 0 if ixs is None else ixs.data_ptr(),
               0 if ixs is None else ixs.stride(0), out.data_ptr(), out.stride(0), c0, nc, m, float(sigma),
               int(bool(patch_nan)), _stream_ptr())
    return out


def scatter_dense(compact: torch.Tensor, ixs: Optional[torch.Tensor], C: int, c0: int = 0,
                  rm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dense ``(C, C)`` float64 adapter: ``rm[c0+r, ixs[r, n]] += compact[r, n]`` (small C only)."""
    nc, m = compact.shape
    if rm is None:
        rm = torch.zeros((C, C), dtype=torch.float64, device=compact.device)
    _cabi.call("velo_dev_scatter_dense", compact.data_ptr(), compact.stride(0), 0 if ixs is None else ixs.data_ptr(),
               0 if ixs is None else ixs.stride(0), rm.data_ptr(), C, c0, nc, m, _stream_ptr())
    return rm


# --------------------------------------------------------------------------- K4: gamma fits
FIT_SLOPE, FIT_SLOPE_OFFSET, FIT_SLOPE_WEIGHTED, FIT_SLOPE_WEIGHTED_OFFSET = 0, 1, 2, 3


def fit_constraints(S_cm: CellMajor, U_cm: CellMajor, fixperc_q: bool = False, limit_gamma: bool = False):
    """Per-gene constraints of the non-default fit options: ``(q_fix, up_gamma)`` float64 CUDA vectors (or None).
    Percentiles of masked subsets (estimation.py:199-204, 221-224) by radix select on the device."""
    G, device = S_cm.G, S_cm.t.device
    qf = torch.empty(G, dtype=torch.float64, device=device) if fixperc_q else None
    up = torch.empty(G, dtype=torch.float64, device=device) if limit_gamma else None
    if qf is not None or up is not None:
        _cabi.call("velo_dev_fit_constraints", S_cm.ptr, U_cm.ptr, S_cm.ld, G, S_cm.C,
                   0 if qf is None else qf.data_ptr(), 0 if up is None else up.data_ptr(), _stream_ptr())
    return qf, up


def fit_gammas(mode: int, S_cm: CellMajor, U_cm: CellMajor, W_cm: Optional[CellMajor] = None,
               cell_mask: Optional[torch.Tensor] = None, lo: float = 0.0, hi: float = 20.0,
               want_r2: bool = False, want_moments: bool = False,
               hi_per_gene: Optional[torch.Tensor] = None, q_fixed: Optional[torch.Tensor] = None):
    """Batched per-gene fit (estimation.py:173-366).  Returns ``(gamma, offset, r2, moments)`` CUDA tensors
    (float32 ``(G,)``; ``r2`` / ``moments`` None unless requested)."""
    G, C, device = S_cm.G, S_cm.C, S_cm.t.device
    assert U_cm.G == G and U_cm.C == C and U_cm.ld == S_cm.ld
    gamma = torch.empty(G, dtype=torch.float32, device=device)
    offset = torch.zeros(G, dtype=torch.float32, device=device)
    r2 = torch.empty(G, dtype=torch.float32, device=device) if want_r2 else None
    mom = torch.empty((14, G), dtype=torch.float64, device=device) if want_moments else None
    if cell_mask is not None:
        cell_mask = cell_mask.to(device=device, dtype=torch.uint8).contiguous()
        assert cell_mask.numel() == C
    if mode >= 2:
        assert W_cm is not None and W_cm.G == G and W_cm.C == C
    _cabi.call("velo_dev_fit_gammas_ex", mode, S_cm.ptr, U_cm.ptr, S_cm.ld,
               0 if W_cm is None else W_cm.ptr, 0 if W_cm is None else W_cm.ld,
               0 if cell_mask is None else cell_mask.data_ptr(), G, C, float(lo), float(hi),
               0 if hi_per_gene is None else hi_per_gene.data_ptr(), 0 if q_fixed is None else q_fixed.data_ptr(),
               gamma.data_ptr(), offset.data_ptr(), 0 if r2 is None else r2.data_ptr(),
               0 if mom is None else mom.data_ptr(), _stream_ptr())
    return gamma, offset, r2, mom


# --------------------------------------------------------------------------- K6: elementwise chain
def velocity_chain(S_cm: CellMajor, U_cm: CellMajor, gamma: torch.Tensor, q: Optional[torch.Tensor] = None,
                   assumption: str = "constant_velocity", dt_shift: float = 1.0, dt_extrap: float = 1.0,
                   clip: bool = True, transform: str = "sqrt", psc: float = 1e-10, eps: Optional[float] = None,
                   want=("Upred", "velocity", "delta_S", "S_t", "d")):
    """predict_U -> calculate_velocity -> calculate_shift -> extrapolate_cell_at_t -> transform, one pass
    (analysis.py:1321-1439, 1577/1597).  Returns a dict of CellMajor outputs named in ``want``."""
    G, C, device = S_cm.G, S_cm.C, S_cm.t.device
    gamma = gamma.to(device=device, dtype=torch.float32).contiguous()
    if q is not None:
        q = q.to(device=device, dtype=torch.float32).contiguous()
    thr = None
    if eps:
        thr = torch.empty(G, dtype=torch.float32, device=device)
        _cabi.call("velo_dev_velocity_threshold", S_cm.ptr, S_cm.ld, gamma.data_ptr(),
                   0 if q is None else q.data_ptr(), G, C, float(eps), thr.data_ptr(), _stream_ptr())
    outs = {k: CellMajor.empty(C, G, device) for k in want}
    ptr = lambda k: outs[k].ptr if k in outs else 0
    _cabi.call("velo_dev_velocity_chain", S_cm.ptr, U_cm.ptr, S_cm.ld, gamma.data_ptr(),
               0 if q is None else q.data_ptr(), 0 if thr is None else thr.data_ptr(), G, C,
               {"constant_velocity": 0, "constant_unspliced": 1}[assumption], float(dt_shift), float(dt_extrap),
               int(bool(clip)), _cabi.TRANSFORMS[transform], float(psc),
               ptr("Upred"), ptr("velocity"), ptr("delta_S"), ptr("S_t"), ptr("d"), _stream_ptr())
    return outs


# --------------------------------------------------------------------------- K5: kNN smoothing
def knn_smooth(indptr, indices, weights, S_cm: CellMajor, maximum: bool = False) -> CellMajor:
    """``Sx[c, :] = sum_n w[c, n] * S[idx[c, n], :]`` for a CSR-by-rows weight matrix (neighbors.py:416-423)."""
    device = S_cm.t.device
    as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(
        device=device, dtype=dt).contiguous()
    indptr, indices, weights = as_t(indptr, torch.int64), as_t(indices, torch.int32), as_t(weights, torch.float32)
    assert indptr.numel() == S_cm.C + 1
    out = CellMajor.empty(S_cm.C, S_cm.G, device)
    _cabi.call("velo_dev_knn_smooth", indptr.data_ptr(), indices.data_ptr(), weights.data_ptr(), S_cm.ptr, out.ptr,
               S_cm.ld, S_cm.G, S_cm.C, int(bool(maximum)), _stream_ptr())
    return out


def delta_transform(delta_S: CellMajor, dt: float, transform: str, psc: float) -> CellMajor:
    """``d = f(dt * delta_S)``: the velocity argument of colDeltaCor* (analysis.py:1577/1594/1597)."""
    out = CellMajor(torch.empty_like(delta_S.t), delta_S.G)
    _cabi.call("velo_dev_delta_transform", delta_S.ptr, out.ptr, delta_S.ld, delta_S.C, float(dt),
               _cabi.TRANSFORMS[transform], float(psc), _stream_ptr())
    return out


def extrapolate(S: CellMajor, delta_S: CellMajor, dt: float, clip: bool) -> CellMajor:
    """``clip(S + dt * delta_S, 0)`` (analysis.py:1429-1431)."""
    out = CellMajor(torch.empty_like(S.t), S.G)
    _cabi.call("velo_dev_extrapolate", S.ptr, delta_S.ptr, out.ptr, S.ld, S.C, float(dt), int(bool(clip)), _stream_ptr())
    return out


def patch_corr(corr: torch.Tensor, ixs: Optional[torch.Tensor], c0: int = 0, patch_nan: bool = True) -> int:
    """In place: self pair -> 0, NaN -> 1 (analysis.py:1604-1612).  Returns the number of NaNs replaced."""
    cnt = torch.zeros(1, dtype=torch.int64, device=corr.device)
    nc, m = corr.shape
    _cabi.call("velo_dev_patch_corr", corr.data_ptr(), corr.stride(0), 0 if ixs is None else ixs.data_ptr(),
               0 if ixs is None else ixs.stride(0), c0, nc, m, int(bool(patch_nan)), cnt.data_ptr(), _stream_ptr())
    return int(cnt.item())


def embedding_shift(tp: torch.Tensor, ixs: torch.Tensor, embedding, c0: int = 0) -> torch.Tensor:
    """``delta_embedding`` without expression scaling (analysis.py:1704-1712), ``(nc, 2)`` float64."""
    emb = torch.as_tensor(np.ascontiguousarray(embedding, dtype=np.float64)).to(tp.device)
    nc, m = tp.shape
    out = torch.empty((nc, 2), dtype=torch.float64, device=tp.device)
    _cabi.call("velo_dev_embedding_shift", tp.data_ptr(), tp.stride(0), ixs.data_ptr(), ixs.stride(0), emb.data_ptr(),
               emb.shape[1], c0, nc, m, out.data_ptr(), _stream_ptr())
    return out


# --------------------------------------------------------------------------- normalisation (analysis.py:535-676)
def cell_sums(X: CellMajor) -> torch.Tensor:
    """``X.sum(0)`` of the reference's gene-major matrix: per-cell totals, float64 CUDA vector."""
    out = torch.empty(X.C, dtype=torch.float64, device=X.t.device)
    _cabi.call("velo_dev_cell_sums", X.ptr, X.ld, X.G, X.C, out.data_ptr(), _stream_ptr())
    return out


def size_normalize(X: CellMajor, factor: Optional[torch.Tensor], pcount: float = 1.0, want_sz: bool = True,
                   want_norm: bool = True, nonfinite_to_zero: bool = False):
    """``(factor * X, log2(factor * X + pcount))`` in one pass; ``factor``: per-cell float64 CUDA vector or None."""
    sz = CellMajor(torch.empty_like(X.t), X.G) if want_sz else None
    nm = CellMajor(torch.empty_like(X.t), X.G) if want_norm else None
    if factor is not None:
        assert factor.is_cuda and factor.dtype == torch.float64 and factor.numel() == X.C and factor.is_contiguous()
    _cabi.call("velo_dev_size_normalize", X.ptr, X.ld, X.G, X.C, 0 if factor is None else factor.data_ptr(),
               float(pcount), int(bool(nonfinite_to_zero)), 0 if sz is None else sz.ptr, 0 if nm is None else nm.ptr,
               _stream_ptr())
    return sz, nm


# --------------------------------------------------------------------------- fit_gammas weights (analysis.py:1179-1219)
WEIGHT_KINDS = {"maxmin_diag": 0, "maxmin": 1, "maxmin_double": 2, "sum": 3, "prod": 4, "maxmin_weighted": 5}


def row_percentiles(M: CellMajor, q) -> torch.Tensor:
    """``np.percentile(M, q, axis=cells)`` per gene ("linear" interpolation), ``(G, len(q))`` float64 CUDA tensor."""
    G, C, device = M.G, M.C, M.t.device
    rows = torch.empty((G, C), dtype=torch.float32, device=device)
    _cabi.call("velo_dev_unpack_genemajor", M.ptr, M.ld, G, C, rows.data_ptr(), 4, _stream_ptr())
    qd = torch.as_tensor(np.atleast_1d(np.asarray(q, dtype=np.float64))).to(device)
    out = torch.empty((G, qd.numel()), dtype=torch.float64, device=device)
    _cabi.call("velo_dev_row_percentiles", rows.data_ptr(), G, C, qd.data_ptr(), qd.numel(), out.data_ptr(), _stream_ptr())
    return out


def fit_weights(kind: str, S: CellMajor, U: CellMajor, Sx: Optional[CellMajor] = None, Ux: Optional[CellMajor] = None,
                maxmin_perc=(2, 98), maxmin_weighted_pow: float = 15) -> CellMajor:
    """Weight matrix of the gamma fit built on the device (per-gene radix-select percentiles)."""
    W = CellMajor.empty(S.C, S.G, S.t.device)
    _cabi.call("velo_dev_fit_weights_ex", WEIGHT_KINDS[kind], S.ptr, U.ptr, 0 if Sx is None else Sx.ptr,
               0 if Ux is None else Ux.ptr, S.ld, S.G, S.C, float(maxmin_perc[0]), float(maxmin_perc[1]),
               float(maxmin_weighted_pow), W.ptr, W.ld, _stream_ptr())
    return W


def logratio(S: CellMajor, delta_S: Optional[CellMajor], dt: float, psc: float, which: int) -> CellMajor:
    """``transform="logratio"`` operands (analysis.py:1582-1583): 0 -> log2(S+psc), 1 -> log2(|S+dt*dS|+psc) - log2(S+psc)."""
    out = CellMajor(torch.empty_like(S.t), S.G)
    _cabi.call("velo_dev_logratio", S.ptr, 0 if delta_S is None else delta_S.ptr, out.ptr, S.ld, S.C, float(dt), float(psc),
               int(which), _stream_ptr())
    return out


def expression_scaling(tp: torch.Tensor, ixs: torch.Tensor, hi_dim: CellMajor, delta_S: CellMajor,
                       penalty: float = 1.0) -> torch.Tensor:
    """``scaling`` of calculate_embedding_shift (analysis.py:1714-1719): the expected expression change
    ``sum_n (P[c,n] - 1/m) * hi_dim[:, ixs[c,n]]`` (a K5 row gather with signed weights) projected on ``delta_S``."""
    nc, m = tp.shape
    indptr = torch.arange(0, nc * m + 1, m, device=tp.device, dtype=torch.int64)
    w = (tp - 1.0 / m).reshape(-1).contiguous()
    estim = knn_smooth(indptr, ixs.reshape(-1), w, hi_dim)
    scale = torch.empty(nc, dtype=torch.float64, device=tp.device)
    _cabi.call("velo_dev_row_cosine_scale", delta_S.ptr, estim.ptr, estim.ld, estim.G, nc, float(penalty), scale.data_ptr(),
               _stream_ptr())
    return scale


def knn_smooth_csr(w_indptr, w_indices, w_weights, S_csr, g0: int = 0, ng: Optional[int] = None,
                   maximum: bool = False) -> CellMajor:
    """kNN smoothing with sparse counts.  ``S_csr``: scipy CSR ``(cells, genes)`` (or a ``(indptr, indices, data, G)``
    tuple of arrays/tensors) with sorted column indices; returns the dense smoothed gene slab ``[g0, g0+ng)`` as a
    ``(cells, ng)`` cell-major matrix (gene-sharding: every rank asks for its own slab)."""
    device = require_cuda()
    as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(
        device=device, dtype=dt).contiguous()
    if isinstance(S_csr, tuple):
        s_ip, s_ix, s_v, G = S_csr
    else:
        S_csr = S_csr.tocsr()
        if not S_csr.has_sorted_indices:
            S_csr = S_csr.sorted_indices()
        s_ip, s_ix, s_v, G = S_csr.indptr, S_csr.indices, S_csr.data, S_csr.shape[1]
    s_ip, s_ix, s_v = as_t(s_ip, torch.int64), as_t(s_ix, torch.int32), as_t(s_v, torch.float32)
    w_indptr, w_indices, w_weights = as_t(w_indptr, torch.int64), as_t(w_indices, torch.int32), as_t(w_weights, torch.float32)
    C = w_indptr.numel() - 1
    assert s_ip.numel() == C + 1
    ng = G - g0 if ng is None else ng
    out = CellMajor.empty(C, ng, device)
    _cabi.call("velo_dev_knn_smooth_csr", w_indptr.data_ptr(), w_indices.data_ptr(), w_weights.data_ptr(),
               s_ip.data_ptr(), s_ix.data_ptr(), s_v.data_ptr(), out.ptr, out.ld, C, int(g0), int(ng),
               int(bool(maximum)), _stream_ptr())
    return out


# --------------------------------------------------------------------------- sparse ingest
class CsrCounts:
    """A ``(genes, cells)`` count matrix held on the device as CSR BY CELL (``indptr`` over cells, sorted gene ids,
    fp32 values) -- the form sparse ``.loom`` / 10x / AnnData HDF5 files store, uploaded as it is (8 bytes per non-zero
    instead of 8 bytes per ELEMENT of the dense float64 matrix the reference's loader builds, analysis.py:56-64)."""

    __slots__ = ("indptr", "genes", "values", "G")

    def __init__(self, indptr: torch.Tensor, genes: torch.Tensor, values: torch.Tensor, G: int):
        assert indptr.is_cuda and indptr.dtype == torch.int64 and genes.dtype == torch.int32 and values.dtype == torch.float32
        self.indptr, self.genes, self.values, self.G = indptr.contiguous(), genes.contiguous(), values.contiguous(), int(G)

    @property
    def C(self) -> int:
        return self.indptr.numel() - 1

    @property
    def nnz(self) -> int:
        return self.genes.numel()

    @classmethod
    def from_arrays(cls, indptr, indices, data, G: int) -> "CsrCounts":
        """From the three arrays of a by-cell compressed layout (HDF5 ``indptr`` / ``indices`` / ``data``)."""
        device = require_cuda()
        as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(
            device=device).to(dt)
        return cls(as_t(indptr, torch.int64), as_t(indices, torch.int32), as_t(data, torch.float32), G)

    @classmethod
    def from_scipy(cls, M) -> "CsrCounts":
        """From a SciPy sparse matrix in the reference's orientation ``(genes, cells)``: its CSC form IS CSR by cell."""
        from scipy import sparse
        M = sparse.csc_matrix(M)
        if not M.has_sorted_indices:
            M = M.sorted_indices()
        return cls.from_arrays(M.indptr, M.indices, M.data, M.shape[0])

    def cell_sums(self) -> torch.Tensor:
        out = torch.empty(self.C, dtype=torch.float64, device=self.values.device)
        _cabi.call("velo_dev_csr_cell_sums_scale", self.indptr.data_ptr(), self.values.data_ptr(), self.C, 0, out.data_ptr(),
                   _stream_ptr())
        return out

    def scaled(self, factor: torch.Tensor) -> "CsrCounts":
        """``factor[cell] * X`` with the same pattern (size normalisation; non-finite products -> 0)."""
        assert factor.is_cuda and factor.dtype == torch.float64 and factor.numel() == self.C
        out = CsrCounts(self.indptr, self.genes, self.values.clone(), self.G)
        _cabi.call("velo_dev_csr_cell_sums_scale", out.indptr.data_ptr(), out.values.data_ptr(), out.C,
                   factor.contiguous().data_ptr(), 0, _stream_ptr())
        return out

    def to_cellmajor(self, g0: int = 0, ng: Optional[int] = None) -> CellMajor:
        """Dense cell-major gene slab ``[g0, g0 + ng)`` (default: all genes) on the device."""
        ng = self.G - g0 if ng is None else ng
        out = CellMajor(torch.empty((self.C, padded_ld(ng)), dtype=torch.float32, device=self.values.device), ng)
        _cabi.call("velo_dev_csr_to_cellmajor", self.indptr.data_ptr(), self.genes.data_ptr(), self.values.data_ptr(),
                   self.C, int(g0), int(ng), out.ptr, out.ld, _stream_ptr())
        return out


# --------------------------------------------------------------------------- exact kNN
KNN_MAX_K = 14000
# metrics the brute-force kernel serves: Euclidean directly; correlation / cosine through the identity
# |x^ - y^|^2 = 2 (1 - <x^, y^>) on rows scaled to unit norm (after centring, for correlation) -- the neighbour ORDER
# under the reference's scikit-learn metric (neighbors.py:239-243) is the Euclidean order of the standardised rows
KNN_DEVICE_METRICS = ("euclidean", "l2", "minkowski", "correlation", "cosine")


def knn(points, k: int, include_self: bool = False, metric: str = "euclidean", q0: int = 0, nq: Optional[int] = None,
        want_dist: bool = True):
    """Exact k nearest neighbours of every row of ``points`` (C x D), ascending by distance, under ``metric``
    ("euclidean" | "correlation" | "cosine").  Returns ``(idx int32 (C, k), dist float64 (C, k))`` CUDA tensors;
    ``dist`` is the metric's own distance (``1 - corr`` / ``1 - cos`` for the two angular metrics).
    ``q0`` / ``nq``: only the query block ``[q0, q0 + nq)`` (rank-sharded searches)."""
    device = require_cuda()
    if metric not in KNN_DEVICE_METRICS:
        raise ValueError(f"metric={metric!r} is not served by the device kNN ({KNN_DEVICE_METRICS})")
    X = (points if isinstance(points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)))
    X = X.to(device=device, dtype=torch.float64).contiguous()
    angular = metric in ("correlation", "cosine")
    if angular:
        if metric == "correlation":
            X = X - X.mean(dim=1, keepdim=True)
        X = (X / torch.linalg.vector_norm(X, dim=1, keepdim=True)).contiguous()
    C, D = X.shape
    nq = C - q0 if nq is None else nq
    idx = torch.empty((nq, k), dtype=torch.int32, device=device)
    dist = torch.empty((nq, k), dtype=torch.float64, device=device) if want_dist else None
    _cabi.call("velo_dev_knn_range", X.data_ptr(), C, D, int(k), int(bool(include_self)), int(q0), int(nq), idx.data_ptr(),
               0 if dist is None else dist.data_ptr(), _stream_ptr())
    if angular and dist is not None:
        dist = dist * dist * 0.5
    return idx, dist


def knn_query(points, queries, k: int):
    """The k nearest ``points`` (C x D) of every row of ``queries`` (nq x D), Euclidean, ascending -- what
    ``NearestNeighbors().fit(points).kneighbors(queries)`` returns (calculate_grid_arrows, analysis.py:1788-1790).
    Returns ``(idx int32 (nq, k), dist float64 (nq, k))`` CUDA tensors."""
    device = require_cuda()
    as_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))).to(
        device=device, dtype=torch.float64).contiguous()
    X, Q = as_t(points), as_t(queries)
    assert X.shape[1] == Q.shape[1]
    idx = torch.empty((Q.shape[0], k), dtype=torch.int32, device=device)
    dist = torch.empty((Q.shape[0], k), dtype=torch.float64, device=device)
    _cabi.call("velo_dev_knn_query", X.data_ptr(), X.shape[0], X.shape[1], Q.data_ptr(), Q.shape[0], int(k), idx.data_ptr(),
               dist.data_ptr(), _stream_ptr())
    return idx, dist


def grid_flow(neighs: torch.Tensor, dists: torch.Tensor, delta, sigma: float):
    """Gaussian-kernel average of the cells' embedding displacements around every grid point
    (analysis.py:1792-1797): returns ``(total_p_mass (npts,), flow (npts, dims))`` float64 CUDA tensors."""
    delta = torch.as_tensor(np.ascontiguousarray(delta, dtype=np.float64)).to(neighs.device)
    npts, k = neighs.shape
    mass = torch.empty(npts, dtype=torch.float64, device=neighs.device)
    flow = torch.empty((npts, delta.shape[1]), dtype=torch.float64, device=neighs.device)
    _cabi.call("velo_dev_grid_flow", neighs.data_ptr(), dists.data_ptr(), npts, k, delta.data_ptr(), delta.shape[1],
               float(sigma), mass.data_ptr(), flow.data_ptr(), _stream_ptr())
    return mass, flow


# --------------------------------------------------------------------------- PCA (analysis.py:678-702)
class PCAResult:
    """The attributes of ``sklearn.decomposition.PCA`` that velocyto reads after ``perform_PCA`` (``explained_variance_ratio_``
    for choosing ``n_pca_dims``, doc/tutorial/analysis.rst:118-121; ``components_`` / ``mean_`` for projections)."""

    def __init__(self, components, explained_variance, explained_variance_ratio, singular_values, mean, n_samples):
        self.components_, self.explained_variance_ = components, explained_variance
        self.explained_variance_ratio_, self.singular_values_, self.mean_ = explained_variance_ratio, singular_values, mean
        self.n_components_, self.n_features_in_, self.n_samples_ = components.shape[0], components.shape[1], n_samples

    def transform(self, X):
        return (np.asarray(X) - self.mean_) @ self.components_.T


PCA_EXACT_MAX_DIM = 4096        # symmetric eigendecomposition up to this order (~1 s); beyond it, subspace iteration


def pca(X: CellMajor, n_components: Optional[int] = None, div_by_std: bool = False, block: int = 8192,
        n_iter: int = 12, seed: int = 0):
    """PCA of the cells (rows = samples, genes = features) as ``PCA(n_components).fit_transform(X.T)`` defines it
    (analysis.py:697-702): features centred, components = leading right singular vectors with scikit-learn's sign rule
    (largest-magnitude loading positive), ``pcs = U * S``.

    The matrix stays on the device (fp32 storage, fp64 arithmetic over cell blocks).  Two solvers:

    * exact -- fp64 second moments (gene x gene covariance, or the cell x cell Gram matrix when there are fewer cells
      than genes) and one symmetric eigendecomposition -- whenever the smaller dimension is at most
      ``PCA_EXACT_MAX_DIM`` or most components are wanted;
    * block subspace iteration with Rayleigh-Ritz extraction otherwise (``n_iter`` applications of X_c^T X_c to a
      random ``n + max(16, n)``-column block, QR after each): the deterministic-seed counterpart of the randomised
      solver scikit-learn's ``svd_solver="auto"`` picks for such shapes (7 power iterations, 10 oversamples, and a
      different random block on every call) -- same class of approximation, run much further.

    cuSOLVER / cuBLAS through torch.linalg: plain library calls, PCA is not part of the hand-written hot path.
    Returns ``(pcs (C, n) float64 CUDA, PCAResult)``."""
    C, G, dv = X.C, X.G, X.t.device
    n = min(C, G) if n_components is None else int(n_components)
    if not 0 < n <= min(C, G):
        raise ValueError(f"n_components={n_components} must be between 1 and min(n_samples, n_features)={min(C, G)}")

    def rows(r0):                                   # one block of samples in fp64 (optionally scaled by 1 / per-cell std)
        Xb = X.t[r0:r0 + block, :G].double()
        if div_by_std:
            Xb = Xb / Xb.std(dim=1, unbiased=False, keepdim=True)            # X.T / X.std(0)   (analysis.py:700)
        return Xb

    mean = torch.zeros(G, dtype=torch.float64, device=dv)
    for r0 in range(0, C, block):
        mean += rows(r0).sum(0)
    mean /= C
    exact = min(C, G) <= PCA_EXACT_MAX_DIM or n > min(C, G) // 4
    if exact and G <= C:                            # covariance route: G x G
        M = torch.zeros((G, G), dtype=torch.float64, device=dv)
        for r0 in range(0, C, block):
            Xb = rows(r0) - mean
            M.addmm_(Xb.t(), Xb)
        lam, V = torch.linalg.eigh(M)
        lam, V = lam.flip(0)[:n].clamp_min(0), V.flip(1)[:, :n]                # descending
        comps = V.t().contiguous()                                              # (n, G)
        del M
    elif exact:                                     # Gram route: C x C, components = X_c^T u / s
        Xc = torch.cat([rows(r0) - mean for r0 in range(0, C, block)])
        lam, Uv = torch.linalg.eigh(Xc @ Xc.t())
        lam, Uv = lam.flip(0)[:n].clamp_min(0), Uv.flip(1)[:, :n]
        comps = (Xc.t() @ Uv / lam.sqrt().clamp_min(1e-300)).t().contiguous()
        comps = comps / torch.linalg.vector_norm(comps, dim=1, keepdim=True).clamp_min(1e-300)
        del Xc
    else:                                           # block subspace iteration on X_c^T X_c (G x G, never formed)
        ell = min(min(C, G), n + max(16, n))
        gen = torch.Generator(device=dv).manual_seed(seed)
        Q = torch.linalg.qr(torch.randn((G, ell), dtype=torch.float64, device=dv, generator=gen)).Q

        def apply(Q):                               # returns X_c^T (X_c Q)  and  (X_c Q)^T (X_c Q)
            Z = torch.zeros((G, ell), dtype=torch.float64, device=dv)
            T = torch.zeros((ell, ell), dtype=torch.float64, device=dv)
            for r0 in range(0, C, block):
                Xb = rows(r0) - mean
                Y = Xb @ Q
                Z.addmm_(Xb.t(), Y)
                T.addmm_(Y.t(), Y)
            return Z, T

        for _ in range(n_iter):
            Z, _ = apply(Q)
            Q = torch.linalg.qr(Z).Q
        _, T = apply(Q)                             # Rayleigh-Ritz on the converged block
        lam, Wv = torch.linalg.eigh(T)
        lam, Wv = lam.flip(0)[:n].clamp_min(0), Wv.flip(1)[:, :n]
        comps = (Q @ Wv).t().contiguous()
    # scikit-learn's svd_flip(u_based_decision=False): the largest-magnitude loading of every component is positive
    piv = comps.abs().argmax(dim=1)
    sign = torch.sign(comps[torch.arange(n, device=dv), piv])
    sign[sign == 0] = 1
    comps = comps * sign[:, None]
    pcs = torch.empty((C, n), dtype=torch.float64, device=dv)
    total_var = torch.zeros((), dtype=torch.float64, device=dv)
    for r0 in range(0, C, block):
        Xb = rows(r0) - mean
        pcs[r0:r0 + block] = Xb @ comps.t()
        total_var += (Xb * Xb).sum()
    ev = lam / (C - 1)
    res = PCAResult(comps.cpu().numpy(), ev.cpu().numpy(), (ev / (total_var / (C - 1))).cpu().numpy(),
                    lam.sqrt().cpu().numpy(), mean.cpu().numpy(), C)
    res.solver = "exact" if exact else f"subspace_iteration(n_iter={n_iter}, block={ell})"
    return pcs, res


# --------------------------------------------------------------------------- device-side randomisation (opt-in)
def sample_neighbors(knn_idx: torch.Tensor, p, m: int, seed: int):
    """Weighted sampling without replacement of ``m`` of the ``W`` candidates of every cell on the device
    (``velo_dev_sample_neighbors``): the distribution of ``np.random.choice(W, m, replace=False, p=p)`` per cell
    (analysis.py:1561-1564), not its MT19937 stream.  Returns ``(neigh_ixs, sampling_ixs)``, ``(C, m)`` int32 CUDA."""
    assert knn_idx.is_cuda and knn_idx.dtype == torch.int32 and knn_idx.dim() == 2 and knn_idx.is_contiguous()
    C, W = knn_idx.shape
    p = np.asarray(p, dtype=np.float64)
    if p.shape != (W,) or not np.all(p > 0):
        raise ValueError("p must hold one positive probability per candidate")
    inv_p = torch.from_numpy((1.0 / p).astype(np.float32)).to(knn_idx.device)
    neigh = torch.empty((C, m), dtype=torch.int32, device=knn_idx.device)
    samp = torch.empty((C, m), dtype=torch.int32, device=knn_idx.device)
    _cabi.call("velo_dev_sample_neighbors", knn_idx.data_ptr(), C, W, inv_p.data_ptr(), int(m),
               int(seed) & 0xFFFFFFFFFFFFFFFF, neigh.data_ptr(), samp.data_ptr(), _stream_ptr())
    return neigh, samp


def permute_rows_nsign(X: CellMajor, seed: int) -> CellMajor:
    """Randomised control of ``estimate_transition_prob`` (``permute_rows_nsign``, analysis.py:2413-2420) on the device:
    every gene gets its own pseudo-random permutation of the cells and independent random signs."""
    out = CellMajor(torch.empty_like(X.t), X.G)
    if X.ld != X.G:
        out.t[:, X.G:].zero_()
    _cabi.call("velo_dev_permute_rows_nsign", X.ptr, out.ptr, X.ld, X.G, X.C, int(seed) & 0xFFFFFFFFFFFFFFFF,
               _stream_ptr())
    return out
