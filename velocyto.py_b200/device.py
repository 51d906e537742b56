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

    __slots__ = ("t", "G")

    def __init__(self, tensor: torch.Tensor, G: int):
        assert tensor.is_cuda and tensor.dtype == torch.float32 and tensor.dim() == 2 and tensor.is_contiguous()
        assert tensor.shape[1] % 4 == 0 and tensor.shape[1] >= G
        self.t, self.G = tensor, int(G)

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
        return cls(torch.zeros((C, padded_ld(G)), dtype=torch.float32, device=device), G)

    @classmethod
    def from_gene_major(cls, arr, chunk_bytes: int = 256 << 20) -> "CellMajor":
        """Upload a host ``(genes, cells)`` array (float64/float32; any strides) or a CUDA tensor.

        The host copy goes over in gene-row chunks and is transposed/converted on the device
        (``velo_dev_pack_cellmajor``); nothing is transposed on the host.
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
            if arr.flags.f_contiguous and not arr.flags.c_contiguous:
                # physically cell-major already (what scipy's sparse product returns)
                out = cls.empty(C, G, device)
                out.t[:, :G].copy_(torch.from_numpy(arr.T).to(device, non_blocking=False))
                return out
            arr = np.ascontiguousarray(arr)
            src = None
        out = cls.empty(C, G, device)
        esz = 8 if (src.dtype == torch.float64 if src is not None else arr.dtype == np.float64) else 4
        rows_per = max(32, min(G, chunk_bytes // max(1, C * esz)))
        for g0 in range(0, G, rows_per):
            g1 = min(G, g0 + rows_per)
            if src is not None:
                blk = src[g0:g1].contiguous()
                if not blk.is_cuda:
                    blk = blk.to(device)
            else:
                blk = torch.from_numpy(arr[g0:g1]).to(device)
            _cabi.call("velo_dev_pack_cellmajor", blk.data_ptr(), esz, g1 - g0, C, out.ptr, out.ld, g0, _stream_ptr())
            del blk
        return out

    def to_gene_major(self, dtype=np.float64) -> np.ndarray:
        """Download as the reference's ``(genes, cells)`` host array."""
        dt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        dst = torch.empty((self.G, self.C), dtype=dt, device=self.t.device)
        _cabi.call("velo_dev_unpack_genemajor", self.ptr, self.ld, self.G, self.C, dst.data_ptr(),
                   8 if dt == torch.float64 else 4, _stream_ptr())
        return dst.cpu().numpy()

    def rows(self, c0: int, nc: int) -> "CellMajor":
        return CellMajor(self.t[c0:c0 + nc], self.G)


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
    _cabi.call("velo_dev_coldeltacor", tr, rule, e_cm.ptr, d_cm.ptr, e_cm.ld, stats.data_ptr(),
               0 if ixs is None else ixs.data_ptr(), m, out.data_ptr(), out.stride(0),
               G, C, c0, nc, m, float(psc), _stream_ptr())
    return out


def transition_prob(corr: torch.Tensor, ixs: Optional[torch.Tensor], sigma: float, c0: int = 0,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Compact ``exp(corr/sigma)`` row-normalised with self->0 / NaN->1 patches (analysis.py:1604-1612,1697-1698)."""
    nc, m = corr.shape
    if out is None:
        out = torch.empty_like(corr)
    _cabi.call("velo_dev_transition_prob", corr.data_ptr(), corr.stride(0), 0 if ixs is None else ixs.data_ptr(),
               0 if ixs is None else ixs.stride(0), out.data_ptr(), out.stride(0), c0, nc, m, float(sigma),
               _stream_ptr())
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
