"""Cell-sharded execution of the transition-probability core across the GPUs of one box.

The correlation rows are independent: rank ``r`` owns the contiguous cell block
``[r*b, min(C, (r+1)*b))`` with ``b = ceil(C / world)`` and needs, besides its own velocity rows
and neighbour lists, the expression rows of ALL cells (a neighbour may live anywhere).  That is
the path's single exchange step: one NCCL all-gather of the per-rank cell-major ``e`` blocks
over NVLink (12 GB at 100k x 30k fp32).  With uniform blocks of ``b`` rows, the gathered buffer
row index equals the global cell id, so neighbour indices need no translation.  Outputs stay
sharded.  One process per GPU; ``torch.distributed`` is only the plumbing.

The reference has no distributed code at all (OpenMP ``prange`` over cells inside one process,
speedboosted.pyx:23); this module replaces that loop-level parallelism at box scale.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def block_size(C: int, world: int) -> int:
    return (C + world - 1) // world


def partition(C: int, world: int) -> List[Tuple[int, int]]:
    """``(c0, nc)`` of every rank; uniform block ``b = ceil(C/world)``, trailing blocks may be short/empty."""
    b = block_size(C, world)
    return [(min(C, r * b), max(0, min(C, (r + 1) * b) - min(C, r * b))) for r in range(world)]


def gather_cell_blocks(local: torch.Tensor, b: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather ``(nc_r, ld)`` row blocks into ``(world*b, ld)``; rows of short blocks are zero-padded.

    Works on CUDA tensors over NCCL (the product path) and on CPU tensors over gloo (the
    world_size-2 logic tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    ld = local.shape[1]
    if world == 1:
        return local
    if local.shape[0] != b:
        padded = torch.zeros((b, ld), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        local = padded
    if out is None:
        out = torch.empty((world * b, ld), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class CellShardedTransitionProb:
    """``estimate_transition_prob`` numeric core (correlation + softmax epilogue), cell-sharded.

    Every rank calls :meth:`run` with ITS block: ``e_local`` / ``d_local`` ``(nc, ld)`` cell-major
    fp32, ``ix_local`` ``(nc, m)`` int32 GLOBAL neighbour ids; it returns the block's ``(nc, m)``
    transition probabilities (or correlations with ``sigma=None``).
    """

    def __init__(self, G: int, C: int, transform: str = "sqrt", psc: float = 1e-10,
                 sigma: Optional[float] = 0.05, group=None):
        self.G, self.C, self.transform, self.psc, self.sigma, self.group = G, C, transform, psc, sigma, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.b = block_size(C, self.world)
        self.c0, self.nc = partition(C, self.world)[self.rank]
        self._e_full: Optional[torch.Tensor] = None

    def run(self, e_local, d_local, ix_local: torch.Tensor, kernel_events=None) -> torch.Tensor:
        from . import device as dev
        assert e_local.C == self.nc and d_local.C == self.nc and ix_local.shape[0] == self.nc
        if self.world > 1:
            if self._e_full is None:
                self._e_full = torch.empty((self.world * self.b, e_local.ld), dtype=torch.float32,
                                           device=e_local.t.device)
            full = gather_cell_blocks(e_local.t, self.b, self.group, self._e_full)
            e_all = dev.CellMajor(full, self.G)
        else:
            e_all = e_local
        stats = dev.cell_stats(d_local)
        if kernel_events is not None:
            kernel_events[0].record()
        corr = dev.coldeltacor(e_all, d_local, ix_local, self.transform, self.psc, c0=self.c0, stats=stats)
        if kernel_events is not None:
            kernel_events[1].record()
        if self.sigma is None:
            return corr
        return dev.transition_prob(corr, ix_local, self.sigma, c0=self.c0, out=corr)


def _host_ptr(a):
    """``(address, row pitch in elements, element bytes)`` of a 2-D host array (NumPy or CPU torch) whose rows are
    contiguous -- e.g. a cell block ``M[:, c0:c0+nc]`` of the reference's gene-major matrix, passed without a copy."""
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.dim() == 2 and (a.stride(1) == 1 or a.shape[1] == 1)
        assert a.dtype in (torch.float64, torch.float32, torch.int64)
        return a.data_ptr(), a.stride(0) if a.shape[0] > 1 else a.shape[1], a.element_size()
    import numpy as np
    assert isinstance(a, np.ndarray) and a.ndim == 2 and (a.strides[1] == a.itemsize or a.shape[1] == 1)
    assert a.dtype in (np.float64, np.float32, np.int64)
    return a.ctypes.data, (a.strides[0] // a.itemsize) if a.shape[0] > 1 else a.shape[1], a.itemsize


def needs_residuals(transform: str, psc: float, elem_bytes: int) -> bool:
    """fp64 inputs + a transform that jumps at zero difference (sqrt with psc > 0, log10 with psc != 1): the fp32
    residuals of ``e`` travel with it so that fp32 ties keep the fp64 sign (same rule as the one-GPU host tier)."""
    import math
    if elem_bytes != 8:
        return False
    if transform == "sqrt":
        return 2.0 * math.sqrt(max(psc, 0.0)) > 1e-4
    if transform in ("log", "log10"):
        return 2.0 * abs(math.log10(psc if psc > 0 else 1e-300)) > 1e-4
    return False


class CellShardedHostTransitionProb:
    """``velo_transition_prob_partial`` across the GPUs of one box, from HOST buffers in the reference's format.

    Every rank calls :meth:`run` with its cell block: ``e_block`` / ``d_block`` gene-major ``(G, nc)`` float64 (or
    float32) host arrays -- column blocks of the reference's matrices, NumPy views are fine, pinned or pageable --
    ``ixs_block`` ``(nc, m)`` int64 GLOBAL neighbour ids and ``out_block`` ``(nc, m)`` float32.  Per rank:

      1. ``velo_upload_cellmajor``: the expression block goes through PCIe straight into this rank's slot of the
         gathered ``(world*b, ld)`` matrix (transpose + fp32 conversion on the device);
      2. the path's one exchange step: an in-place NCCL all-gather of those slots over NVLink;
      3. ``velo_transition_prob_partial_sharded``: velocity rows, neighbour lists and results move in cell chunks
         underneath the correlation kernel; the first velocity chunk uploads while the all-gather is in flight.

    Nothing but plumbing happens in Python (the reference's loop-level OpenMP parallelism,
    speedboosted.pyx:22-23, at box scale)."""

    def __init__(self, G: int, C: int, transform: str = "sqrt", psc: float = 1e-10, sigma: Optional[float] = 0.05,
                 group=None):
        self.G, self.C, self.transform, self.psc, self.sigma, self.group = G, C, transform, psc, sigma, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.b = block_size(C, self.world)
        self.c0, self.nc = partition(C, self.world)[self.rank]
        self._e_full = self._lo_full = None

    def run(self, e_block, d_block, ixs_block, out_block):
        from . import _cabi, device as dev
        device = dev.require_cuda()
        G, nc, b, ld = self.G, self.nc, self.b, dev.padded_ld(self.G)
        e_ptr, e_pitch, e_sz = _host_ptr(e_block)
        d_ptr, d_pitch, d_sz = _host_ptr(d_block)
        ix_ptr, ix_pitch, ix_sz = _host_ptr(ixs_block)
        o_ptr, o_pitch, o_sz = _host_ptr(out_block)
        m = ixs_block.shape[1]
        assert tuple(e_block.shape) == (G, nc) and tuple(d_block.shape) == (G, nc) and e_sz == d_sz
        assert tuple(ixs_block.shape) == (nc, m) and tuple(out_block.shape) == (nc, m)
        assert ix_sz == 8 and o_sz == 4 and ix_pitch == m and o_pitch == m
        if self._e_full is None:
            self._e_full = torch.zeros((self.world * b, ld), dtype=torch.float32, device=device)
        want_lo = needs_residuals(self.transform, self.psc, e_sz)
        if want_lo and self._lo_full is None:
            self._lo_full = torch.zeros((self.world * b, ld), dtype=torch.float32, device=device)
        stream = torch.cuda.current_stream().cuda_stream
        import ctypes
        # The block goes up in SUB-BLOCKS of cells; the all-gather of sub-block s (NCCL's own stream, NVLink) runs while
        # sub-block s + 1 is still crossing PCIe.  All ranks cut at the same local rows (multiples of sb inside the
        # uniform block b), so the collectives line up; rows past a short last block are padding nobody indexes.
        # The residual sub-blocks are gathered optimistically whenever the input CAN need them (their traffic hides
        # under the upload); whether the kernel uses them is decided once, by one all-reduce of the per-rank flags.
        S = 4 if (self.world > 1 and b >= 4 * 1024) else 1
        sb = (b + S - 1) // S
        nz_any, works = 0, []
        for s0 in range(0, b, sb):
            s1 = min(b, s0 + sb)
            n_up = max(0, min(nc, s1) - s0)                       # rows of this sub-block that exist on this rank
            lo_off = self.rank * b + s0
            if n_up:
                nz = ctypes.c_int(0)
                _cabi.call("velo_upload_cellmajor", e_ptr + s0 * e_sz, e_sz, G, n_up, e_pitch,
                           self._e_full[lo_off:].data_ptr(), self._lo_full[lo_off:].data_ptr() if want_lo else 0,
                           ctypes.addressof(nz) if want_lo else 0, ld, stream)
                nz_any |= nz.value
            if self.world > 1:
                if S == 1:
                    mine = self._e_full[self.rank * b:(self.rank + 1) * b]
                    dist.all_gather_into_tensor(self._e_full, mine, group=self.group)          # in place: slot `rank`
                    if want_lo:
                        dist.all_gather_into_tensor(self._lo_full, self._lo_full[self.rank * b:(self.rank + 1) * b],
                                                    group=self.group)
                else:
                    for full in ((self._e_full, self._lo_full) if want_lo else (self._e_full,)):
                        outs = [full[r * b + s0:r * b + s1] for r in range(self.world)]
                        works.append(dist.all_gather(outs, full[lo_off:self.rank * b + s1], group=self.group, async_op=True))
        for w in works:
            w.wait()                                              # stream-level: the current stream waits for the gathers
        use_lo = False
        if want_lo:
            flag = torch.tensor([nz_any], device=device, dtype=torch.int32)
            if self.world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            use_lo = bool(flag.item())                      # any rank's block not fp32-representable
        if nc:
            _cabi.call("velo_transition_prob_partial_sharded", _cabi.TRANSFORMS[self.transform], self._e_full.data_ptr(),
                       self._lo_full.data_ptr() if use_lo else 0, ld, stream, d_ptr, d_sz, d_pitch, ix_ptr, o_ptr,
                       G, self.C, self.c0, nc, m, float(self.psc), float(self.sigma) if self.sigma else 0.0)
        return out_block


# --------------------------------------------------------------------------- gene-sharded stages (K4 / K5 / K6)
def gene_partition(G: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """``(g0, ng)`` of every rank: contiguous gene blocks, starts aligned to ``align`` genes (128-byte rows)."""
    b = (G + world - 1) // world
    b = (b + align - 1) // align * align
    return [(min(G, r * b), max(0, min(G, (r + 1) * b) - min(G, r * b))) for r in range(world)]


def genes_to_cells(x_local: torch.Tensor, G: int, group=None) -> torch.Tensor:
    """Re-shard from gene blocks to cell blocks: the hand-over between the gene-sharded stages
    (kNN smoothing, gamma fit, elementwise chain -- every gene is independent) and the cell-sharded
    correlation stage.

    x_local: ``(C, ng_r)`` -- ALL cells, this rank's gene block (cell-major).  Returns ``(nc_r, G)``: this rank's
    cell block (``partition``), all genes.  One all-to-all exchange; works on CUDA/NCCL and CPU/gloo tensors."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    C = x_local.shape[0]
    if world == 1:
        return x_local[:, :G].contiguous()
    rank = dist.get_rank(group)
    cparts, gparts = partition(C, world), gene_partition(G, world)
    assert x_local.shape[1] >= gparts[rank][1]
    ng = gparts[rank][1]
    send = [x_local[c0:c0 + nc, :ng].contiguous() for c0, nc in cparts]        # my genes, each rank's cells
    nc_me = cparts[rank][1]
    recv = [torch.empty((nc_me, gparts[r][1]), dtype=x_local.dtype, device=x_local.device) for r in range(world)]
    # pairwise exchange posted as one batch (NCCL fuses it into a single grouped all-to-all over NVLink; gloo,
    # which has no alltoall, runs the same code in the CPU logic tests)
    recv[rank].copy_(send[rank])
    ops = []
    for r in range(world):
        if r != rank:
            ops.append(dist.P2POp(dist.isend, send[r], r, group))
            ops.append(dist.P2POp(dist.irecv, recv[r], r, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return torch.cat(recv, dim=1)                                               # (nc_me, G) in gene order
