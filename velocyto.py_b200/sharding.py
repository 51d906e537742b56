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
