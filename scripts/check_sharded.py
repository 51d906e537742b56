"""Run under torchrun with N >= 2 GPUs: the cell-sharded NCCL path must reproduce the single-GPU result bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from velocyto_b200 import device as dev
from velocyto_b200.sharding import CellShardedTransitionProb, partition

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
G, C, m, psc = 3001, 1003, 77, 1.0            # uneven blocks on purpose (1003 cells over `world` ranks)
rng = np.random.default_rng(0)                # same data on every rank
e = rng.gamma(2.0, 1.0, (G, C)).astype(np.float32).astype(np.float64)
e[rng.uniform(size=e.shape) < 0.3] = 0
z = rng.normal(size=(G, C))
d = (np.sqrt(np.abs(z) + psc) * np.sign(z)).astype(np.float32).astype(np.float64)
ixs = np.stack([(c + 1 + rng.choice(C - 1, m, replace=False)) % C for c in range(C)])
e_all, d_all = dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d)
ix_all = dev.indices_to_device(ixs, C)
core = CellShardedTransitionProb(G, C, "sqrt", psc, 0.05)
c0, nc = partition(C, world)[rank]
assert (core.c0, core.nc) == (c0, nc)
mine = core.run(e_all.rows(c0, nc), d_all.rows(c0, nc), ix_all[c0:c0 + nc].contiguous())
# single-GPU reference of the same rows (no collective)
whole = dev.transition_prob(dev.coldeltacor(e_all, d_all, ix_all, "sqrt", psc), ix_all, 0.05)
ok = torch.equal(mine, whole[c0:c0 + nc])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded x{world} == single-GPU: {bool(flag.item())}; rows sum to 1: {float((whole.sum(1) - 1).abs().max()):.2e}")

# ---- stage 2: gene-sharded smoothing + fit + chain, re-shard (all-to-all), cell-sharded correlation ----
from velocyto_b200.sharding import gene_partition, genes_to_cells
k = 25
nbr = np.stack([np.concatenate([[c], (c + 1 + rng.choice(C - 1, k, replace=False)) % C]) for c in range(C)])
indptr = np.arange(0, C * (k + 1) + 1, k + 1)
wts = np.full(C * (k + 1), 1.0 / (k + 1), dtype=np.float32)
U = (e * rng.uniform(0.2, 0.8, (G, 1))).astype(np.float32).astype(np.float64)

def pipeline(S_cm, U_cm):
    """smooth -> OLS fit -> velocity chain on whatever gene range the matrices cover"""
    Sx = dev.knn_smooth(indptr, nbr.reshape(-1), wts, S_cm)
    Ux = dev.knn_smooth(indptr, nbr.reshape(-1), wts, U_cm)
    gam, q, _, _ = dev.fit_gammas(dev.FIT_SLOPE_OFFSET, Sx, Ux)
    out = dev.velocity_chain(Sx, Ux, gam, q, transform="sqrt", psc=psc, want=("d",))
    return Sx, out["d"], gam

Sx_all, d2_all, gam_all = pipeline(e_all, dev.CellMajor.from_gene_major(U))            # single GPU, all genes
ref2 = dev.transition_prob(dev.coldeltacor(Sx_all, d2_all, ix_all, "sqrt", psc), ix_all, 0.05)
g0, ng = gene_partition(G, world)[rank]
Sx_g, d_g, gam_g = pipeline(dev.CellMajor.from_gene_major(e[g0:g0 + ng]), dev.CellMajor.from_gene_major(U[g0:g0 + ng]))
ok_g = torch.equal(gam_g, gam_all[g0:g0 + ng]) and torch.equal(Sx_g.t[:, :ng], Sx_all.t[:, g0:g0 + ng])

def to_cm(x):                                       # (nc, G) -> padded cell-major container
    cm = dev.CellMajor.empty(x.shape[0], G)
    cm.t[:, :G] = x
    return cm

Sx_c = to_cm(genes_to_cells(Sx_g.t[:, :ng].contiguous(), G))
d_c = to_cm(genes_to_cells(d_g.t[:, :ng].contiguous(), G))
mine2 = core.run(Sx_c, d_c, ix_all[c0:c0 + nc].contiguous())
ok2 = ok_g and torch.equal(mine2, ref2[c0:c0 + nc])
flag2 = torch.tensor([1 if ok2 else 0], device="cuda")
dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"gene-sharded smooth+fit+chain -> all-to-all -> cell-sharded corr == single-GPU: {bool(flag2.item())}")

# ---- stage 3: the host-buffer front (CellShardedHostTransitionProb): sub-block upload pipelined with the all-gathers,
#      raw float64 data (residual matrix + EXACT kernel), blocks handed over as strided views of the full host matrices ----
from velocyto_b200.sharding import CellShardedHostTransitionProb
G3, C3, m3 = 257, 4 * 4096 * world + 37, 33                  # b >= 4096: the 4-sub-block path; uneven last block
rng3 = np.random.default_rng(5)
e3 = rng3.gamma(2.0, 1.0, (G3, C3))
e3[rng3.uniform(size=e3.shape) < 0.3] = 0
d3 = np.sqrt(np.abs(rng3.normal(size=(G3, C3))) + psc) * np.sign(rng3.normal(size=(G3, C3)))
ix3 = ((np.arange(C3)[:, None] + 1 + rng3.integers(0, C3 - 1, (C3, m3))) % C3).astype(np.int64)
host = CellShardedHostTransitionProb(G3, C3, "sqrt", psc, 0.05)
out3 = np.empty((host.nc, m3), dtype=np.float32)
host.run(e3[:, host.c0:host.c0 + host.nc], d3[:, host.c0:host.c0 + host.nc], ix3[host.c0:host.c0 + host.nc], out3)
host.run(e3[:, host.c0:host.c0 + host.nc], d3[:, host.c0:host.c0 + host.nc], ix3[host.c0:host.c0 + host.nc], out3)   # buffers reused
e3d, d3d = dev.CellMajor.from_gene_major(e3, residual=True), dev.CellMajor.from_gene_major(d3)
ix3d = dev.indices_to_device(ix3, C3)
ref3 = dev.transition_prob(dev.coldeltacor(e3d, d3d, ix3d, "sqrt", psc), ix3d, 0.05)
ok3 = e3d.lo is not None and torch.equal(torch.from_numpy(out3).cuda(), ref3[host.c0:host.c0 + host.nc])
flag3 = torch.tensor([1 if ok3 else 0], device="cuda")
dist.all_reduce(flag3, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"host-buffer sharded front (sub-block upload + all-gather pipeline, EXACT path) == single-GPU: {bool(flag3.item())}")
dist.destroy_process_group()
sys.exit(0 if (flag.item() and flag2.item() and flag3.item()) else 1)
