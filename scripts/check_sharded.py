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
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
