"""K1 launches for an ncu capture: uniform-random vs embedding-local neighbour lists, fast vs tie-resolving (EXACT)
variant, on the full 100k x 30k expression matrix but only 2 cells per SM (296 cells), so that one launch is ~16 ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velocyto_b200 import device as dev

C, G, m, nc = int(os.environ.get("C", 100_000)), 30_000, 3_000, int(os.environ.get("NC", 296))
gen = torch.Generator(device="cuda").manual_seed(0)
e = dev.CellMajor.empty(C, G)
for r0 in range(0, C, 4096):
    n = min(4096, C - r0)
    u = torch.rand((n, G), device="cuda", generator=gen)
    v = -torch.log(torch.rand_like(u).clamp_min(1e-7)) - torch.log(u.clamp_min(1e-7))
    v[torch.rand_like(u) < 0.3] = 0
    e.t[r0:r0 + n, :G] = v
d = dev.CellMajor.empty(nc, G)
z = torch.randn((nc, G), device="cuda", generator=gen)
d.t[:, :G] = torch.sign(z) * torch.sqrt(z.abs() + 1.0)
c0 = 40_000
ar = torch.arange(c0, c0 + nc, device="cuda")[:, None]
ix_rand = ((ar + 1 + torch.randint(0, C - 1, (nc, m), device="cuda", generator=gen)) % C).to(torch.int32).contiguous()
win = 10_000
pick = torch.rand((nc, win), device="cuda", generator=gen).topk(m, dim=1).indices
off = pick - win // 2
off = off + (off >= 0).to(off.dtype)
ix_local = ((ar + off) % C).to(torch.int32).contiguous()
lo = torch.zeros_like(e.t)
lo[:, :G] = (torch.rand((C, G), device="cuda", generator=gen) - 0.5) * 2.0 ** -27 * e.t[:, :G]
stats = dev.cell_stats(d)
out = torch.empty((nc, m), dtype=torch.float32, device="cuda")
for name, em, ix in (("random", e, ix_rand), ("local", e, ix_local), ("random_exact", dev.CellMajor(e.t, G, lo), ix_rand)):
    for _ in range(2):
        dev.coldeltacor(em, d, ix, "sqrt", 1.0, c0=c0, stats=stats, out=out)
    torch.cuda.synchronize()
    print("done", name)
