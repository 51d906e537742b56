"""K1 tie-resolving (EXACT) variant vs the fast variant at the bench's gene/neighbour shape, one library per process
(VELO_B200_LIB selects a tuning build, e.g. -DVELO_K1_THREADS_EXACT=1024)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from velocyto_b200 import device as dev

def timeit(fn, iters=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

C, G, m = int(os.environ.get("C", 20000)), 30000, 3000
gen = torch.Generator(device="cuda").manual_seed(0)
e, d = dev.CellMajor.empty(C, G), dev.CellMajor.empty(C, G)
lo = torch.zeros_like(e.t)
for r0 in range(0, C, 4096):
    n = min(4096, C - r0)
    u = torch.rand((n, G), device="cuda", generator=gen)
    v = -torch.log(torch.rand_like(u).clamp_min(1e-7)) - torch.log(u.clamp_min(1e-7))
    v[torch.rand_like(u) < 0.3] = 0
    e.t[r0:r0+n, :G] = v
    lo[r0:r0+n, :G] = (torch.rand((n, G), device="cuda", generator=gen) - 0.5) * 2.0 ** -27 * v
    z = torch.randn((n, G), device="cuda", generator=gen)
    d.t[r0:r0+n, :G] = torch.sign(z) * torch.sqrt(z.abs() + 1.0)
ar = torch.arange(C, device="cuda")[:, None]
ix = ((ar + 1 + torch.randint(0, C - 1, (C, m), device="cuda", generator=gen)) % C).to(torch.int32).contiguous()
stats = dev.cell_stats(d)
out = torch.empty((C, m), dtype=torch.float32, device="cuda")
alg = C * (m + 2) * G * 4
e_exact = dev.CellMajor(e.t, G, lo)
for name, em in (("fast", e), ("exact", e_exact)):
    ms = timeit(lambda: dev.coldeltacor(em, d, ix, "sqrt", 1.0, stats=stats, out=out))
    print(json.dumps(dict(lib=os.environ.get("VELO_B200_LIB", "default"), variant=name, ms=ms, gbs=alg / ms / 1e6,
                          cells_per_s=C / ms * 1e3)))
