"""K1 limiter experiments: access pattern (random vs consecutive neighbour rows) x transform (sqrt vs linear)."""
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
for r0 in range(0, C, 4096):
    n = min(4096, C - r0)
    e.t[r0:r0+n, :G] = torch.rand((n, G), device="cuda", generator=gen) * 3
    d.t[r0:r0+n, :G] = torch.randn((n, G), device="cuda", generator=gen)
ar = torch.arange(C, device="cuda")[:, None]
ix_rand = ((ar + 1 + torch.randint(0, C - 1, (C, m), device="cuda", generator=gen)) % C).to(torch.int32).contiguous()
ix_seq = ((ar + 1 + torch.arange(m, device="cuda")[None, :]) % C).to(torch.int32).contiguous()
stats = dev.cell_stats(d)
out = torch.empty((C, m), dtype=torch.float32, device="cuda")
alg = C * (m + 2) * G * 4
for pat, ix in (("random", ix_rand), ("consecutive", ix_seq)):
    for tr in ("sqrt", "linear", "log10"):
        ms = timeit(lambda: dev.coldeltacor(e, d, ix, tr, 1.0, stats=stats, out=out))
        print(json.dumps(dict(pattern=pat, transform=tr, ms=ms, gbs=alg / ms / 1e6, cells_per_s=C / ms * 1e3)))
