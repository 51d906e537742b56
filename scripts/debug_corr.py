"""GPU diagnostic: per-cell error pattern of the correlation kernel vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import velo_oracle as vo
from velocyto_b200 import device as dev

def run(G, C, m, psc=1.0, seed=0):
    rng = np.random.default_rng(seed)
    e = rng.gamma(2.0, 1.0, (G, C)); z = rng.normal(size=(G, C))
    d = np.sqrt(np.abs(z) + psc) * np.sign(z)
    ixs = np.stack([(c + 1 + rng.choice(C - 1, m, replace=False)) % C for c in range(C)])
    want = vo.colDeltaCorSqrtpartial(e, d, ixs, psc=psc)[np.arange(C)[:, None], ixs]
    e_cm, d_cm = dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d)
    ix = dev.indices_to_device(ixs, C)
    got = dev.coldeltacor(e_cm, d_cm, ix, "sqrt", psc).cpu().numpy()
    err = np.abs(got - want).max(1)
    bad = np.where(err > 5e-7)[0]
    print(f"G={G} C={C} m={m}: max err {err.max():.3e}; bad cells {len(bad)}; first bad {bad[:12]}; "
          f"err cells<148 {err[:148].max():.2e} cells>=148 {err[148:].max() if C>148 else 0:.2e}")
    if len(bad):
        c = bad[0]
        print("   per-neighbour err of first bad cell:", np.round(np.abs(got[c]-want[c])[:10], 7))
        # stats check
        st = dev.cell_stats(d_cm).cpu().numpy()
        mu = d.astype(np.float32).astype(np.float64).mean(0)
        print("   stats mean err", np.abs(st[:,0]-mu).max())
    # two launches of <=148 cells each through the sharded entry
    parts = []
    for c0 in range(0, C, 148):
        nc = min(148, C - c0)
        parts.append(dev.coldeltacor(e_cm, d_cm.rows(c0, nc), ix[c0:c0+nc].contiguous(), "sqrt", psc, c0=c0).cpu().numpy())
    got2 = np.concatenate(parts)
    print(f"   single-cell-per-CTA launches: max err {np.abs(got2-want).max():.3e}")

for cfg in [(1200, 256, 40), (1200, 128, 40), (100, 256, 40), (1200, 600, 8), (37, 400, 6), (512, 300, 40), (640, 300, 40)]:
    run(*cfg)
