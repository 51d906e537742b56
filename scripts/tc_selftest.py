"""GPU self-test / diagnostics of the tensor-core all-pairs linear kernel (K2g, csrc/coldeltacor_tc.cu).

    python scripts/tc_selftest.py random G C [c0 nc]   raw products P, Q and corr vs fp64 / vs the fp32 kernel K2
    python scripts/tc_selftest.py onehot                single-gene inputs: which gene positions reach the MMA
    python scripts/tc_selftest.py time G C              K2g vs K2 timing (CUDA events)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import velocyto_b200  # noqa: E402,F401
from velocyto_b200 import device as dev  # noqa: E402


def make(G, C, seed=0, c0=0, nc=None):
    rng = np.random.default_rng(seed)
    gene_mu = rng.gamma(0.6, 2.0, G)[:, None] + 0.05
    e = rng.gamma(2.0, 1.0, (G, C)) * gene_mu                   # genes with very different levels
    z = rng.normal(size=(G, C))
    d = np.sqrt(np.abs(z) + 1.0) * np.sign(z) + 0.3 * (e - e.mean(1, keepdims=True)) / (e.std(1, keepdims=True) + 1e-9)
    nc = C - c0 if nc is None else nc
    return e.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64), c0, nc


def reference(e, d, c0, nc):
    """fp64: centred operands, raw products and the correlation (speedboosted.pyx:30-78 algebra)."""
    X = e - e.mean(1, keepdims=True)
    X = X - X.mean(0, keepdims=True)                             # (G, C) centred over genes per cell
    B = d[:, c0:c0 + nc] - d[:, c0:c0 + nc].mean(0, keepdims=True)
    P = B.T @ X                                                  # nc x C
    Q = X[:, c0:c0 + nc].T @ X
    qd = (X * X).sum(0)
    pcc = np.einsum("gc,gc->c", B, X[:, c0:c0 + nc])
    sbb = (B * B).sum(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        corr = (P - pcc[:, None]) / np.sqrt((qd[c0:c0 + nc, None] + qd[None, :] - 2 * Q) * sbb[:, None])
    corr[np.arange(nc), c0 + np.arange(nc)] = np.nan
    return P, Q, corr


def run_random(G, C, c0=0, nc=None):
    e, d, c0, nc = make(G, C, 0, c0, nc)
    E = dev.CellMajor.from_gene_major(e)
    D = dev.CellMajor.from_gene_major(d[:, c0:c0 + nc])
    out, P, Q = dev.coldeltacor_linear_tc(E, D, c0=c0, debug=True)
    torch.cuda.synchronize()
    Pw, Qw, cw = reference(e, d, c0, nc)
    P, Q, out = P.cpu().numpy().astype(np.float64), Q.cpu().numpy().astype(np.float64), out.cpu().numpy().astype(np.float64)
    sP, sQ = np.abs(Pw).max(), np.abs(Qw).max()
    eP, eQ = np.abs(P - Pw).max() / sP, np.abs(Q - Qw).max() / sQ
    print(f"[random G={G} C={C} c0={c0} nc={nc}] P err/max {eP:.3e}  Q err/max {eQ:.3e}")
    if eP > 1e-3 or eQ > 1e-3:
        np.set_printoptions(precision=4, linewidth=200, suppress=True)
        print("P got\n", P[:4, :8], "\nP want\n", Pw[:4, :8])
        print("Q got\n", Q[:4, :8], "\nQ want\n", Qw[:4, :8])
        bad = np.abs(Q - Qw) > 1e-3 * sQ
        print("Q bad fraction", bad.mean(), "bad rows mod 8", np.bincount(np.where(bad)[0] % 8, minlength=8),
              "bad cols mod 8", np.bincount(np.where(bad)[1] % 8, minlength=8))
        print("ratio got/want median", np.median(Q[~np.isnan(Q)] / Qw[~np.isnan(Q)]))
    ok = ~np.isnan(cw)
    nan_match = np.array_equal(np.isnan(out), np.isnan(cw))
    ec = np.abs(out - cw)[ok & ~np.isnan(out)].max()
    # the fp32 register-tiled kernel on the same inputs
    k2 = dev.coldeltacor(E, D, None, "linear", 0.0, c0=c0).cpu().numpy().astype(np.float64)
    ek2 = np.abs(k2 - cw)[ok & ~np.isnan(k2)].max()
    print(f"   corr max abs err: K2g {ec:.3e}   K2(fp32) {ek2:.3e}   NaN pattern equal: {nan_match}")
    return eP < 1e-4 and eQ < 1e-4 and ec < 1e-6 and nan_match


def run_onehot():
    G, C = 256, 128
    for g0 in (0, 1, 7, 8, 15, 16, 31, 32, 47, 63, 64, 65, 127, 128, 200, 255):
        rng = np.random.default_rng(g0)
        e = np.zeros((G, C))
        e[g0] = rng.normal(size=C) * 3
        d = np.zeros((G, C))
        d[g0] = rng.normal(size=C)
        d[(g0 + 1) % G] = 0.01
        E = dev.CellMajor.from_gene_major(e)
        D = dev.CellMajor.from_gene_major(d)
        _, P, Q = dev.coldeltacor_linear_tc(E, D, debug=True)
        torch.cuda.synchronize()
        Pw, Qw, _ = reference(e.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64), 0, C)
        eP = np.abs(P.cpu().numpy() - Pw).max() / np.abs(Pw).max()
        eQ = np.abs(Q.cpu().numpy() - Qw).max() / np.abs(Qw).max()
        print(f"[onehot g0={g0:3d}] P err {eP:.2e}  Q err {eQ:.2e}")


def run_time(G, C):
    rng = np.random.default_rng(1)
    E = dev.CellMajor.empty(C, G)
    D = dev.CellMajor.empty(C, G)
    E.t[:, :G] = torch.from_numpy(rng.gamma(2.0, 1.0, (C, G)).astype(np.float32)).cuda()
    D.t[:, :G] = torch.from_numpy(rng.normal(size=(C, G)).astype(np.float32)).cuda()
    stats = dev.cell_stats(D)
    out = torch.empty((C, C), dtype=torch.float32, device="cuda")
    from velocyto_b200 import _cabi
    lib = _cabi.load()
    for name, fn in (("K2g tensor", lambda: dev.coldeltacor_linear_tc(E, D, stats=stats, out=out)),
                     ("K2 fp32  ", lambda: dev.coldeltacor(E, D, None, "linear", 0.0, stats=stats, out=out))):
        lib.velo_set_tensor_cores(0 if name.startswith("K2 ") else 1)
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        pg = G * C * C
        print(f"[time G={G} C={C}] {name}: {ms:9.2f} ms   {pg / ms / 1e9:8.2f} T pair-gene/s"
              f"   ({12 * pg / ms / 1e12:7.3f} PFLOP/s at 12 flop per pair-gene)")
    lib.velo_set_tensor_cores(1)


def run_bias(G, C):
    """Accumulation bias of long tensor-core sums: drain the TMEM block sums every D 64-gene groups."""
    e, d, c0, nc = make(G, C, 3)
    E = dev.CellMajor.from_gene_major(e)
    D = dev.CellMajor.from_gene_major(d)
    Pw, Qw, cw = reference(e, d, 0, C)
    ok = ~np.isnan(cw)
    big = np.abs(Qw) > 0.05 * np.abs(Qw).max()
    for dg in (1, 2, 4, 16, 64, 4096):
        os.environ["VELO_TC_DRAIN_GROUPS"] = str(dg)
        out, P, Q = dev.coldeltacor_linear_tc(E, D, debug=True)
        torch.cuda.synchronize()
        P, Q, out = (t.cpu().numpy().astype(np.float64) for t in (P, Q, out))
        relQ = (Q - Qw)[big] / Qw[big]
        print(f"[bias G={G} C={C}] drain every {dg:4d} x 64 genes: Q rel err mean {relQ.mean():+.3e} rms {relQ.std():.3e}"
              f"   |P| err/max {np.abs(P - Pw).max() / np.abs(Pw).max():.3e}   corr max abs err {np.abs(out - cw)[ok].max():.3e}")
    os.environ.pop("VELO_TC_DRAIN_GROUPS")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "random":
        a = [int(x) for x in sys.argv[2:]]
        ok = run_random(*a)
        sys.exit(0 if ok else 1)
    elif mode == "onehot":
        run_onehot()
    elif mode == "time":
        run_time(int(sys.argv[2]), int(sys.argv[3]))
    elif mode == "bias":
        run_bias(int(sys.argv[2]), int(sys.argv[3]))
