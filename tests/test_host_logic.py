"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the product
fails loudly without a GPU (no CPU fallback), argument contracts, and the cell-sharding logic over gloo."""
import os
import re
import socket
import sys

import numpy as np
from scipy import sparse
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "velo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(velo_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from velocyto_b200 import _cabi
    syms = _header_symbols()
    assert len(syms) >= 25 and "velo_colDeltaCorSqrtpartial" in syms and "velo_dev_coldeltacor" in syms
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/velo_b200.h but not exported by libvelo_b200.so"
        assert s in _cabi.PROTOTYPES, f"{s} has no ctypes prototype in _cabi.py"
    assert set(_cabi.PROTOTYPES) <= set(syms)
    assert _cabi.load().velo_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import velocyto_b200.estimation as est
    from velocyto_b200 import VeloError, device as dev
    e = np.random.default_rng(0).random((8, 6))
    ixs = np.tile(np.arange(1, 4), (6, 1))
    with pytest.raises(VeloError, match="no CUDA device|CPU fallback"):
        est.colDeltaCorSqrtpartial(e, e.copy(), ixs, psc=1.0)
    with pytest.raises(VeloError):
        dev.CellMajor.from_gene_major(e)
    with pytest.raises(VeloError):
        est.fit_slope(e, e)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "velocyto.py_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "velo_oracle" not in src and "liboracle" not in src and "from oracle" not in src, f


def test_argument_contract_is_checked_before_the_device():
    import velocyto_b200.estimation as est
    e = np.random.default_rng(0).random((8, 6))
    ixs = np.tile(np.arange(1, 4), (6, 1))
    with pytest.raises(ValueError, match="C-contiguous"):
        est.colDeltaCorSqrtpartial(e, np.asfortranarray(e), ixs)
    with pytest.raises(ValueError, match="dtype"):
        est.colDeltaCorpartial(e, e.astype(np.float32), ixs)
    with pytest.raises(ValueError):
        est.colDeltaCorLog10partial(e, e.copy(), ixs[:3])
    with pytest.raises(ValueError):
        est.fit_slope(e, e[:, :3])


def test_partition_covers_all_cells():
    from velocyto_b200.sharding import block_size, partition
    for C in (1, 7, 100, 100_000, 100_003):
        for w in (1, 2, 3, 4, 8):
            parts = partition(C, w)
            b = block_size(C, w)
            assert len(parts) == w and sum(nc for _, nc in parts) == C
            pos = 0
            for r, (c0, nc) in enumerate(parts):
                assert c0 == min(C, r * b) and 0 <= nc <= b and c0 == pos
                pos += nc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, C, ld, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from velocyto_b200.sharding import block_size, gather_cell_blocks, partition
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        full = torch.arange(C * ld, dtype=torch.float32).reshape(C, ld)
        c0, nc = partition(C, world)[rank]
        b = block_size(C, world)
        got = gather_cell_blocks(full[c0:c0 + nc].clone(), b)
        ok = got.shape == (world * b, ld) and torch.equal(got[:C], full) and bool((got[C:] == 0).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _gloo_reshard_worker(rank, world, port, C, G, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from velocyto_b200.sharding import gene_partition, genes_to_cells, partition
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        full = torch.arange(C * G, dtype=torch.float32).reshape(C, G)
        g0, ng = gene_partition(G, world)[rank]
        c0, nc = partition(C, world)[rank]
        got = genes_to_cells(full[:, g0:g0 + ng].clone(), G)
        q.put((rank, bool(torch.equal(got, full[c0:c0 + nc]))))
    finally:
        dist.destroy_process_group()


def test_gene_to_cell_reshard_world2_gloo():
    """Hand-over between the gene-sharded stages (K4/K5/K6) and the cell-sharded correlation: one all-to-all."""
    import torch.multiprocessing as mp
    from velocyto_b200.sharding import gene_partition
    parts = gene_partition(100, 3)
    assert sum(n for _, n in parts) == 100 and all(g0 % 32 == 0 for g0, n in parts if n)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_reshard_worker, args=(r, 2, port, 11, 70, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(2)) == {0: True, 1: True}


@pytest.mark.parametrize("C", [10, 11])
def test_all_gather_of_cell_blocks_world2_gloo(C):
    """The path's one exchange step: row index of the gathered buffer == global cell id (uneven C included)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, C, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def test_velocytoloom_attribute_surface():
    """Same method names / keyword arguments as the reference's hot methods (analysis.py:933,1120,1321,1452,1670)."""
    import inspect
    from velocyto_b200.analysis import VelocytoLoom
    want = {
        "knn_imputation": ["k", "pca_space", "metric", "diag", "n_pca_dims", "maximum", "size_norm", "balanced",
                           "b_sight", "b_maxl", "group_constraint", "n_jobs"],
        "fit_gammas": ["steady_state_bool", "use_imputed_data", "use_size_norm", "fit_offset", "fixperc_q", "weighted",
                       "weights", "limit_gamma", "maxmin_perc", "maxmin_weighted_pow"],
        "predict_U": ["which_gamma", "which_S", "which_offset"],
        "calculate_velocity": ["kind", "eps"],
        "calculate_shift": ["assumption", "delta_t"],
        "extrapolate_cell_at_t": ["delta_t", "clip"],
        "estimate_transition_prob": ["hidim", "embed", "transform", "ndims", "n_sight", "psc", "knn_random",
                                     "sampled_fraction", "sampling_probs", "max_dist_embed", "n_jobs", "threads",
                                     "calculate_randomized", "random_seed"],
        "calculate_embedding_shift": ["sigma_corr", "expression_scaling", "scaling_penalty"],
    }
    for name, params in want.items():
        sig = inspect.signature(getattr(VelocytoLoom, name))
        assert list(sig.parameters)[1:1 + len(params)] == params, name
    vlm = VelocytoLoom(S=np.ones((3, 4)), U=np.ones((3, 4)))
    vlm.Sx_sz = np.arange(12.0).reshape(3, 4)
    assert vlm.Sx_sz.shape == (3, 4)
    with pytest.raises(AttributeError):
        vlm.Upred


def test_balanced_knn_matches_reference_golden():
    """Host-side restatement of BalancedKNN (neighbors.py:13-321) against the reference's own output."""
    from velocyto_b200.neighbors import BalancedKNN
    g = np.load(os.path.join(ROOT, "tests", "golden", "knn_smoothing_small.npz"))
    pts, groups = g["bknn_points"], g["bknn_groups"]
    for tag, cons in (("plain", None), ("grouped", groups)):
        b = BalancedKNN(k=8, sight_k=30, maxl=12, constraint=cons, mode="distance", n_jobs=1, search="host").fit(pts)
        gph = b.kneighbors_graph(mode="distance")
        assert np.array_equal(gph.indices, g[f"bknn_{tag}_indices"])
        np.testing.assert_allclose(gph.data, g[f"bknn_{tag}_data"], rtol=0, atol=0)
        assert np.array_equal(b.l, g[f"bknn_{tag}_l"]) and b.l.max() <= 12


def test_connectivity_with_diagonal_equals_scipy_path():
    """Fast CSR construction == ``(knn > 0).astype(float); setdiag(diag)`` of the reference (analysis.py:1006-1009)."""
    import warnings
    from scipy import sparse
    from velocyto_b200.analysis import connectivity_with_diagonal
    rng = np.random.default_rng(5)
    n, k = 60, 7
    idx = np.stack([rng.choice(n, k, replace=False) for _ in range(n)])            # may contain the diagonal
    dist = rng.uniform(0.1, 1.0, (n, k))
    dist[rng.uniform(size=dist.shape) < 0.1] = 0.0                                  # zero-distance edges are dropped
    knn = sparse.csr_matrix((dist.ravel(), idx.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    for diag in (1, 8, 0.5):
        want = (knn > 0).astype(float)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want.setdiag(diag)
        got = connectivity_with_diagonal(knn, diag)
        assert np.array_equal(got.toarray(), want.toarray())


def test_integration_stub_binds_against_the_built_library():
    """The ctypes stub INTEGRATION.md tells a maintainer to drop in as velocyto/speedboosted.py must load and bind
    against the library as built (no compute call here: there is no GPU in this tier)."""
    import re
    from velocyto_b200 import _cabi
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# velocyto/speedboosted\.py.*?)```", text, re.S)
    assert m, "INTEGRATION.md lost its speedboosted.py stub"
    src = m.group(1).replace('"libvelo_b200.so"', repr(_cabi.LIB_PATH))
    ns = {}
    exec(compile(src, "INTEGRATION.md:speedboosted.py", "exec"), ns)
    for name in ("_colDeltaCor", "_colDeltaCorSqrt", "_colDeltaCorLog10", "_colDeltaCorpartial",
                 "_colDeltaCorSqrtpartial", "_colDeltaCorLog10partial"):            # speedboosted.pyx:542-610
        assert callable(ns[name]), name
    # the typed-memoryview contract of the .pyx is kept: non-contiguous / non-float64 input is refused before any call
    bad = np.zeros((4, 6), dtype=np.float32)
    with pytest.raises(ValueError):
        ns["_colDeltaCor"](bad, bad, bad, 1)
    with pytest.raises(ValueError):
        ns["_colDeltaCor"](np.zeros((4, 6)).T, np.zeros((6, 4)), np.zeros((6, 6)), 1)


def test_exponential_clock_sampling_is_numpy_choice_without_replacement():
    """The math behind the opt-in device sampler (csrc/random.cu): keeping the `size` smallest keys -log(u)/p, in ascending
    order, has the distribution of np.random.choice(W, size, replace=False, p) -- successive sampling, order included
    (analysis.py:1561-1564).  Checked here with NumPy on both sides: first pick, ordered (first, second) pair and
    per-candidate inclusion frequencies."""
    W, size, n = 7, 3, 40000
    p = np.linspace(0.5, 0.1, W)
    p /= p.sum()
    rng = np.random.default_rng(0)
    keys = -np.log(rng.uniform(size=(n, W))) / p
    mine = np.argsort(keys, axis=1)[:, :size]
    np.random.seed(1)
    ref = np.stack([np.random.choice(W, size=(size,), replace=False, p=p) for _ in range(n)])

    def stats(s):
        first = np.bincount(s[:, 0], minlength=W) / n
        pair = np.zeros((W, W))
        np.add.at(pair, (s[:, 0], s[:, 1]), 1.0 / n)
        incl = np.zeros(W)
        np.add.at(incl, s.ravel(), 1.0 / n)
        return first, pair, incl

    f1, p1, i1 = stats(mine)
    f2, p2, i2 = stats(ref)
    tol = 5 * np.sqrt(0.25 / n)                              # 5 sigma of a frequency estimate, both sides sampled
    assert np.abs(f1 - p).max() < tol and np.abs(f2 - p).max() < tol
    assert np.abs(p1 - p2).max() < 2 * tol
    assert np.abs(i1 - i2).max() < 2 * tol
    # closed form of the ordered pair under successive sampling: p_a * p_b / (1 - p_a)
    want = p[:, None] * p[None, :] / (1 - p[:, None])
    np.fill_diagonal(want, 0.0)
    assert np.abs(p1 - want).max() < tol


def _feistel_numpy(C, g, seed=(0x1234ABCD, 0x9E37)):
    """k_permute_gene_rows (csrc/random.cu) restated in NumPy: pi_g(c) for all c, and the sign bits."""
    def mix32(x):
        x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
        x ^= x >> 16; x = (x * 0x85EBCA6B) & 0xFFFFFFFF
        x ^= x >> 13; x = (x * 0xC2B2AE35) & 0xFFFFFFFF
        x ^= x >> 16
        return x

    h = 1
    while (1 << (2 * h)) < C:
        h += 1
    mask = (1 << h) - 1
    kg = (mix32(np.uint64(g) ^ np.uint64(seed[0])) + seed[1]) & 0xFFFFFFFF
    rk = [int(mix32((kg + 0x9E3779B9 * (r + 1)) & 0xFFFFFFFF)) for r in range(4)]
    x = np.arange(C, dtype=np.uint64)
    todo = np.ones(C, dtype=bool)
    first = True
    while todo.any():
        L, R = x[todo] >> h, x[todo] & mask
        for r in range(4):
            f = ((((R + rk[r]) & 0xFFFFFFFF) * 0x9E3779B1) & 0xFFFFFFFF) >> (32 - h)
            L, R = R, L ^ f
        x[todo] = (L << h) | R
        todo = x >= C if first else todo & (x >= C)
        first = False
    sk = int(mix32(kg ^ 0x5bd1e995))
    c = np.arange(C, dtype=np.uint64)
    sign = (((((c * 0x9E3779B1) & 0xFFFFFFFF) ^ sk) * 0x85EBCA6B) & 0xFFFFFFFF) >> 31
    return x.astype(np.int64), sign.astype(np.int64)


def test_feistel_cycle_walk_is_a_bijection():
    """The permutation behind the device randomised control: a 4-round Feistel network on 2h bits, cycle-walked into
    [0, C), is a bijection of the cells for every gene key -- and, over genes, behaves like independent uniform
    permutations with fair signs (what permute_rows_nsign, analysis.py:2413-2420, draws from numba's generator)."""
    for C in (1, 2, 3, 72, 1000, 4097):
        for g in (0, 1, 12345):
            x, _ = _feistel_numpy(C, g)
            assert np.array_equal(np.sort(x), np.arange(C)), (C, g)
    C, n_genes = 500, 4000
    perms, signs = zip(*(_feistel_numpy(C, g) for g in range(n_genes)))
    perms, signs = np.array(perms), np.array(signs)
    # where cell 0 / cell 137 come from: uniform over the C cells (chi-square, 499 dof: mean 499, sd 31.6)
    for c in (0, 137):
        cnt = np.bincount(perms[:, c], minlength=C)
        chi2 = ((cnt - n_genes / C) ** 2 / (n_genes / C)).sum()
        assert chi2 < 499 + 5 * 31.6, chi2
    # neighbouring cells are not mapped to neighbouring sources, fixed points are ~1/C, signs are fair and independent
    d = np.abs(perms[:, 1:] - perms[:, :-1])
    assert abs((d == 1).mean() - 2.0 / C) < 1.5e-3
    assert abs((perms == np.arange(C)[None, :]).mean() - 1.0 / C) < 5e-4
    assert abs(signs.mean() - 0.5) < 2e-3
    assert abs(np.corrcoef(signs[:, 3], signs[:, 4])[0, 1]) < 0.06 and abs(np.corrcoef(signs[7], signs[8])[0, 1]) < 0.15
    # two genes never share a permutation
    assert len({tuple(p[:16]) for p in perms}) == n_genes


def test_normalize_dispatch_matches_reference_interface(monkeypatch):
    """VelocytoLoom.normalize routes `which` to the four internal normalisers with the reference's keyword mapping
    (analysis.py:635-676): (S, U) for "both", (Sx, Ux) for "imputed", target_size[0] / [1], use_S_size_for_U ->
    use_S_size / use_Sx_size; anything else does nothing.  No GPU: the normalisers are replaced by recorders."""
    from velocyto_b200.analysis import VelocytoLoom
    calls = []
    for n in ("S", "U", "Sx", "Ux"):
        monkeypatch.setattr(VelocytoLoom, "_normalize_" + n,
                            lambda self, _n=n, **kw: calls.append((_n, kw)), raising=True)
    vlm = VelocytoLoom.__new__(VelocytoLoom)
    rel = np.arange(3.0)
    expect = {"both": ["S", "U"], "S": ["S"], "U": ["U"], "imputed": ["Sx", "Ux"], "Sx": ["Sx"], "Ux": ["Ux"], "bogus": []}
    for which, names in expect.items():
        calls.clear()
        vlm.normalize(which, size=False, log=True, pcount=0.5, relative_size=rel, use_S_size_for_U=True, target_size=(7.0, 9.0))
        assert [c[0] for c in calls] == names, which
        for n, kw in calls:
            assert kw["size"] is False and kw["log"] is True and kw["pcount"] == 0.5 and kw["relative_size"] is rel
            assert kw["target_size"] == (7.0 if n in ("S", "Sx") else 9.0)
            assert ("use_S_size" in kw) == (n == "U") and ("use_Sx_size" in kw) == (n == "Ux")
            if n in ("U", "Ux"):
                assert kw["use_S_size" if n == "U" else "use_Sx_size"] is True


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) prints ONE JSON line with the contract's
    keys; runs the reference's compiled kernel (oracle/_ref) or the oracle port on a tiny bounded sample."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "0.5", "--genes", "2000", "--neighbors", "300"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "cells/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in line, key
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_balanced_knn_routes_unserved_metrics_to_scikit_learn():
    """The device search serves Euclidean / correlation / cosine; any other scikit-learn metric must go to
    scikit-learn with THAT metric (the reference passes it through, neighbors.py:239-243) -- never silently to the
    Euclidean kernel.  Runs without a GPU because the manhattan search never touches the device."""
    from sklearn.neighbors import NearestNeighbors
    from velocyto_b200.neighbors import BalancedKNN
    rng = np.random.default_rng(5)
    X = rng.normal(size=(120, 6))
    bk = BalancedKNN(k=5, sight_k=20, maxl=10, metric="manhattan")
    assert bk._device_metric() is None
    bk.fit(X)
    dist, idx, _ = bk.kneighbors()
    d_ref, i_ref = NearestNeighbors(n_neighbors=21, metric="manhattan").fit(X).kneighbors(X)
    assert np.array_equal(bk.dsi, i_ref) and np.allclose(bk.dist, d_ref)
    d_euc, i_euc = NearestNeighbors(n_neighbors=21).fit(X).kneighbors(X)
    assert not np.array_equal(bk.dsi, i_euc)                            # and it is NOT the Euclidean answer
    assert BalancedKNN(metric="euclidean")._device_metric() == "euclidean"
    assert BalancedKNN(metric="correlation")._device_metric() == "correlation"
    assert BalancedKNN(metric="euclidean", search="host")._device_metric() is None


def test_sharded_host_front_helpers():
    from velocyto_b200.sharding import _host_ptr, needs_residuals
    a = np.zeros((7, 40))
    ptr, pitch, esz = _host_ptr(a[:, 8:20])                             # a cell block of the gene-major matrix: no copy
    assert ptr == a.ctypes.data + 8 * 8 and pitch == 40 and esz == 8
    import torch
    t = torch.zeros((5, 12), dtype=torch.float32)
    assert _host_ptr(t) == (t.data_ptr(), 12, 4)
    with pytest.raises(AssertionError):
        _host_ptr(a.T)                                                  # genes must stay the slow axis
    # same rule as the one-GPU host tier (capi.cu): residuals only for fp64 data under a transform that jumps at 0
    assert needs_residuals("sqrt", 1.0, 8) and not needs_residuals("sqrt", 1e-10, 8)
    assert not needs_residuals("sqrt", 1.0, 4) and not needs_residuals("linear", 0.0, 8)
    assert needs_residuals("log10", 0.5, 8) and not needs_residuals("log10", 1.0, 8)


def test_host_sampler_consumes_numpys_legacy_stream_bit_for_bit():
    """The C++ restatement of the per-cell ``np.random.choice(..., replace=False, p=p)`` loop (analysis.py:1561-1564):
    identical samples AND identical generator state afterwards, for the reference's linear probability ramps, a steep
    ramp with many collisions, and the degenerate size == W case.  No GPU involved."""
    from velocyto_b200.analysis import _sample_neighbors_numpy_stream
    for C, W, frac, ramp in ((40, 31, 0.5, (0.5, 0.1)), (60, 2001, 0.3, (0.5, 0.1)), (6, 20001, 0.3, (0.5, 0.1)),
                             (9, 5, 1.0, (0.5, 0.1)), (25, 501, 0.9, (0.9, 0.001))):
        p = np.linspace(ramp[0], ramp[1], W)
        p = p / p.sum()
        size = int(frac * W)
        np.random.seed(15071990)
        want = np.stack([np.random.choice(W, size=(size,), replace=False, p=p) for _ in range(C)], 0)
        state_want = np.random.get_state()
        np.random.seed(1)                                              # whatever was there before must not matter
        got = _sample_neighbors_numpy_stream(15071990, C, W, p, size)
        state_got = np.random.get_state()
        assert np.array_equal(got, want), (C, W)
        assert np.array_equal(state_got[1], state_want[1]) and state_got[2] == state_want[2]
        assert np.random.random() == (np.random.set_state(state_want) or np.random.random())   # and the NEXT draw agrees


def test_connectivity_fast_path_equals_general_path():
    """Regular kNN graphs (k entries per row, what the searches return) take an (n, k)-block shortcut in
    connectivity_with_diagonal; zero distances or self edges fall back to the per-edge path.  Both must equal
    ``(knn > 0).astype(float)`` + ``setdiag(diag)`` (analysis.py:1006-1009)."""
    import warnings
    from velocyto_b200.analysis import connectivity_with_diagonal
    rng = np.random.default_rng(0)
    n, k = 300, 9
    idx = np.stack([rng.choice(np.delete(np.arange(n), c), k, replace=False) for c in range(n)])
    dat = rng.uniform(0.1, 1.0, (n, k))

    def ref(knn, diag):
        conn = (knn > 0).astype(float)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            conn.setdiag(diag)
        return conn

    regular = sparse.csr_matrix((dat.ravel(), idx.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    dat0 = dat.copy()
    dat0[5, 2] = 0.0                                                   # a zero-distance edge: dropped by `knn > 0`
    with_zero = sparse.csr_matrix((dat0.ravel(), idx.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    idx_self = idx.copy()
    idx_self[7, 0] = 7                                                 # a self edge: overwritten by setdiag
    with_self = sparse.csr_matrix((dat.ravel(), idx_self.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    ragged = sparse.csr_matrix(regular.toarray() * (rng.uniform(size=(n, n)) < 0.7))
    for knn in (regular, with_zero, with_self, ragged):
        for diag in (1, 8):
            assert (connectivity_with_diagonal(knn, diag) != ref(knn, diag)).nnz == 0
    # (SciPy's `knn > 0` in ref() canonicalised `regular` in place -- the reference does the same to self.knn -- so the
    # order check uses a fresh matrix)
    fresh = sparse.csr_matrix((dat.ravel(), idx.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    fast = connectivity_with_diagonal(fresh, 1)
    assert np.array_equal(fast.indices.reshape(n, k + 1)[:, 0], np.arange(n))      # diagonal first, kNN order kept
    assert np.array_equal(fast.indices.reshape(n, k + 1)[:, 1:], idx)


def test_velocytoloom_accepts_sparse_layers_and_lazy_dense_attributes():
    """Host-side behaviour that needs no GPU: sparse count layers are kept sparse (no dense float64 copy), the
    by-cell triplet constructor builds the same object, and the dense (cells x cells) result attributes are lazy
    descriptors with ordinary set / get / delete semantics."""
    from velocyto_b200.analysis import VelocytoLoom
    rng = np.random.default_rng(1)
    S = rng.poisson(0.3, (40, 25)).astype(np.float64)
    U = rng.poisson(0.2, (40, 25)).astype(np.float64)
    dense = VelocytoLoom(S=S, U=U)
    sp = VelocytoLoom(S=sparse.csr_matrix(S), U=sparse.coo_matrix(U))
    assert sparse.issparse(sp.S) and sparse.issparse(sp.U) and sp.S.shape == (40, 25)
    np.testing.assert_array_equal(sp.initial_cell_size, dense.initial_cell_size)
    np.testing.assert_array_equal(sp.initial_Ucell_size, dense.initial_Ucell_size)
    cs, cu = sparse.csc_matrix(S), sparse.csc_matrix(U)
    tri = VelocytoLoom.from_csr((cs.data, cs.indices, cs.indptr), (cu.data, cu.indices, cu.indptr), n_genes=40)
    assert (tri.S != sp.S).nnz == 0 and (tri.U != sp.U).nnz == 0
    with pytest.raises(AttributeError):
        dense.corrcoef                                                  # nothing computed yet
    assert not hasattr(dense, "transition_prob")
    dense.corrcoef = np.eye(3)                                          # a user assignment is kept as it is
    assert np.array_equal(dense.corrcoef, np.eye(3))
    del dense.corrcoef
    assert not hasattr(dense, "corrcoef")


def test_smoothing_weights_from_knn_equals_reference_sequence():
    """knn_imputation's graph preparation in one step (smoothing_weights_from_knn) == the reference's sequence
    ``(knn > 0).astype(float)``, ``setdiag(diag)``, ``connectivity_to_weights`` (analysis.py:1006-1010) -- closed form for
    regular kNN graphs, general path for zero distances, self edges and ragged rows."""
    import warnings
    from velocyto_b200.analysis import smoothing_weights_from_knn
    from velocyto_b200.neighbors import connectivity_to_weights
    rng = np.random.default_rng(3)
    n, k = 180, 11
    idx = np.stack([rng.choice(np.delete(np.arange(n), c), k, replace=False) for c in range(n)])
    dist = rng.uniform(0.1, 1.0, (n, k))
    mk = lambda d, i: sparse.csr_matrix((d.ravel(), i.astype(np.int32).ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))

    def ref(knn, diag):
        conn = (knn > 0).astype(float)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            conn.setdiag(diag)
        return sparse.csr_matrix(connectivity_to_weights(conn))

    d0 = dist.copy()
    d0[3, 2] = 0.0
    i_self = idx.copy()
    i_self[9, 0] = 9
    ragged = sparse.csr_matrix(mk(dist, idx).toarray() * (rng.uniform(size=(n, n)) < 0.6))
    for knn_fn in (lambda: mk(dist, idx), lambda: mk(d0, idx), lambda: mk(dist, i_self), lambda: ragged.copy()):
        for diag in (1, 8, 0.5):
            got, want = smoothing_weights_from_knn(knn_fn(), diag), ref(knn_fn(), diag)
            assert abs(got - want).max() < 1e-15
            np.testing.assert_allclose(np.asarray(got.sum(1)).ravel(), 1.0, rtol=1e-13)
    w = smoothing_weights_from_knn(mk(dist, idx), 1)                     # regular graph: diagonal first, kNN order kept
    assert np.array_equal(w.indices.reshape(n, k + 1)[:, 0], np.arange(n))
    assert np.array_equal(w.indices.reshape(n, k + 1)[:, 1:], idx)
