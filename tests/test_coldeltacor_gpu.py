"""GPU parity of the colDeltaCor kernels against the oracle and the reference golden vectors.

Everything goes through the C ABI (host tier via ``velocyto_b200.estimation``, device tier via
``velocyto_b200.device``).  Tolerances: the north-star bar is 1e-5 relative on transition
probabilities; ``p ~ exp(corr / 0.05)`` amplifies an absolute correlation error 20x, so the
correlations themselves are held to 5e-7 absolute (5e-6 on the tiny golden case with 37 genes,
where single fp32 roundings are not averaged out).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("sqrt", 1e-10), ("sqrt", 1.0), ("log10", 1.0), ("log10", 0.5), ("linear", 0.0)]


def _d(z, name, psc):
    if name == "sqrt":
        return np.sqrt(np.abs(z) + psc) * np.sign(z)
    if name == "log10":
        return np.log10(np.abs(z) + psc) * np.sign(z)
    return z


def _api(name, partial):
    import velocyto_b200.estimation as est
    return {("linear", False): est.colDeltaCor, ("linear", True): est.colDeltaCorpartial,
            ("sqrt", False): est.colDeltaCorSqrt, ("sqrt", True): est.colDeltaCorSqrtpartial,
            ("log10", False): est.colDeltaCorLog10, ("log10", True): est.colDeltaCorLog10partial}[(name, partial)]


def _call(name, partial, e, d, ixs, psc, **kw):
    fn = _api(name, partial)
    args = (e, d, ixs) if partial else (e, d)
    if name != "linear":
        kw["psc"] = psc
    return fn(*args, **kw)


def _degenerate_pairs(e):
    C = e.shape[1]
    return np.array([[np.array_equal(e[:, i], e[:, c]) for i in range(C)] for c in range(C)])


def synth(G, C, seed, name="sqrt", psc=1e-10, fp32_exact=True):
    """Seeded inputs.  With ``fp32_exact`` the values are exactly representable in float32 (still handed over
    as float64, as the reference API demands).  Why: for psc > 0 the reference's transforms are DISCONTINUOUS at
    zero difference (sign(t)*sqrt(|t|+psc) jumps by 2*sqrt(psc); speedboosted.pyx:372-378), so the sign of a
    difference of two values closer than one fp32 ulp -- which the fp32 device layout cannot see -- changes
    that element by O(1).  Identical-input parity is therefore stated on fp32-representable inputs; the default
    psc of the sqrt transform (1e-10, analysis.py:1523-1524) and log10 with psc=1 are continuous and are also
    tested on raw float64 inputs."""
    rng = np.random.default_rng(seed)
    e = rng.gamma(2.0, 1.0, (G, C))
    e[rng.uniform(size=(G, C)) < 0.3] = 0.0          # realistic sparsity -> exact zero differences
    d = _d(rng.normal(size=(G, C)), name, psc)
    if fp32_exact:
        e, d = e.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64)
    return e, d


def rand_ixs(C, m, seed):
    """m distinct neighbours per cell, never the cell itself (a self pair is a zero-variance column:
    NaN or rounding garbage in the reference depending on the variant)."""
    rng = np.random.default_rng(seed)
    out = np.empty((C, m), dtype=np.int64)
    for c in range(C):
        pick = rng.choice(C - 1, m, replace=False)
        out[c] = pick + (pick >= c)
    return out


@pytest.mark.parametrize("name,psc", CASES)
def test_golden_small_all_variants(golden, name, psc):
    g = golden("coldeltacor_small.npz")
    e, z, ixs = g["e"], g["z"], g["ixs"]
    d = _d(z, name, psc)
    tag = f"{name}_{psc:g}"
    ok = ~_degenerate_pairs(e)
    full = _call(name, False, e, d, None, psc)
    part = _call(name, True, e, d, ixs, psc)
    assert full.shape == part.shape == (e.shape[1],) * 2 and full.dtype == np.float64
    for got, want in ((full, g[f"full_{tag}"]), (part, g[f"partial_{tag}"])):
        assert not np.isnan(got[ok]).any()
        # 37 genes: fp32 roundings are not averaged out, and psc=1 full variants carry a large common
        # mean (t == 0 -> -sqrt(psc)); the fp32 one-pass sums are good to ~2.5e-6 here, ~1e-8 at G >= 1000
        np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=5e-6)
    sampled = np.zeros_like(ok)
    sampled[np.arange(ixs.shape[0])[:, None], ixs] = True
    assert np.all(part[~sampled] == 0)                # only sampled entries are touched
    if name == "sqrt":                                # zero rule -> NaN on identical columns, as the reference
        deg = ~ok & sampled
        assert np.array_equal(np.isnan(part[deg]), np.isnan(g[f"partial_{tag}"][deg]))


@pytest.mark.parametrize("name,psc", CASES)
def test_partial_matches_oracle_medium(oracle, name, psc):
    G, C, m = 2000, 600, 96
    e, d = synth(G, C, 1, name, psc)
    ixs = rand_ixs(C, m, 2)
    want = oracle.coldeltacor(e, d, ixs, name, psc)
    got = _call(name, True, e, d, ixs, psc)
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-7)
    comp = _call(name, True, e, d, ixs, psc, compact=True)
    assert comp.shape == (C, m) and comp.dtype == np.float32
    np.testing.assert_allclose(comp, want[np.arange(C)[:, None], ixs], rtol=0, atol=5e-7)


@pytest.mark.parametrize("name,psc", CASES)
def test_full_matches_oracle_medium(oracle, name, psc):
    G, C = 1500, 257
    e, d = synth(G, C, 3, name, psc)
    want = oracle.coldeltacor(e, d, None, name, psc)
    got = _call(name, False, e, d, None, psc)
    off = ~np.eye(C, dtype=bool)                      # the diagonal is degenerate in every full variant
    np.testing.assert_allclose(got[off], want[off], rtol=0, atol=5e-7)


@pytest.mark.parametrize("name,psc", [("sqrt", 1e-10), ("log10", 1.0), ("linear", 0.0)])
def test_continuous_transforms_on_raw_float64_inputs(oracle, name, psc):
    """Inputs NOT representable in fp32: the transforms that are continuous at zero difference keep the
    5e-7 bound through the fp32 device layout."""
    G, C, m = 3000, 300, 64
    e, d = synth(G, C, 13, name, psc, fp32_exact=False)
    ixs = rand_ixs(C, m, 14)
    want = oracle.coldeltacor(e, d, ixs, name, psc)
    got = _call(name, True, e, d, ixs, psc)
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-7)


@pytest.mark.parametrize("name,psc", [("sqrt", 1.0), ("log10", 0.5)])
def test_discontinuous_transforms_on_raw_float64_inputs(oracle, name, psc):
    """Raw float64 inputs with the transforms that JUMP at zero difference: the host tier keeps the fp32
    residuals of e and resolves fp32 ties to the sign the fp64 reference sees, so the 5e-7 bound still holds."""
    G, C, m = 3000, 300, 64
    e, d = synth(G, C, 15, name, psc, fp32_exact=False)
    ixs = rand_ixs(C, m, 16)
    want = oracle.coldeltacor(e, d, ixs, name, psc)
    got = _call(name, True, e, d, ixs, psc)
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-7)
    off = ~np.eye(C, dtype=bool)
    np.testing.assert_allclose(_call(name, False, e, d, None, psc)[off], oracle.coldeltacor(e, d, None, name, psc)[off],
                               rtol=0, atol=5e-7)


def test_fp32_ties_resolved_like_fp64_reference(oracle):
    """Adversarial: every cell's value of a gene rounds to the SAME float32 (relative spread 1e-10), so all fp32
    differences are exactly 0 while the fp64 reference sees definite signs -> +-sqrt(psc) per element."""
    from velocyto_b200 import device as dev
    G, C, m, psc = 257, 96, 17, 1.0
    rng = np.random.default_rng(17)
    base = rng.gamma(2.0, 1.0, (G, 1)).astype(np.float32).astype(np.float64)
    e = base * (1.0 + 1e-10 * rng.integers(-50, 50, (G, C)))
    e[::5] = 0.0                                                   # exact zero rows stay exact ties
    assert np.all(e.astype(np.float32) == e.astype(np.float32)[:, :1])
    d = _d(rng.normal(size=(G, C)), "sqrt", psc)
    ixs = rand_ixs(C, m, 18)
    want = oracle.coldeltacor(e, d, ixs, "sqrt", psc)
    got = _call("sqrt", True, e, d, ixs, psc)                      # host tier: residuals kept automatically
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-7, equal_nan=True)
    e_cm = dev.CellMajor.from_gene_major(e, residual=True)          # device tier
    assert e_cm.lo is not None
    ix = dev.indices_to_device(ixs, C)
    comp = dev.coldeltacor(e_cm, dev.CellMajor.from_gene_major(d), ix, "sqrt", psc).cpu().numpy()
    np.testing.assert_allclose(comp, want[np.arange(C)[:, None], ixs], rtol=0, atol=5e-7, equal_nan=True)
    # fp32-representable data carries no residual matrix
    assert dev.CellMajor.from_gene_major(e.astype(np.float32).astype(np.float64), residual=True).lo is None


@pytest.mark.parametrize("G", [30001, 61003, 129, 130])
def test_ragged_and_multislab_gene_axis(oracle, G):
    """G not a multiple of 4/128, and G large enough for 2 and 3 shared-memory slabs."""
    C, m = 24, 7
    e, d = synth(G, C, 4, "sqrt", 1.0)
    ixs = rand_ixs(C, m, 5)
    want = oracle.coldeltacor(e, d, ixs, "sqrt", 1.0)
    got = _call("sqrt", True, e, d, ixs, 1.0)
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-7)


def test_more_neighbours_than_one_chunk(oracle):
    """m > 4096 exercises the neighbour-chunk loop of the kernel."""
    G, C, m = 48, 4300, 4200
    e, d = synth(G, C, 6, "linear", 0.0)
    rng = np.random.default_rng(7)
    ixs = np.stack([rng.permutation(C)[:m] for _ in range(C)])
    want = oracle.coldeltacor(e, d, ixs, "linear", 0.0)
    got = _call("linear", True, e, d, ixs, 0.0, compact=True)
    ref = want[np.arange(C)[:, None], ixs]
    ok = ixs != np.arange(C)[:, None]
    np.testing.assert_allclose(got[ok], ref[ok], rtol=0, atol=5e-6)


def test_duplicate_indices_accumulate(oracle):
    """`rm[c, i] +=` per sampled slot: a repeated index counts twice (speedboosted.pyx:336)."""
    G, C, m = 300, 40, 6
    e, d = synth(G, C, 8, "sqrt", 1.0)
    ixs = rand_ixs(C, m, 9)
    ixs[:, 1] = ixs[:, 0]
    want = oracle.coldeltacor(e, d, ixs, "sqrt", 1.0)
    got = _call("sqrt", True, e, d, ixs, 1.0)
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(want), rtol=0, atol=1e-6)


def test_reference_argument_contract():
    """dmat must already be C-contiguous float64; emat / ixs are coerced (estimation.py:59-60)."""
    import velocyto_b200.estimation as est
    G, C, m = 64, 16, 4
    e, d = synth(G, C, 10)
    ixs = rand_ixs(C, m, 0)
    with pytest.raises(ValueError):
        est.colDeltaCorSqrtpartial(e, np.asfortranarray(d), ixs, psc=1.0)
    with pytest.raises(ValueError):
        est.colDeltaCorSqrtpartial(e, d.astype(np.float32), ixs, psc=1.0)
    a = est.colDeltaCorSqrtpartial(np.asfortranarray(e), d, ixs.astype(np.int32), psc=1.0)
    b = est.colDeltaCorSqrtpartial(e, d, ixs, threads=3, psc=1.0)
    assert np.array_equal(a, b)
    from velocyto_b200 import VeloError
    bad = ixs.copy()
    bad[0, 0] = C
    with pytest.raises(VeloError):
        est.colDeltaCorSqrtpartial(e, d, bad, psc=1.0)


def test_device_tier_sharded_rows_match_whole(oracle):
    """Cell-sharded call (c0, nc) == the same rows of the whole-problem call: the multi-GPU contract."""
    import torch
    from velocyto_b200 import device as dev
    G, C, m = 700, 300, 33
    e, d = synth(G, C, 11, "sqrt", 1e-10)
    ixs = rand_ixs(C, m, 12)
    e_cm, d_cm = dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d)
    ix = dev.indices_to_device(ixs, C)
    whole = dev.coldeltacor(e_cm, d_cm, ix, "sqrt", 1e-10)
    c0, nc = 117, 90
    part = dev.coldeltacor(e_cm, d_cm.rows(c0, nc), ix[c0:c0 + nc].contiguous(), "sqrt", 1e-10, c0=c0)
    assert torch.equal(whole[c0:c0 + nc], part)       # bit-identical: same kernel, same order
    want = oracle.coldeltacor(e, d, ixs, "sqrt", 1e-10)[np.arange(C)[:, None], ixs]
    np.testing.assert_allclose(whole.cpu().numpy(), want, rtol=0, atol=5e-7)
    # round trip of the layout converters
    np.testing.assert_allclose(e_cm.to_gene_major(), e.astype(np.float32).astype(np.float64), rtol=0, atol=0)


def test_transition_prob_matches_reference_golden(golden):
    """corrcoef -> transition_prob within 1e-5 relative of the reference (analysis.py:1604-1612,1697-1698)."""
    import torch
    from velocyto_b200 import device as dev
    g = golden("pipeline_small.npz")
    Sx = g["Sx_sz"]
    C = Sx.shape[1]
    neigh = g["neigh_ixs"].astype(np.int64)
    psc = 1.0
    e_cm = dev.CellMajor.from_gene_major(Sx)
    ix = dev.indices_to_device(neigh, C)
    for dkey, pkey in (("delta_S", "transition_prob"), ("delta_S_rndm", "transition_prob_random")):
        delta = (Sx + 1.0 * g[dkey]) - Sx
        dmat = np.sqrt(np.abs(delta) + psc) * np.sign(delta)
        corr = dev.coldeltacor(e_cm, dev.CellMajor.from_gene_major(dmat), ix, "sqrt", psc)
        tp = dev.transition_prob(corr, ix, 0.05).cpu().numpy().astype(np.float64)
        want = g[pkey][np.arange(C)[:, None], neigh]
        np.testing.assert_allclose(tp, want, rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(tp.sum(1), 1.0, rtol=1e-6)


def test_full_size_properties_and_spot_check(oracle):
    """BASELINE config-4 shape (100k cells x 30k genes) at reduced m: size-independent properties
    (|corr| <= 1, sign flip under d -> -d, invariance under d -> 3d, determinism) plus an oracle
    spot check of a few cells on the gathered sub-problem."""
    import torch
    from velocyto_b200 import device as dev
    free, _ = torch.cuda.mem_get_info()
    C, G, m = (100_000, 30_000, 24) if free > 60e9 else (20_000, 8_000, 24)
    gen = torch.Generator(device="cuda").manual_seed(0)
    e_cm = dev.CellMajor.empty(C, G)
    d_cm = dev.CellMajor.empty(C, G)
    blk = 10_000
    for c0 in range(0, C, blk):
        u = torch.rand((min(blk, C - c0), G), device="cuda", generator=gen)
        v = -torch.log(torch.rand_like(u).clamp_min(1e-7)) - torch.log(u.clamp_min(1e-7))   # Gamma(2,1)
        v[torch.rand_like(u) < 0.3] = 0
        e_cm.t[c0:c0 + v.shape[0], :G] = v
        z = torch.randn((v.shape[0], G), device="cuda", generator=gen)
        d_cm.t[c0:c0 + v.shape[0], :G] = torch.sign(z) * torch.sqrt(z.abs() + 1.0)
    # m neighbours per cell, never the cell itself (a self pair is degenerate: NaN, as in the reference)
    ix = ((torch.arange(C, device="cuda")[:, None] + 1 +
           torch.randint(0, C - 1, (C, m), device="cuda", generator=gen)) % C).to(torch.int32).contiguous()
    out = dev.coldeltacor(e_cm, d_cm, ix, "sqrt", 1.0)
    assert torch.isfinite(out).all() and float(out.abs().max()) <= 1.0 + 1e-6
    assert torch.equal(out, dev.coldeltacor(e_cm, d_cm, ix, "sqrt", 1.0))                   # deterministic
    d_cm.t.mul_(-3.0)
    flipped = dev.coldeltacor(e_cm, d_cm, ix, "sqrt", 1.0)
    assert float((flipped + out).abs().max()) < 2e-6
    d_cm.t.div_(-3.0)
    # spot check: cells and their neighbours gathered into a small dense problem for the oracle
    for c in (0, C // 3, C - 1):
        nb = ix[c].cpu().numpy().astype(np.int64)
        cols = np.concatenate([[c], nb])
        e_sub = e_cm.t[torch.from_numpy(cols).cuda(), :G].cpu().numpy().astype(np.float64).T
        d_sub = np.zeros_like(e_sub)
        d_sub[:, 0] = d_cm.t[c, :G].cpu().numpy()
        sub_ix = np.zeros((cols.size, m), dtype=np.int64)
        sub_ix[0] = np.arange(1, m + 1)
        d_sub[:, 1:] = 1.0 + np.arange(G)[:, None]     # any non-degenerate filler
        want = oracle.coldeltacor(e_sub, d_sub, sub_ix, "sqrt", 1.0)[0, 1:]
        np.testing.assert_allclose(out[c].cpu().numpy(), want, rtol=0, atol=5e-7)


def test_empty_and_degenerate_shapes():
    """No sampled neighbours, a single gene, a single cell: the wrappers return what the reference's loops leave behind
    (untouched zeros; NaN for zero-variance pairs) instead of failing."""
    import velocyto_b200.estimation as est
    from velocyto_b200 import VeloError
    e, d = synth(50, 12, 20)
    out = est.colDeltaCorSqrtpartial(e, d, np.empty((12, 0), dtype=np.int64), psc=1.0)          # m == 0
    assert out.shape == (12, 12) and not out.any()
    comp = est.colDeltaCorSqrtpartial(e, d, np.empty((12, 0), dtype=np.int64), psc=1.0, compact=True)
    assert comp.shape == (12, 0)
    one_gene = est.colDeltaCorpartial(e[:1].copy(), d[:1].copy(), rand_ixs(12, 3, 21))               # G == 1: no variance
    assert np.isnan(one_gene[np.arange(12)[:, None], rand_ixs(12, 3, 21)]).all()
    one_cell = est.colDeltaCorSqrt(e[:, :1].copy(), d[:, :1].copy(), psc=1.0)                        # C == 1: only the self pair
    assert one_cell.shape == (1, 1)
    with pytest.raises(VeloError):
        est.colDeltaCor(np.empty((0, 5)), np.empty((0, 5)))                                          # no genes at all


# ----------------------------------------------------------------------------------------------------------
# K2g: the all-pairs linear variant on the tensor cores (csrc/coldeltacor_tc.cu)

def _tc_inputs(G, C, seed):
    """Genes with very different expression levels (the common profile must cancel in e_i - e_c) and a velocity
    field correlated with the expression differences, so that correlations are not all ~0."""
    rng = np.random.default_rng(seed)
    level = rng.gamma(0.6, 2.0, G)[:, None] + 0.05
    e = rng.gamma(2.0, 1.0, (G, C)) * level
    e[rng.uniform(size=(G, C)) < 0.3] = 0.0
    z = rng.normal(size=(G, C))
    d = z + 0.5 * (e - e.mean(1, keepdims=True)) / (e.std(1, keepdims=True) + 1e-9)
    return e.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("G,C,c0,nc", [(3000, 700, 0, 700), (777, 333, 100, 200), (64, 128, 0, 128), (5, 9, 2, 4),
                                       (4097, 130, 1, 129)])
def test_tensor_core_linear_matches_oracle(oracle, G, C, c0, nc):
    """x_colDeltaCor (speedboosted.pyx:13-87) via tcgen05: ragged gene blocks (G % 64 != 0), ragged tiles
    (C % 128 != 0), a cell shard (c0, nc), and tiny shapes; 5e-7 absolute as for the other kernels."""
    from velocyto_b200 import device as dev
    e, d = _tc_inputs(G, C, G + C)
    want = oracle.coldeltacor(e, d, None, "linear", 0.0)[c0:c0 + nc]
    E, D = dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d[:, c0:c0 + nc])
    got = dev.coldeltacor_linear_tc(E, D, c0=c0).cpu().numpy()
    self_pair = np.zeros((nc, C), dtype=bool)
    self_pair[np.arange(nc), c0 + np.arange(nc)] = True
    assert np.isnan(got[self_pair]).all()                       # 0 * inf in the reference (probed: NaN)
    assert not np.isnan(got[~self_pair]).any()
    np.testing.assert_allclose(got[~self_pair], want[~self_pair], rtol=0, atol=5e-7)


def test_tensor_core_linear_is_the_default_full_linear_path(oracle):
    """estimation.colDeltaCor (host tier, dense fp64 output) runs K2g; switching the tensor cores off gives the fp32
    kernel K2; both within tolerance of the oracle, K2g launches its own kernels."""
    from velocyto_b200 import _cabi
    import velocyto_b200.estimation as est
    G, C = 2500, 300
    e, d = _tc_inputs(G, C, 5)
    want = oracle.coldeltacor(e, d, None, "linear", 0.0)
    off = ~np.eye(C, dtype=bool)
    lib = _cabi.load()
    assert lib.velo_get_tensor_cores() == 1
    got_tc = est.colDeltaCor(e, d)
    try:
        lib.velo_set_tensor_cores(0)
        got_k2 = est.colDeltaCor(e, d)
    finally:
        lib.velo_set_tensor_cores(1)
    np.testing.assert_allclose(got_tc[off], want[off], rtol=0, atol=5e-7)
    # K2 keeps one fp32 FMA chain per sum over all genes: on expression levels this heavy-tailed it reaches ~1e-6
    np.testing.assert_allclose(got_k2[off], want[off], rtol=0, atol=2e-6)
    assert not np.array_equal(got_tc[off], got_k2[off])         # really two different kernels
    assert np.abs(got_tc - want)[off].max() <= np.abs(got_k2 - want)[off].max()   # split-fp16 + block sums beat fp32 FMA chains


def test_tensor_core_linear_degenerate_cells(oracle):
    """Coincident cells (A == 0 everywhere) and constant velocity rows are 0 * inf = NaN in the reference."""
    from velocyto_b200 import device as dev
    G, C = 500, 140
    e, d = _tc_inputs(G, C, 9)
    e[:, 17] = e[:, 3]                                          # two identical cells
    d[:, 50] = 0.25                                             # zero-variance velocity
    want = oracle.coldeltacor(e, d, None, "linear", 0.0)
    got = dev.coldeltacor_linear_tc(dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d)).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=5e-7)


def test_tensor_core_linear_transition_probabilities(oracle):
    """The north-star bar: transition probabilities within 1e-5 relative of the reference path at 8000 genes."""
    from velocyto_b200 import device as dev
    G, C = 8000, 256
    e, d = _tc_inputs(G, C, 21)
    corr_want = oracle.patch_corrcoef(oracle.coldeltacor(e, d, None, "linear", 0.0))
    tp_want = oracle.transition_prob(corr_want, np.ones((C, C)), 0.05)   # full mode: the mask keeps the self term
    corr = dev.coldeltacor_linear_tc(dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d))
    tp = dev.transition_prob(corr, None, 0.05).cpu().numpy()
    np.testing.assert_allclose(tp, tp_want, rtol=1e-5, atol=0)


def test_host_tier_pipelined_chunks_match_single_pass(oracle, monkeypatch):
    """The host tier moves d / ixs / out in cell chunks underneath the kernel of the previous chunk (capi.cu);
    forcing 7 ragged chunks on a small problem must reproduce the one-chunk result bit for bit, and a bad index
    in a LATER chunk must still be rejected before its kernel runs."""
    import velocyto_b200.estimation as est
    from velocyto_b200 import _cabi
    G, C, m = 900, 500, 40
    e, d = synth(G, C, 31, "sqrt", 1.0)
    ixs = rand_ixs(C, m, 32)
    monkeypatch.delenv("VELO_HOST_CHUNK_CELLS", raising=False)
    one = est.colDeltaCorSqrtpartial(e, d, ixs, psc=1.0, compact=True)
    monkeypatch.setenv("VELO_HOST_CHUNK_CELLS", "77")
    many = est.colDeltaCorSqrtpartial(e, d, ixs, psc=1.0, compact=True)
    assert np.array_equal(one, many, equal_nan=True)
    want = oracle.coldeltacor(e, d, ixs, "sqrt", 1.0)[np.arange(C)[:, None], ixs]
    np.testing.assert_allclose(many, want, rtol=0, atol=5e-7)
    bad = ixs.copy()
    bad[C - 3, 5] = C + 7
    with pytest.raises((_cabi.VeloError, ValueError)):
        est.colDeltaCorSqrtpartial(e, d, bad, psc=1.0, compact=True)


# ----------------------------------------------------------------------------------------------------------
# round 2: direct parity at the benchmark shape, the tie-resolving variant at scale, the staged / sharded host tier

def _gather_subproblem(e_cm, lo, d_cm, ix, cells, G):
    """Selected cells + all their neighbours as a small fp64 problem for ``oracle.coldeltacor_cells``."""
    import torch
    cols, sub_ix, pos = [], [], 0
    for c in cells:
        nb = ix[c].to(torch.int64)
        cols.append(torch.cat([torch.tensor([c], device=nb.device), nb]))
        sub_ix.append(np.arange(pos + 1, pos + 1 + nb.numel()))
        pos += 1 + nb.numel()
    cols = torch.cat(cols)
    e_sub = e_cm.t[cols, :G].to(torch.float64)
    if lo is not None:
        e_sub = e_sub + lo[cols, :G].to(torch.float64)             # the fp64 values the reference would be given
    sel = np.array([s[0] - 1 for s in sub_ix])
    d_sel = torch.stack([d_cm.t[c, :G] for c in cells], 1).to(torch.float64).cpu().numpy()
    return np.ascontiguousarray(e_sub.cpu().numpy().T), d_sel, sel, np.stack(sub_ix)


@pytest.mark.parametrize("exact", [False, True])
def test_bench_shape_rows_match_oracle(oracle, exact):
    """BASELINE config 4 exactly as bench.py runs it -- 100k cells x 30k genes, m = 3000 neighbours, sqrt, psc = 1 (two
    gene slabs, the in-kernel sort, 3000 accumulators) -- for a handful of cells: correlations within 5e-7 and
    transition probabilities within 1e-5 of the oracle.  ``exact``: the expression matrix is genuinely fp64 (quantised
    fp32 values + non-zero fp32 residuals, so fp32 ties between different cells are COMMON) and the tie-resolving
    kernel variant k_coldeltacor<.,.,EXACT> runs at full scale."""
    import torch
    from velocyto_b200 import device as dev
    free, _ = torch.cuda.mem_get_info()
    C, G, m = (100_000, 30_000, 3_000) if free > 60e9 else (20_000, 8_000, 3_000)
    psc, sigma = 1.0, 0.05
    gen = torch.Generator(device="cuda").manual_seed(5 + exact)
    e_cm, d_cm = dev.CellMajor.empty(C, G), dev.CellMajor.empty(C, G)
    if exact:
        e_cm.lo = torch.zeros_like(e_cm.t)
    blk = 10_000
    for c0 in range(0, C, blk):
        n = min(blk, C - c0)
        u = torch.rand((n, G), device="cuda", generator=gen)
        v = -torch.log(torch.rand_like(u).clamp_min(1e-7)) - torch.log(u.clamp_min(1e-7))      # Gamma(2,1)
        v[torch.rand_like(u) < 0.3] = 0
        if exact:
            v = torch.round(v * 4) / 4                                # coarse grid: ties between cells everywhere
            lo = (torch.rand((n, G), device="cuda", generator=gen) - 0.5) * 2.0 ** -27 * v     # |lo| < half an ulp of v
            e_cm.lo[c0:c0 + n, :G] = lo
        e_cm.t[c0:c0 + n, :G] = v
        z = torch.randn((n, G), device="cuda", generator=gen)
        d_cm.t[c0:c0 + n, :G] = torch.sign(z) * torch.sqrt(z.abs() + psc)
    cells = [11, C // 2, C - 1]
    ix = ((torch.arange(C, device="cuda")[:, None] + 1 +
           torch.randint(0, C - 1, (C, m), device="cuda", generator=gen)) % C).to(torch.int32)
    rows = torch.tensor(cells, device="cuda")
    got = torch.cat([dev.coldeltacor(e_cm, d_cm.rows(c, 1), ix[c:c + 1].contiguous(), "sqrt", psc, c0=c) for c in cells])
    tp = dev.transition_prob(got.clone(), ix[rows].contiguous(), sigma, c0=C + 5).cpu().numpy()   # c0 past the end: no row is "self"
    e_sub, d_sel, sel, sub_ix = _gather_subproblem(e_cm, e_cm.lo, d_cm, ix, cells, G)
    want = oracle.coldeltacor_cells(e_sub, d_sel, sel, sub_ix, "sqrt", psc)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=5e-7)
    np.testing.assert_allclose(tp, oracle.transition_prob_compact(want, None, sigma), rtol=1e-5, atol=0)
    if exact:
        # the residuals matter: dropping them (what fp32 storage alone would see) moves the correlations far past the bar
        plain = dev.CellMajor(e_cm.t, G)
        off = dev.coldeltacor(plain, d_cm.rows(cells[0], 1), ix[cells[0]:cells[0] + 1].contiguous(), "sqrt", psc, c0=cells[0])
        assert float((off - got[:1]).abs().max()) > 1e-4


def _host_problem(G, C, m, seed):
    e, d = synth(G, C, seed, "sqrt", 1.0, fp32_exact=False)          # raw float64: the EXACT path with residuals
    return e, d, rand_ixs(C, m, seed + 1)


def test_host_tier_pageable_equals_pinned(oracle, monkeypatch):
    """NumPy (pageable) buffers go through the library's pinned staging ring with several host threads; page-locked
    buffers are copied directly.  Same bits either way, single- and multi-chunk, and both match the oracle."""
    import torch
    from velocyto_b200 import _cabi
    G, C, m, psc, sigma = 1100, 640, 48, 1.0, 0.05
    e, d, ixs = _host_problem(G, C, m, 41)
    want_c = oracle.coldeltacor(e, d, ixs, "sqrt", psc)[np.arange(C)[:, None], ixs]
    want = oracle.transition_prob_compact(want_c, None, sigma)
    res = {}
    for chunk in (None, "96"):
        if chunk is None:
            monkeypatch.delenv("VELO_HOST_CHUNK_CELLS", raising=False)
        else:
            monkeypatch.setenv("VELO_HOST_CHUNK_CELLS", chunk)
        for kind in ("pageable", "pinned"):
            if kind == "pinned":
                bufs = [torch.from_numpy(a).pin_memory() for a in (e, d, ixs)]
                out = torch.empty((C, m), dtype=torch.float32).pin_memory()
            else:
                bufs = [torch.from_numpy(a) for a in (e, d, ixs)]
                out = torch.empty((C, m), dtype=torch.float32)
            _cabi.call("velo_transition_prob_partial", _cabi.SQRT, bufs[0].data_ptr(), bufs[1].data_ptr(), 8,
                       bufs[2].data_ptr(), out.data_ptr(), G, C, m, psc, sigma)
            res[(chunk, kind)] = out.numpy().copy()
    first = res[(None, "pageable")]
    for k, v in res.items():
        assert np.array_equal(first, v), k
    np.testing.assert_allclose(first, want, rtol=1e-5, atol=0)


def test_sharded_host_entry_matches_single_call(oracle, monkeypatch):
    """The cell-sharded host tier on ONE GPU playing both ranks of a 2-way split: velo_upload_cellmajor of each block
    into its slot (column blocks of the full host matrix, passed as strided views), then
    velo_transition_prob_partial_sharded per block == the single-GPU host call, bit for bit."""
    import ctypes
    import torch
    from velocyto_b200 import _cabi, device as dev
    from velocyto_b200.sharding import partition, block_size, _host_ptr
    G, C, m, psc, sigma = 900, 501, 37, 1.0, 0.05
    e, d, ixs = _host_problem(G, C, m, 43)
    whole = np.empty((C, m), dtype=np.float32)
    _cabi.call("velo_transition_prob_partial", _cabi.SQRT, e.ctypes.data, d.ctypes.data, 8, ixs.ctypes.data,
               whole.ctypes.data, G, C, m, psc, sigma)
    world, ld = 2, dev.padded_ld(G)
    b = block_size(C, world)
    e_full = torch.zeros((world * b, ld), dtype=torch.float32, device="cuda")
    lo_full = torch.zeros_like(e_full)
    stream = torch.cuda.current_stream().cuda_stream
    nz_any = 0
    for r, (c0, nc) in enumerate(partition(C, world)):
        ptr, pitch, esz = _host_ptr(e[:, c0:c0 + nc])                 # a view: row pitch = C values
        assert pitch == C and esz == 8
        nz = ctypes.c_int(0)
        _cabi.call("velo_upload_cellmajor", ptr, 8, G, nc, pitch, e_full[r * b:].data_ptr(), lo_full[r * b:].data_ptr(),
                   ctypes.addressof(nz), ld, stream)
        nz_any |= nz.value
    assert nz_any == 1                                                # raw float64 data: residuals present
    monkeypatch.setenv("VELO_HOST_CHUNK_CELLS", "64")
    for r, (c0, nc) in enumerate(partition(C, world)):
        out = np.empty((nc, m), dtype=np.float32)
        dptr, dpitch, _ = _host_ptr(d[:, c0:c0 + nc])
        _cabi.call("velo_transition_prob_partial_sharded", _cabi.SQRT, e_full.data_ptr(), lo_full.data_ptr(), ld, stream,
                   dptr, 8, dpitch, ixs[c0:c0 + nc].ctypes.data, out.ctypes.data, G, C, c0, nc, m, psc, sigma)
        assert np.array_equal(out, whole[c0:c0 + nc]), f"block {r}"
    # the Python front (world size 1: no process group) drives the same two entry points
    from velocyto_b200.sharding import CellShardedHostTransitionProb
    out1 = np.empty((C, m), dtype=np.float32)
    CellShardedHostTransitionProb(G, C, "sqrt", psc, sigma).run(e, d, ixs, out1)
    assert np.array_equal(out1, whole)


def test_transition_prob_small_sigma_and_nan_rules():
    """Row-max subtraction: sigma_corr far below the fp32 overflow point of exp(corr / sigma) (0.0113) still gives the
    fp64 result; NaN -> 1 only when asked (knn_random branch), otherwise the row is NaN (full branch)."""
    import torch
    from velocyto_b200 import device as dev
    rng = np.random.default_rng(3)
    nc, m = 64, 200
    corr = rng.uniform(-1, 1, (nc, m)).astype(np.float32)
    corr[5, 7] = np.nan
    ix = np.stack([(c + 1 + rng.choice(300 - 1, m, replace=False)) % 300 for c in range(nc)])
    ix[9, 3] = 9                                                       # a self pair -> 0
    for sigma in (0.05, 0.004):
        c64 = corr.astype(np.float64)
        c64[9, 3] = 0
        c64[np.isnan(c64)] = 1
        w = np.exp((c64 - c64.max(1, keepdims=True)) / sigma)
        want = w / w.sum(1, keepdims=True)
        got = dev.transition_prob(torch.from_numpy(corr).cuda(), dev.indices_to_device(ix, 300), sigma).cpu().numpy()
        assert np.isfinite(got).all()
        np.testing.assert_allclose(got, want, rtol=5e-5 if sigma < 0.01 else 2e-6, atol=1e-30)
    got = dev.transition_prob(torch.from_numpy(corr).cuda(), dev.indices_to_device(ix, 300), 0.05, patch_nan=False).cpu().numpy()
    assert np.isnan(got[5]).all() and np.isfinite(np.delete(got, 5, 0)).all()


@pytest.mark.parametrize("C", [1300, 1537])
def test_tensor_core_symmetric_scheme_matches_plain(oracle, C, monkeypatch):
    """K2g symmetric-Q scheme (launch 1: tiles on/right of the diagonal compute P and Q and keep Q; launch 2: tiles left
    of it compute two P tiles and read Q[c, i] = Q[i, c]) against the plain scheme and the oracle -- odd and even
    numbers of 128-cell blocks, a ragged last block."""
    from velocyto_b200 import device as dev
    G = 900
    e, d = _tc_inputs(G, C, 77)
    e_cm, d_cm = dev.CellMajor.from_gene_major(e), dev.CellMajor.from_gene_major(d)
    monkeypatch.setenv("VELO_TC_SYMMETRIC", "0")
    plain = dev.coldeltacor_linear_tc(e_cm, d_cm).cpu().numpy()
    monkeypatch.setenv("VELO_TC_SYMMETRIC", "1")
    sym, P, Q = dev.coldeltacor_linear_tc(e_cm, d_cm, debug=True)
    sym, Q = sym.cpu().numpy(), Q.cpu().numpy()
    off = ~np.eye(C, dtype=bool)
    assert np.array_equal(np.isnan(sym), np.isnan(plain))
    np.testing.assert_allclose(sym[off], plain[off], rtol=0, atol=2e-7)
    np.testing.assert_allclose(Q, Q.T, rtol=1e-6, atol=1e-6 * np.abs(Q).max())          # what the scheme relies on
    want = oracle.coldeltacor(e, d, None, "linear", 0.0)
    np.testing.assert_allclose(sym[off], want[off], rtol=0, atol=5e-7)
