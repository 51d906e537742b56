"""Pin the oracle (oracle/) on outputs of the unmodified reference.

The reference has no tests / golden vectors; the fixtures in
tests/golden/ were produced by tests/golden/make_golden.py, which runs the
reference's own compiled kernel and Python layers.  When oracle/_ref is present
the C restatement is additionally compared with the live reference kernel.
"""
import numpy as np
import pytest
from scipy import sparse

CASES = [("sqrt", 1e-10), ("sqrt", 1.0), ("log10", 1.0), ("log10", 0.5), ("linear", 0.0)]


def _d(z, name, psc):
    if name == "sqrt":
        return np.sqrt(np.abs(z) + psc) * np.sign(z)
    if name == "log10":
        return np.log10(np.abs(z) + psc) * np.sign(z)
    return z


def _degenerate_pairs(e):
    C = e.shape[1]
    return np.array([[np.array_equal(e[:, i], e[:, c]) for i in range(C)] for c in range(C)])


@pytest.mark.parametrize("name,psc", CASES)
def test_coldeltacor_oracle_matches_golden(oracle, golden, name, psc):
    g = golden("coldeltacor_small.npz")
    e, z, ixs = g["e"], g["z"], g["ixs"]
    d = _d(z, name, psc)
    tag = f"{name}_{psc:g}"
    full = oracle.coldeltacor(e, d, None, name, psc, threads=2)
    part = oracle.coldeltacor(e, d, ixs, name, psc, threads=2)
    # pairs of identical columns (i == c, or the duplicated cell) have a zero-variance A: the reference
    # returns NaN or rounding garbage there depending on the variant; callers overwrite them
    # (analysis.py:1604-1612, 1666).  Everything else must agree to fp64 rounding.
    ok = ~_degenerate_pairs(e)
    for got, want in ((full, g[f"full_{tag}"]), (part, g[f"partial_{tag}"])):
        assert not np.isnan(got[ok]).any() and not np.isnan(want[ok]).any()
        np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=2e-12)
    # the partial sqrt variant zeroes exact-zero differences -> NaN on identical columns in both
    if name == "sqrt":
        deg = ~ok
        assert np.array_equal(np.isnan(part[deg]), np.isnan(g[f"partial_{tag}"][deg]))
    # partial touches only the sampled entries
    mask = np.zeros_like(part, dtype=bool)
    mask[np.arange(ixs.shape[0])[:, None], ixs] = True
    assert np.all(part[~mask] == 0)


@pytest.mark.parametrize("name,psc", CASES)
def test_selected_cells_oracle_matches_golden(oracle, golden, name, psc):
    """``coldeltacor_cells`` (the few-cells entry used for the bench-shape spot checks) against the reference's golden
    vectors, both zero rules: it must be the same arithmetic as the whole-problem oracle pinned above."""
    g = golden("coldeltacor_small.npz")
    e, ixs = g["e"], g["ixs"].astype(np.int64)
    d = _d(g["z"], name, psc)
    tag = f"{name}_{psc:g}"
    C = e.shape[1]
    cells = np.array([0, 3, C // 2, C - 1])
    ok = ~_degenerate_pairs(e)
    part = oracle.coldeltacor_cells(e, d[:, cells], cells, ixs[cells], name, psc, partial=True, threads=2)
    want = g[f"partial_{tag}"][cells[:, None], ixs[cells]]
    sel = ok[cells[:, None], ixs[cells]]
    np.testing.assert_allclose(part[sel], want[sel], rtol=0, atol=2e-12)
    everyone = np.tile(np.arange(C), (cells.size, 1))
    full = oracle.coldeltacor_cells(e, d[:, cells], cells, everyone, name, psc, partial=False, threads=2)
    np.testing.assert_allclose(full[ok[cells]], g[f"full_{tag}"][cells][ok[cells]], rtol=0, atol=2e-12)
    # and bit-for-bit the whole-problem oracle (same operation order per pair)
    whole = oracle.coldeltacor(e, d, ixs, name, psc, threads=2)[cells[:, None], ixs[cells]]
    assert np.array_equal(part, whole, equal_nan=True)


@pytest.mark.parametrize("name,psc", CASES)
def test_coldeltacor_oracle_matches_live_reference(oracle, name, psc):
    if oracle.load_ref_speedboosted() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(42)
    G, C, m = 211, 67, 19
    e = rng.gamma(2.0, 1.0, (G, C))
    d = _d(rng.normal(size=(G, C)), name, psc)
    ixs = np.stack([rng.choice(C, m, replace=False) for _ in range(C)])
    for ix in (None, ixs):
        got = oracle.coldeltacor(e, d, ix, name, psc)
        want = oracle.ref_coldeltacor(e, d, ix, name, psc)
        off = ~np.eye(C, dtype=bool)      # diagonals are degenerate (NaN or rounding garbage) in both
        if ix is not None:
            sampled = np.zeros((C, C), dtype=bool)
            sampled[np.arange(C)[:, None], ix] = True
            assert np.all(got[~sampled] == 0) and np.all(want[~sampled] == 0)   # only sampled entries are touched
            off &= sampled
        assert not np.isnan(got[off]).any() and not np.isnan(want[off]).any()
        np.testing.assert_allclose(got[off], want[off], rtol=0, atol=5e-13)


def test_fit_slopes_oracle_matches_golden(oracle, golden):
    g = golden("fit_slopes_small.npz")
    X, Y, W = g["X"], g["Y"], g["W"]
    eq = dict(rtol=1e-6, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(oracle.fit_slope(Y, X), g["slope"], **eq)
    s, q = oracle.fit_slope_offset(Y, X)
    np.testing.assert_allclose(s, g["slope_offset_g"], **eq)
    np.testing.assert_allclose(q, g["slope_offset_q"], **eq)
    s, q = oracle.fit_slope_offset(Y, X, fixperc_q=True)
    np.testing.assert_allclose(s, g["slope_offset_fix_g"], **eq)
    np.testing.assert_allclose(q, g["slope_offset_fix_q"], **eq)
    s, r2 = oracle.fit_slope_weighted(Y, X, W, return_R2=True)
    np.testing.assert_allclose(s, g["weighted_g"], **eq)
    np.testing.assert_allclose(r2, g["weighted_R2"], **eq)
    np.testing.assert_allclose(oracle.fit_slope_weighted(Y, X, W, limit_gamma=True), g["weighted_lim_g"], **eq)
    s, q, r2 = oracle.fit_slope_weighted_offset(Y, X, W, return_R2=True)
    np.testing.assert_allclose(s, g["weighted_offset_g"], **eq)
    np.testing.assert_allclose(q, g["weighted_offset_q"], **eq)
    np.testing.assert_allclose(r2, g["weighted_offset_R2"], **eq)
    s, q = oracle.fit_slope_weighted_offset(Y, X, W, fixperc_q=True, return_R2=False)
    np.testing.assert_allclose(s, g["weighted_offset_fix_g"], **eq)
    np.testing.assert_allclose(q, g["weighted_offset_fix_q"], **eq)
    # degenerate genes: x == 0 -> NaN, y == 0 -> 0 (estimation.py:176-179)
    assert np.isnan(g["slope"][2]) and g["slope"][5] == 0


def test_knn_smoothing_oracle_matches_golden(oracle, golden):
    g = golden("knn_smoothing_small.npz")
    C = g["S"].shape[1]
    knn = sparse.csr_matrix((g["knn_data"], g["knn_indices"], g["knn_indptr"]), shape=(C, C))
    for diag in (1, 8):
        Sx, Ux, w = oracle.knn_imputation(g["S"], g["U"], knn, diag=diag)
        np.testing.assert_allclose(sparse.csr_matrix(w).toarray(), g[f"w_dense_diag{diag}"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(Sx, g[f"Sx_diag{diag}"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(Ux, g[f"Ux_diag{diag}"], rtol=1e-13, atol=1e-13)


def test_pipeline_oracle_matches_golden(oracle, golden):
    g = golden("pipeline_small.npz")
    C = g["S_sz"].shape[1]
    knn = sparse.csr_matrix((g["knn_data"], g["knn_indices"], g["knn_indptr"]), shape=(C, C))
    Sx, Ux, _ = oracle.knn_imputation(g["S_sz"], g["U_sz"], knn, diag=1)
    np.testing.assert_allclose(Sx, g["Sx_sz"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(Ux, g["Ux_sz"], rtol=1e-13, atol=1e-13)
    Sx, Ux = g["Sx_sz"], g["Ux_sz"]
    np.testing.assert_allclose(oracle.fit_slope(Ux, Sx), g["gammas_nnls"], rtol=1e-6)
    gam, q = oracle.fit_slope_offset(Ux, Sx)
    np.testing.assert_allclose(gam, g["gammas_ols"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(q, g["q_ols"], rtol=1e-6, atol=1e-8)
    W = oracle.gamma_fit_weights("maxmin_diag", Sx, Ux, Sx, Ux)
    gd, qd, r2 = oracle.fit_slope_weighted_offset(Ux, Sx, W, return_R2=True)
    np.testing.assert_allclose(gd, g["gammas_default"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(qd, g["q_default"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(r2, g["R2_default"], rtol=1e-5, atol=1e-7)
    Upred = oracle.predict_U(gam.astype(np.float32), Sx, q.astype(np.float32))
    np.testing.assert_allclose(Upred, g["Upred"], rtol=1e-12, atol=1e-12)
    vel = oracle.calculate_velocity(Ux, Upred)
    np.testing.assert_allclose(vel, g["velocity"], rtol=1e-12, atol=1e-12)
    dS = oracle.calculate_shift(vel, 1.0)
    np.testing.assert_allclose(dS, g["delta_S"], rtol=1e-12, atol=1e-12)
    St = oracle.extrapolate_cell_at_t(Sx, dS, 1.0)
    np.testing.assert_allclose(St, g["Sx_sz_t"], rtol=1e-12, atol=1e-12)
    # estimate_transition_prob, knn_random mode (analysis.py:1528-1612)
    psc = 1.0
    neigh = g["neigh_ixs"].astype(np.int64)
    for dkey, ckey, pkey in (("delta_S", "corrcoef", "transition_prob"),
                             ("delta_S_rndm", "corrcoef_random", "transition_prob_random")):
        delta = (Sx + 1.0 * g[dkey]) - Sx
        corr = oracle.colDeltaCorSqrtpartial(Sx, oracle.velocity_transform(delta, "sqrt", psc), neigh, psc=psc)
        oracle.patch_corrcoef(corr)
        np.testing.assert_allclose(corr, g[ckey], rtol=0, atol=1e-12)
        tp = oracle.transition_prob(corr, oracle.neighbors_to_csr(neigh), 0.05)
        np.testing.assert_allclose(tp, g[pkey], rtol=1e-9, atol=1e-15)
    de = oracle.embedding_shift(g["transition_prob"], g["embedding"], oracle.neighbors_to_csr(neigh))
    np.testing.assert_allclose(de, g["delta_embedding"], rtol=1e-9, atol=1e-12)
    # full mode (analysis.py:1613-1668)
    delta = (Sx + 1.0 * g["delta_S"]) - Sx
    corr = oracle.colDeltaCorSqrt(Sx, oracle.velocity_transform(delta, "sqrt", psc), psc=psc)
    np.fill_diagonal(corr, 0)
    # cells 7 and 22 converge to identical profiles after imputation: a zero-variance column, where the
    # full sqrt variant returns rounding garbage (no zero rule, speedboosted.pyx:110-114) -- excluded.
    ok = ~_degenerate_pairs(Sx)
    assert (~ok).sum() == C + 2
    np.testing.assert_allclose(corr[ok], g["full_corrcoef"][ok], rtol=0, atol=1e-12)
    fk = g["full_knn_indices"]
    mask = sparse.csr_matrix((np.ones(fk.size), fk.ravel(), np.arange(0, fk.size + 1, fk.shape[1])), shape=(C, C))
    tp = oracle.transition_prob(np.where(ok | np.eye(C, dtype=bool), corr, g["full_corrcoef"]), mask, 0.05)
    np.testing.assert_allclose(tp, g["full_transition_prob"], rtol=1e-9, atol=1e-15)


def test_sampler_matches_golden(oracle, golden):
    """np.random.choice stream of the neighbour sampler (analysis.py:1552-1566)."""
    from sklearn.neighbors import NearestNeighbors
    g = golden("pipeline_small.npz")
    nn = NearestNeighbors(n_neighbors=31, n_jobs=1).fit(g["embedding"])
    knn_idx = nn.kneighbors_graph(mode="connectivity").indices.reshape(-1, 31)
    neigh, samp = oracle.sample_neighbors(knn_idx, 0.5, (0.5, 0.1), 15071990)
    assert np.array_equal(samp, g["sampling_ixs"])
    assert np.array_equal(neigh, g["neigh_ixs"])


def test_normalize_oracle_matches_golden(oracle, golden):
    """analysis.py:535-676 restated in oracle.size_log_normalize == the reference's normalize() outputs."""
    g = golden("normalize_small.npz")
    S, U, Sx, Ux = g["S"], g["U"], g["Sx"], g["Ux"]
    C = S.shape[1]

    def check(tag, name_sz, name_norm, got):
        X_sz, X_norm = got[0], got[1]
        np.testing.assert_allclose(X_sz, g[f"{tag}__{name_sz}"], rtol=1e-14, atol=0, equal_nan=True)
        if f"{tag}__{name_norm}" in g.files:
            np.testing.assert_allclose(X_norm, g[f"{tag}__{name_norm}"], rtol=1e-14, atol=1e-15, equal_nan=True)
        else:
            assert X_norm is None

    r = oracle.size_log_normalize(S)
    check("default", "S_sz", "S_norm", r)
    np.testing.assert_array_equal(r[2], g["default__cell_size"])
    check("default", "U_sz", "U_norm", oracle.size_log_normalize(U, guard=True))
    check("opts", "S_sz", "S_norm", oracle.size_log_normalize(S, pcount=0.5, target_size=1000.0))
    check("opts", "U_sz", "U_norm", oracle.size_log_normalize(U, pcount=0.5, cell_size=S.sum(0), target_size=500.0, guard=True))
    check("nosize", "S_sz", "S_norm", oracle.size_log_normalize(S, size=False))
    check("nolog", "U_sz", "U_norm", oracle.size_log_normalize(U, log=False, guard=True))
    rel = np.linspace(50.0, 400.0, C)
    check("relsize", "S_sz", "S_norm", oracle.size_log_normalize(S, cell_size=rel))
    check("relsize", "U_sz", "U_norm", oracle.size_log_normalize(U, cell_size=rel, guard=True))
    check("imputed", "Sx_sz", "Sx_norm", oracle.size_log_normalize(Sx))
    check("imputed", "Ux_sz", "Ux_norm", oracle.size_log_normalize(Ux, guard=True))
    check("imputed_opts", "Sx_sz", "Sx_norm", oracle.size_log_normalize(Sx, pcount=2.0, target_size=800.0))
    # use_Sx_size without a previous S normalisation: hasattr(self, "cell_size") is False -> Sx.sum(0) (analysis.py:607-610)
    check("imputed_opts", "Ux_sz", "Ux_norm", oracle.size_log_normalize(Ux, pcount=2.0, cell_size=Sx.sum(0), guard=True))


def test_gemm_identity_and_fp16_pair_split_reproduce_linear_coldeltacor(oracle):
    """Numerical design of the tensor-core kernel K2g (csrc/coldeltacor_tc.cu), checked on the host against the oracle:
    (1) x_colDeltaCor (speedboosted.pyx:13-87) == the two gene-axis products P = B X^T, Q = X X^T plus the epilogue;
    (2) scaling every operand row by a power of two and splitting it into an fp16 (hi, lo) pair, with the three
        products hi*hi + hi*lo + lo*hi, keeps the correlation within 5e-8 -- a bf16 pair does not (> 2e-7)."""
    import torch
    G, C = 1500, 96
    rng = np.random.default_rng(3)
    level = rng.gamma(0.6, 2.0, G)[:, None] + 0.05
    e = (rng.gamma(2.0, 1.0, (G, C)) * level).astype(np.float32).astype(np.float64)
    d = rng.normal(size=(G, C)).astype(np.float32).astype(np.float64)
    want = oracle.colDeltaCor(e, d)
    off = ~np.eye(C, dtype=bool)
    X = e - e.mean(1, keepdims=True)
    X = (X - X.mean(0, keepdims=True)).astype(np.float32).astype(np.float64)
    B = (d - d.mean(0, keepdims=True)).astype(np.float32).astype(np.float64)
    qd, pcc, sbb = (X * X).sum(0), (B * X).sum(0), (B * B).sum(0)

    def corr_from(P, Q):
        with np.errstate(invalid="ignore", divide="ignore"):
            return (P - pcc[:, None]) / np.sqrt((qd[:, None] + qd[None, :] - 2 * Q) * sbb[:, None])

    # (1) -- with the centred operands rounded to fp32 as on the device (that rounding alone costs ~4e-9)
    np.testing.assert_allclose(corr_from(B.T @ X, X.T @ X)[off], want[off], rtol=0, atol=1e-8)

    def split(A, dtype):
        s = 2.0 ** (13 - np.floor(np.log2(np.abs(A).max(0)))) if dtype == torch.float16 else np.ones(A.shape[1])
        t = torch.from_numpy(A * s)
        hi = t.to(dtype).double()
        lo = (t - hi).to(dtype).double()
        return hi.numpy(), lo.numpy(), s

    errs = {}
    for dtype in (torch.float16, torch.bfloat16):
        Xh, Xl, sx = split(X, dtype)
        Bh, Bl, sb = split(B, dtype)
        P = (Bh.T @ Xh + Bh.T @ Xl + Bl.T @ Xh) / sb[:, None] / sx[None, :]
        Q = (Xh.T @ Xh + Xh.T @ Xl + Xl.T @ Xh) / sx[:, None] / sx[None, :]
        errs[dtype] = np.abs(corr_from(P, Q) - want)[off].max()
    assert errs[torch.float16] < 5e-8, errs                                                      # (2)
    assert errs[torch.bfloat16] > 2e-7, errs


def test_config1_fit_gammas_reference_path_at_full_shape(oracle):
    """BASELINE config 1 (1k cells x 2k genes, fit_gammas on the reference CPU path; plumbing, no GPU): the oracle's
    restatement of the unweighted fits (estimation.py:173-188, 244-297: one SciPy nnls / leastsq call per gene) at the
    full shape against the solver-independent closed forms SURVEY.md section 7 established -- nnls == max(0, Sxy / Sxx) to
    float32 rounding, leastsq == OLS to ~1e-5 -- which are also what the device kernel K4 evaluates."""
    rng = np.random.default_rng(1)
    G, C = 2000, 1000
    mu = rng.gamma(0.6, 2.0, G)
    s = rng.gamma(2.0, 0.5, C)
    gam = rng.uniform(0.05, 1.0, G)
    S = rng.poisson(mu[:, None] * s[None, :]).astype(np.float64) + rng.uniform(0, 0.5, (G, C))
    U = rng.poisson(mu[:, None] * s[None, :] * gam[:, None]).astype(np.float64) + rng.uniform(0, 0.5, (G, C))
    got = oracle.fit_slope(U, S)                                       # Y = U, X = S (analysis.py:1254-1256)
    want = np.maximum(0.0, (S * U).sum(1) / (S * S).sum(1))
    assert got.dtype == np.float32 and got.shape == (G,)
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=0)
    g_o, q_o = oracle.fit_slope_offset(U, S)
    xm, ym = S.mean(1, keepdims=True), U.mean(1, keepdims=True)
    slope = ((S - xm) * (U - ym)).sum(1) / ((S - xm) ** 2).sum(1)
    np.testing.assert_allclose(g_o, slope, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(q_o, (ym - slope[:, None] * xm).ravel(), rtol=2e-5, atol=2e-5)
