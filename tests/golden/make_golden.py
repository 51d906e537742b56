"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference and oracle/_ref):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The reference ships no tests and no golden vectors, so
parity is pinned on outputs of the reference itself: its Cython kernel compiled
from the untouched ``speedboosted.pyx`` (oracle/build_ref.sh) and its untouched
``estimation.py`` / ``neighbors.py`` / ``analysis.py`` imported from where they lie.
``velocyto/__init__.py`` is bypassed (it imports pysam/loompy/h5py, absent here);
the shims below only provide import-time stand-ins and the NumPy-2/SciPy API
names the 2019 sources still use.  No reference source is modified or copied.

The fixtures are small ``.npz`` files (inputs + reference outputs); the tests
never need /root/reference.
"""
import importlib
import os
import sys
import types

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VELO_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference():
    """Import reference modules without executing velocyto/__init__.py."""
    from oracle import velo_oracle as vo
    sb = vo.load_ref_speedboosted()
    assert sb is not None, "run oracle/build_ref.sh first"
    # NumPy 2 / SciPy compatibility names used by the 2019 sources
    if not hasattr(np, "NAN"):
        np.NAN = np.nan                       # estimation.py:177,195,217,248
    if not hasattr(np, "string_"):
        np.string_ = np.bytes_                # analysis.py:187
    _stack = np.stack

    def stack(arrays, *a, **k):               # analysis.py:1561 passes a generator
        if not isinstance(arrays, (list, tuple, np.ndarray)):
            arrays = list(arrays)
        return _stack(arrays, *a, **k)
    np.stack = stack
    for cls in (sparse.csr_matrix, sparse.csc_matrix, sparse.coo_matrix, sparse.lil_matrix):
        if not hasattr(cls, "A"):
            cls.A = property(lambda self: self.toarray())      # analysis.py:1697
    if not hasattr(sparse, "csr") or not hasattr(getattr(sparse, "csr", None), "csr_matrix"):
        sparse.csr = types.SimpleNamespace(csr_matrix=sparse.csr_matrix)   # neighbors.py:379,385 annotations
    # import-time stand-ins for absent plotting / IO dependencies
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors",
                 "matplotlib.cm", "loompy", "h5py"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                sys.modules[name] = m
    mpl = sys.modules["matplotlib"]
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "cm"):
        class _CM:
            def __getattr__(self, k):
                return lambda x=None, *a, **kw: np.zeros((np.size(x) if x is not None else 1, 4))
        plt.cm = _CM()
    for sub in ("pyplot", "gridspec", "colors", "cm"):
        if not hasattr(mpl, sub):
            setattr(mpl, sub, sys.modules["matplotlib." + sub])
    if not hasattr(sys.modules["matplotlib.colors"], "Normalize"):
        sys.modules["matplotlib.colors"].Normalize = object
        sys.modules["matplotlib.colors"].LinearSegmentedColormap = object
    pkg = types.ModuleType("velocyto")
    pkg.__path__ = [os.path.join(REF, "velocyto")]
    sys.modules["velocyto"] = pkg
    sys.modules["velocyto.speedboosted"] = sb
    est = importlib.import_module("velocyto.estimation")
    nb = importlib.import_module("velocyto.neighbors")
    try:
        an = importlib.import_module("velocyto.analysis")
    except Exception as exc:                 # pragma: no cover - reported, not fatal
        print("analysis.py import failed:", repr(exc))
        an = None
    return sb, est, nb, an


def synth_counts(G, C, seed):
    """SURVEY.md 8(d) pipeline-level generator."""
    rng = np.random.default_rng(seed)
    mu = rng.gamma(0.6, 2.0, G)
    s = rng.gamma(2.0, 0.5, C)
    gam = rng.uniform(0.05, 1.0, G)
    S = rng.poisson(mu[:, None] * s[None, :]).astype(np.float64)
    U = rng.poisson(mu[:, None] * s[None, :] * gam[:, None] * rng.uniform(0.5, 1.5, (G, C))).astype(np.float64)
    return S, U


def golden_coldeltacor(est, out):
    rng = np.random.default_rng(0)
    G, C, m = 37, 23, 6
    e = rng.gamma(2.0, 1.0, (G, C))
    e[:, 5] = e[:, 4]                      # two identical cells -> NaN / zero-variance branch
    e[3:9, :] = 0.0                        # genes with all-zero differences
    e[11, ::2] = e[11, 0]                  # exact zero differences inside a column
    z = rng.normal(size=(G, C))
    ixs = np.stack([rng.choice(C, m, replace=False) for _ in range(C)])
    ixs[2, 0] = 2                          # a row sampling itself (allowed by the sampler, analysis.py:1556-1560)
    ixs[4, 1] = 5                          # identical-cell pair
    res = {"e": e, "z": z, "ixs": ixs}
    for name, psc in (("sqrt", 1e-10), ("sqrt", 1.0), ("log10", 1.0), ("log10", 0.5), ("linear", 0.0)):
        tag = f"{name}_{psc:g}"
        if name == "sqrt":
            d = np.sqrt(np.abs(z) + psc) * np.sign(z)
            res[f"full_{tag}"] = est.colDeltaCorSqrt(e, d, threads=1, psc=psc)
            res[f"partial_{tag}"] = est.colDeltaCorSqrtpartial(e, d, ixs, threads=1, psc=psc)
        elif name == "log10":
            d = np.log10(np.abs(z) + psc) * np.sign(z)
            res[f"full_{tag}"] = est.colDeltaCorLog10(e, d, threads=1, psc=psc)
            res[f"partial_{tag}"] = est.colDeltaCorLog10partial(e, d, ixs, threads=1, psc=psc)
        else:
            d = z
            res[f"full_{tag}"] = est.colDeltaCor(e, d, threads=1)
            res[f"partial_{tag}"] = est.colDeltaCorpartial(e, d, ixs, threads=1)
    np.savez_compressed(os.path.join(out, "coldeltacor_small.npz"), **res)


def golden_fits(est, an, out):
    G, C = 28, 60
    S, U = synth_counts(G, C, 3)
    rng = np.random.default_rng(4)
    X = S + rng.uniform(0, 0.5, S.shape)
    Y = U + rng.uniform(0, 0.5, U.shape)
    X[2] = 0.0                             # x == 0 -> NaN slope
    Y[5] = 0.0                             # y == 0 -> 0 slope
    Y[7] = 3.0 * X[7] + 0.25               # exact line
    Y[9] = np.maximum(0, 2.0 - X[9])       # negative OLS slope -> nnls clamps to 0
    W = rng.uniform(0.0, 1.0, (G, C))
    W[11] = (rng.uniform(size=C) < 0.2).astype(float)
    res = {"X": X, "Y": Y, "W": W}
    res["slope"] = est.fit_slope(Y, X)
    res["slope_offset_g"], res["slope_offset_q"] = est.fit_slope_offset(Y, X)
    res["slope_offset_fix_g"], res["slope_offset_fix_q"] = est.fit_slope_offset(Y, X, fixperc_q=True)
    res["weighted_g"], res["weighted_R2"] = est.fit_slope_weighted(Y, X, W, return_R2=True)
    res["weighted_lim_g"] = est.fit_slope_weighted(Y, X, W, limit_gamma=True)
    g, q, r2 = est.fit_slope_weighted_offset(Y, X, W, return_R2=True)
    res["weighted_offset_g"], res["weighted_offset_q"], res["weighted_offset_R2"] = g, q, r2
    g, q = est.fit_slope_weighted_offset(Y, X, W, fixperc_q=True, return_R2=False)
    res["weighted_offset_fix_g"], res["weighted_offset_fix_q"] = g, q
    np.savez_compressed(os.path.join(out, "fit_slopes_small.npz"), **res)


def golden_smoothing(nb, out):
    from sklearn.neighbors import NearestNeighbors
    G, C, k = 19, 40, 6
    S, U = synth_counts(G, C, 5)
    rng = np.random.default_rng(6)
    space = rng.normal(size=(C, 5))
    space[7] = space[3]                    # zero-distance edge: dropped by (knn > 0) (analysis.py:1006)
    knn = nb.knn_distance_matrix(space, metric="euclidean", k=k, mode="distance", n_jobs=1)
    res = {"S": S, "U": U, "knn_data": knn.data, "knn_indices": knn.indices, "knn_indptr": knn.indptr}
    for diag in (1, 8):
        conn = (knn > 0).astype(float)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            conn.setdiag(diag)
        w = nb.connectivity_to_weights(conn)
        res[f"w_dense_diag{diag}"] = np.asarray(sparse.csr_matrix(w).toarray())
        res[f"Sx_diag{diag}"] = np.asarray(nb.convolve_by_sparse_weights(S, w))
        res[f"Ux_diag{diag}"] = np.asarray(nb.convolve_by_sparse_weights(U, w))
    # balanced kNN (neighbors.py:186-321): hub-limited graph, plain and group-constrained
    rng2 = np.random.default_rng(7)
    pts = np.concatenate([rng2.normal(size=(70, 4)), rng2.normal(size=(10, 4)) * 0.05])   # a tight hub cluster
    groups = (np.arange(80) % 3).astype(np.int64)
    for tag, cons in (("plain", None), ("grouped", groups)):
        b = nb.BalancedKNN(k=8, sight_k=30, maxl=12, constraint=cons, mode="distance", n_jobs=1)
        b.fit(pts)
        gph = b.kneighbors_graph(mode="distance")
        res[f"bknn_{tag}_data"], res[f"bknn_{tag}_indices"], res[f"bknn_{tag}_l"] = gph.data, gph.indices, b.l
    res["bknn_points"], res["bknn_groups"] = pts, groups
    np.savez_compressed(os.path.join(out, "knn_smoothing_small.npz"), **res)


def golden_pipeline(an, out):
    """knn_imputation -> ... -> calculate_embedding_shift on a VelocytoLoom built by hand."""
    if an is None:
        print("skipping pipeline golden (analysis.py not importable)")
        return
    G, C = 48, 72
    S, U = synth_counts(G, C, 11)
    keep = (S.sum(1) > 0) & (U.sum(1) > 0)
    S, U = S[keep], U[keep]
    G = S.shape[0]
    vlm = an.VelocytoLoom.__new__(an.VelocytoLoom)
    vlm.S, vlm.U, vlm.A = S, U, np.zeros_like(S)
    vlm.ca, vlm.ra = {"CellID": np.arange(C)}, {"Gene": np.arange(G)}
    vlm.initial_cell_size = S.sum(0)
    vlm.initial_Ucell_size = U.sum(0)
    # size normalisation as _normalize_S/_normalize_U(size=True, log=False) would leave it
    vlm.S_sz = S / S.sum(0) * np.mean(S.sum(0))
    vlm.U_sz = U / np.maximum(U.sum(0), 1) * np.mean(U.sum(0))
    Sn = np.log2(vlm.S_sz + 1)
    Sc = Sn - Sn.mean(1)[:, None]
    u, s, vt = np.linalg.svd(Sc.T, full_matrices=False)
    vlm.pcs = u[:, :8] * s[:8]
    vlm.knn_imputation(k=9, pca_space=True, n_pca_dims=6, balanced=False, n_jobs=1)
    res = {"S_sz": vlm.S_sz, "U_sz": vlm.U_sz, "pcs": vlm.pcs,
           "knn_data": vlm.knn.data, "knn_indices": vlm.knn.indices, "knn_indptr": vlm.knn.indptr,
           "Sx_sz": np.ascontiguousarray(vlm.Sx_sz), "Ux_sz": np.ascontiguousarray(vlm.Ux_sz)}
    # make the arrays C-contiguous as a gene filter would
    for a in ("Sx", "Ux", "Sx_sz", "Ux_sz"):
        setattr(vlm, a, np.ascontiguousarray(getattr(vlm, a)))
    vlm.fit_gammas(weighted=False, fit_offset=False)
    res["gammas_nnls"] = vlm.gammas.copy()
    vlm.fit_gammas(weighted=False, fit_offset=True)
    res["gammas_ols"], res["q_ols"] = vlm.gammas.copy(), vlm.q.copy()
    vlm.fit_gammas()                                           # default: maxmin_diag weights + offset (L-BFGS-B)
    res["gammas_default"], res["q_default"], res["R2_default"] = vlm.gammas.copy(), vlm.q.copy(), vlm.R2.copy()
    vlm.fit_gammas(weighted=False, fit_offset=True)            # the pipeline continues from the OLS fit
    vlm.predict_U()
    vlm.calculate_velocity()
    vlm.calculate_shift(assumption="constant_velocity")
    vlm.extrapolate_cell_at_t(delta_t=1.0)
    res.update(Upred=vlm.Upred, velocity=vlm.velocity, delta_S=vlm.delta_S, Sx_sz_t=vlm.Sx_sz_t)
    vlm.ts = vlm.pcs[:, :2].copy()
    vlm.estimate_transition_prob(hidim="Sx_sz", embed="ts", transform="sqrt", psc=1, n_neighbors=30,
                                 knn_random=True, sampled_fraction=0.5, n_jobs=1, threads=1)
    vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=False)
    res.update(embedding=vlm.embedding, sampling_ixs=vlm.sampling_ixs,
               neigh_ixs=vlm.embedding_knn.indices.reshape(C, -1).copy(),
               delta_S_rndm=vlm.delta_S_rndm, corrcoef=vlm.corrcoef, corrcoef_random=vlm.corrcoef_random,
               transition_prob=vlm.transition_prob, transition_prob_random=vlm.transition_prob_random,
               delta_embedding=vlm.delta_embedding)
    # expression scaling (analysis.py:1714-1731) on the same correlations
    vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=True, scaling_penalty=1.0)
    res.update(scaling=vlm.scaling, scaling_rndm=vlm.scaling_rndm, delta_embedding_scaled=vlm.delta_embedding,
               delta_embedding_random_scaled=vlm.delta_embedding_random)
    # transform="logratio" (analysis.py:1582-1590)
    vlm.estimate_transition_prob(hidim="Sx_sz", embed="ts", transform="logratio", psc=1, n_neighbors=30,
                                 knn_random=True, sampled_fraction=0.5, n_jobs=1, threads=1, calculate_randomized=False)
    res.update(logratio_neigh_ixs=vlm.embedding_knn.indices.reshape(C, -1).copy(), logratio_corrcoef=vlm.corrcoef)
    # full (knn_random=False) mode on the same object
    vlm.estimate_transition_prob(hidim="Sx_sz", embed="ts", transform="sqrt", psc=1, n_neighbors=30,
                                 knn_random=False, calculate_randomized=False, n_jobs=1, threads=1)
    vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=False)
    res.update(full_knn_indices=vlm.embedding_knn.indices.reshape(C, -1).copy(),
               full_corrcoef=vlm.corrcoef, full_transition_prob=vlm.transition_prob,
               full_delta_embedding=vlm.delta_embedding)
    np.savez_compressed(os.path.join(out, "pipeline_small.npz"), **res)


def golden_pipeline_medium(an, out):
    """The whole tutorial chain (doc/tutorial/analysis.rst:108-165) through the reference's OWN methods at a size where
    fp32 storage roundings average out (>= 1000 genes): normalize -> perform_PCA (scikit-learn) -> knn_imputation ->
    fit_gammas -> predict_U -> calculate_velocity -> calculate_shift -> extrapolate_cell_at_t ->
    estimate_transition_prob (psc = 1 as in the tutorial, and the default psc = 1e-10) -> calculate_embedding_shift ->
    calculate_grid_arrows.  Stored compactly: inputs, per-gene vectors, per-cell x neighbour matrices, a 40-gene
    slice of the big matrices."""
    if an is None:
        print("skipping medium pipeline golden (analysis.py not importable)")
        return
    import warnings
    G, C = 1400, 260
    S, U = synth_counts(G, C, 31)
    keep = (S.sum(1) > 3) & (U.sum(1) > 3)
    S, U = S[keep][:1100], U[keep][:1100]
    G = S.shape[0]
    vlm = an.VelocytoLoom.__new__(an.VelocytoLoom)
    vlm.S, vlm.U, vlm.A = S, U, np.zeros_like(S)
    vlm.ca, vlm.ra = {"CellID": np.arange(C)}, {"Gene": np.arange(G)}
    vlm.initial_cell_size = S.sum(0)
    vlm.initial_Ucell_size = U.sum(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vlm.normalize("both", size=True, log=True)
    vlm.perform_PCA(n_components=12)
    res = {"S": S, "U": U, "pcs": vlm.pcs.copy(), "pca_explained_variance_ratio": vlm.pca.explained_variance_ratio_.copy(),
           "pca_explained_variance": vlm.pca.explained_variance_.copy(), "pca_components_head": vlm.pca.components_[:, :60].copy()}
    vlm.knn_imputation(k=25, pca_space=True, n_pca_dims=10, balanced=False, n_jobs=1)
    res.update(knn_data=vlm.knn.data, knn_indices=vlm.knn.indices, knn_indptr=vlm.knn.indptr)
    for a in ("Sx", "Ux", "Sx_sz", "Ux_sz"):                   # C-contiguous, as a gene filter leaves them
        setattr(vlm, a, np.ascontiguousarray(getattr(vlm, a)))
    res["Sx_sz_head"], res["Ux_sz_head"] = vlm.Sx_sz[:40].copy(), vlm.Ux_sz[:40].copy()
    vlm.fit_gammas(weighted=False, fit_offset=False)
    res["gammas_nnls"] = vlm.gammas.copy()
    vlm.fit_gammas()                                           # default: maxmin_diag weights + offset (L-BFGS-B iterate)
    res["gammas_default"], res["q_default"] = vlm.gammas.copy(), vlm.q.copy()
    vlm.fit_gammas(weighted=False, fit_offset=True)            # the chain continues from the OLS fit (parity-able to 1e-5)
    res["gammas_ols"], res["q_ols"] = vlm.gammas.copy(), vlm.q.copy()
    vlm.predict_U()
    vlm.calculate_velocity()
    vlm.calculate_shift(assumption="constant_velocity")
    vlm.extrapolate_cell_at_t(delta_t=1.0)
    res.update(Upred_head=vlm.Upred[:40].copy(), velocity_head=vlm.velocity[:40].copy(), delta_S_head=vlm.delta_S[:40].copy(),
               Sx_sz_t_head=vlm.Sx_sz_t[:40].copy())
    vlm.ts = vlm.pcs[:, :2].copy()
    rows = np.arange(C)[:, None]
    for tag, psc in (("psc1", 1), ("pscdef", None)):
        vlm.estimate_transition_prob(hidim="Sx_sz", embed="ts", transform="sqrt", psc=psc, n_neighbors=80,
                                     knn_random=True, sampled_fraction=0.5, n_jobs=1, threads=1)
        vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=True)
        nb_ix = vlm.embedding_knn.indices.reshape(C, -1).copy()
        res.update({f"{tag}_neigh_ixs": nb_ix, f"{tag}_sampling_ixs": vlm.sampling_ixs.copy(),
                    f"{tag}_corrcoef": vlm.corrcoef[rows, nb_ix].copy(), f"{tag}_corrcoef_random": vlm.corrcoef_random[rows, nb_ix].copy(),
                    f"{tag}_transition_prob": vlm.transition_prob[rows, nb_ix].copy(),
                    f"{tag}_transition_prob_random": vlm.transition_prob_random[rows, nb_ix].copy(),
                    f"{tag}_delta_embedding": vlm.delta_embedding.copy(), f"{tag}_delta_embedding_random": vlm.delta_embedding_random.copy(),
                    f"{tag}_scaling": vlm.scaling.copy()})
    vlm.delta_ts, vlm.delta_ts_random = vlm.delta_embedding, vlm.delta_embedding_random     # what the tutorial's plots read
    vlm.calculate_grid_arrows(embed="ts", smooth=0.8, steps=(14, 11), n_neighbors=40, n_jobs=1)
    res.update(flow_grid=vlm.flow_grid, flow=vlm.flow, total_p_mass=vlm.total_p_mass, flow_norm=vlm.flow_norm,
               flow_norm_magnitude=vlm.flow_norm_magnitude, flow_rndm=vlm.flow_rndm, flow_norm_rndm=vlm.flow_norm_rndm)
    np.savez_compressed(os.path.join(out, "pipeline_medium.npz"), **res)


def golden_normalize(an, out):
    """The size/log normalisation family (analysis.py:535-676) on hand-built objects."""
    if an is None:
        print("skipping normalize golden (analysis.py not importable)")
        return
    G, C = 40, 31
    S, U = synth_counts(G, C, 21)
    U[:, 4] = 0.0                                   # a cell without unspliced molecules: 0/0 -> guard (analysis.py:581)
    rng = np.random.default_rng(22)
    Sx = S + rng.uniform(0, 1, S.shape)
    Ux = U + rng.uniform(0, 1, U.shape)

    def fresh():
        vlm = an.VelocytoLoom.__new__(an.VelocytoLoom)
        vlm.S, vlm.U, vlm.A = S.copy(), U.copy(), np.zeros_like(S)
        vlm.Sx, vlm.Ux = Sx.copy(), Ux.copy()
        vlm.ca, vlm.ra = {"CellID": np.arange(C)}, {"Gene": np.arange(G)}
        return vlm

    res = {"S": S, "U": U, "Sx": Sx, "Ux": Ux}
    cases = {"default": dict(which="both"),
             "opts": dict(which="both", pcount=0.5, use_S_size_for_U=True, target_size=(1000.0, 500.0)),
             "nosize": dict(which="both", size=False),
             "nolog": dict(which="both", log=False),
             "relsize": dict(which="both", relative_size=np.linspace(50.0, 400.0, C)),
             "imputed": dict(which="imputed"),
             "imputed_opts": dict(which="imputed", pcount=2.0, use_S_size_for_U=True, target_size=(800.0, None))}
    import warnings
    for tag, kw in cases.items():
        vlm = fresh()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                vlm.normalize(**kw)
        for name in ("S_sz", "S_norm", "U_sz", "U_norm", "Sx_sz", "Sx_norm", "Ux_sz", "Ux_norm", "cell_size", "avg_size",
                     "norm_factor", "Ucell_size", "Uavg_size", "Unorm_factor", "xcell_size", "xavg_size", "xnorm_factor",
                     "xUcell_size", "xUavg_size", "xUnorm_factor"):
            if hasattr(vlm, name):
                res[f"{tag}__{name}"] = np.asarray(getattr(vlm, name), dtype=np.float64)
    np.savez_compressed(os.path.join(out, "normalize_small.npz"), **res)


def main():
    sb, est, nb, an = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "normalize":     # add one fixture without touching the others
        golden_normalize(an, HERE)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "medium":
        golden_pipeline_medium(an, HERE)
        return
    golden_normalize(an, HERE)
    golden_coldeltacor(est, HERE)
    golden_fits(est, an, HERE)
    golden_smoothing(nb, HERE)
    golden_pipeline(an, HERE)
    golden_pipeline_medium(an, HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
