"""Wall-clock of the VelocytoLoom hot methods through the public mirror API, host arrays in, attributes out, next to the
device time of the kernels each method launches (SURVEY.md 8(d) secondary measurements).

    python scripts/bench_pipeline.py                 # BASELINE config 2: 10k cells x 20k genes, k = 500
    C=100000 G=5000 K=500 NN=10000 python scripts/bench_pipeline.py     # 100k cells (post-filter gene count)

For every method: wall-clock (synchronised) of the call, and the device-busy time of the same call measured with CUDA
events around it (an upper bound of its kernel time: it includes gaps).  ratio = wall / device."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from velocyto_b200.analysis import VelocytoLoom
from velocyto_b200 import _cabi

C, G, k = int(os.environ.get("C", 10_000)), int(os.environ.get("G", 20_000)), int(os.environ.get("K", 500))
NN = int(os.environ.get("NN", 2000))
rng = np.random.default_rng(0)
mu = rng.gamma(0.6, 2.0, G)[:, None].astype(np.float32); sc = rng.gamma(2.0, 0.5, C)[None, :].astype(np.float32)
t0 = time.perf_counter()
S = rng.poisson(mu * sc).astype(np.float64)
U = rng.poisson(mu * sc * rng.uniform(0.05, 1.0, (G, 1))).astype(np.float64)
print(f"synthetic counts {G} x {C}: {time.perf_counter() - t0:.1f} s", file=sys.stderr)
vlm = VelocytoLoom(S=S, U=U)
res, launches0 = {}, _cabi.launch_count()


def timed(name, fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _cabi.launch_count()
    t0 = time.perf_counter(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    res[name] = {"wall_ms": round(wall, 2), "cuda_event_ms": round(a.elapsed_time(b), 2), "library_launches": _cabi.launch_count() - l0}


timed("normalize('both') [first use: H2D of S and U]", lambda: vlm.normalize("both", size=True, log=True))
timed("normalize('both') again (resident)", lambda: vlm.normalize("both", size=True, log=True))
timed("perform_PCA(n_components=25) [first call: cuSOLVER / cuBLAS fp64 initialisation]", lambda: vlm.perform_PCA(n_components=25))
timed("perform_PCA(n_components=25) again", lambda: vlm.perform_PCA(n_components=25))
timed(f"knn_imputation(k={k}, n_pca_dims=20) [device kNN + K5 x2]", lambda: vlm.knn_imputation(k=k, n_pca_dims=20, n_jobs=8))
timed("knn_imputation again", lambda: vlm.knn_imputation(k=k, n_pca_dims=20, n_jobs=8))
timed("fit_gammas() default (weights + box-constrained fit) [first call]", lambda: vlm.fit_gammas())
timed("fit_gammas() default again", lambda: vlm.fit_gammas())
timed("fit_gammas(weighted=False, fit_offset=False)", lambda: vlm.fit_gammas(weighted=False, fit_offset=False))
vlm.fit_gammas(weighted=False, fit_offset=True)
timed("predict_U + calculate_velocity + calculate_shift + extrapolate_cell_at_t",
      lambda: (vlm.predict_U(), vlm.calculate_velocity(), vlm.calculate_shift(), vlm.extrapolate_cell_at_t(delta_t=1.0)))
vlm.ts = vlm.pcs[:, :2].copy()
kw = dict(hidim="Sx_sz", embed="ts", transform="sqrt", psc=1, n_neighbors=NN, knn_random=True, sampled_fraction=0.3, n_jobs=8)
timed(f"estimate_transition_prob(n_neighbors={NN}, frac 0.3, randomized) random_backend='device' [first call]",
      lambda: vlm.estimate_transition_prob(random_backend="device", **kw))
if os.environ.get("PROFILE", "0") == "1":
    import cProfile, pstats, io
    pr = cProfile.Profile()
    pr.enable()
    vlm.estimate_transition_prob(random_backend="device", **kw)
    torch.cuda.synchronize()
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(18)
    print(buf.getvalue(), file=sys.stderr)
timed(f"estimate_transition_prob(n_neighbors={NN}, frac 0.3, randomized) random_backend='device' again",
      lambda: vlm.estimate_transition_prob(random_backend="device", **kw))
timed("calculate_embedding_shift(expression_scaling=True) [after device backend]",
      lambda: vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=True))
if os.environ.get("SKIP_REFERENCE_RNG", "0") != "1":
    timed(f"estimate_transition_prob(n_neighbors={NN}, frac 0.3, randomized) default: reference RNG streams (C++ sampler + numba shuffle)",
          lambda: vlm.estimate_transition_prob(**kw))
    timed("calculate_embedding_shift(expression_scaling=True)", lambda: vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=True))
vlm.delta_ts, vlm.delta_ts_random = vlm.delta_embedding, vlm.delta_embedding_random
timed("calculate_grid_arrows(steps=(40, 40), n_neighbors=100)", lambda: vlm.calculate_grid_arrows(embed="ts", smooth=0.5, steps=(40, 40), n_neighbors=100))
timed("read back Sx_sz as float64 (genes x cells)", lambda: vlm.Sx_sz)
if C <= 20000:
    timed("read corrcoef (dense cells x cells, built lazily)", lambda: vlm.corrcoef)
out = {"shape": f"{C} cells x {G} genes, k={k}, n_neighbors={NN}", "methods": res, "library_kernel_launches": _cabi.launch_count() - launches0}
print(json.dumps(out, indent=1))
