"""Wall-clock of the VelocytoLoom hot methods through the public mirror API at BASELINE config 2
(10k cells x 20k genes, k = 500), host arrays in, attributes out -- SURVEY.md 8(d) secondary measurements."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from velocyto_b200.analysis import VelocytoLoom
from velocyto_b200 import _cabi

C, G, k = int(os.environ.get("C", 10_000)), int(os.environ.get("G", 20_000)), 500
rng = np.random.default_rng(0)
mu = rng.gamma(0.6, 2.0, G)[:, None]; sc = rng.gamma(2.0, 0.5, C)[None, :]
S = rng.poisson(mu * sc).astype(np.float64)
U = rng.poisson(mu * sc * rng.uniform(0.05, 1.0, (G, 1))).astype(np.float64)
vlm = VelocytoLoom(S=S, U=U)
vlm.S_sz = S / np.maximum(S.sum(0), 1) * S.sum(0).mean()
vlm.U_sz = U / np.maximum(U.sum(0), 1) * U.sum(0).mean()
Sn = np.log2(vlm.S_sz[:2000] + 1); Sn -= Sn.mean(1)[:, None]
u, s, _ = np.linalg.svd(Sn.T, full_matrices=False)
vlm.pcs = u[:, :25] * s[:25]
vlm.ts = vlm.pcs[:, :2].copy()
res, launches0 = {}, _cabi.launch_count()

def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    res[name] = round((time.perf_counter() - t0) * 1e3, 2)

timed("knn_imputation(k=500, n_pca_dims=20) [upload + device kNN + K5 x2]", lambda: vlm.knn_imputation(k=k, n_pca_dims=20, n_jobs=8))
timed("knn_imputation again (matrices resident)", lambda: vlm.knn_imputation(k=k, n_pca_dims=20, n_jobs=8))
timed("fit_gammas() default (weights + box-constrained fit)", lambda: vlm.fit_gammas())
timed("fit_gammas(weighted=False, fit_offset=False)", lambda: vlm.fit_gammas(weighted=False, fit_offset=False))
vlm.fit_gammas(weighted=False, fit_offset=True)
timed("predict_U + calculate_velocity + calculate_shift + extrapolate_cell_at_t", lambda: (vlm.predict_U(), vlm.calculate_velocity(), vlm.calculate_shift(), vlm.extrapolate_cell_at_t(delta_t=1.0)))
timed("estimate_transition_prob(sqrt, n_neighbors=2000, frac 0.3, randomized) [device kNN + host sampler + K1 x2]",
      lambda: vlm.estimate_transition_prob(hidim="Sx_sz", embed="ts", transform="sqrt", psc=1, n_neighbors=2000, knn_random=True,
                                           sampled_fraction=0.3, n_jobs=8))
timed("calculate_embedding_shift(expression_scaling=True)", lambda: vlm.calculate_embedding_shift(sigma_corr=0.05, expression_scaling=True))
timed("read back Sx_sz as float64 (genes x cells)", lambda: vlm.Sx_sz)
res["library_kernel_launches"] = _cabi.launch_count() - launches0
res["shape"] = f"{C} cells x {G} genes"
print(json.dumps(res, indent=1))
