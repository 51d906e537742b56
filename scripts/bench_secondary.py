"""Secondary measurements: the other kernels of the path at BASELINE config 2
(10k cells x 20k genes, k = 500) and the full (all-pairs) correlation at a reduced config-3 shape.
Prints one JSON object per kernel: CUDA-event time, algorithmic bytes, achieved GB/s vs the measured HBM peak.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from velocyto_b200 import _cabi, device as dev

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

def timeit(fn, iters=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def report(name, ms, alg_bytes, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel=name, ms=ms, algorithmic_bytes=alg_bytes, achieved_gbs=gbs, frac_of_measured_hbm=gbs / peak, **extra)))

def main():
    C, G, k = 10_000, 20_000, 500
    gen = torch.Generator(device="cuda").manual_seed(0)
    S, U = dev.CellMajor.empty(C, G), dev.CellMajor.empty(C, G)
    S.t[:, :G] = torch.poisson(torch.rand((C, G), device="cuda", generator=gen) * 2)
    U.t[:, :G] = torch.poisson(torch.rand((C, G), device="cuda", generator=gen))
    ld = S.ld
    # K5: kNN smoothing, k+1 entries per row
    idx = ((torch.arange(C, device="cuda")[:, None] + torch.randint(1, C, (C, k), device="cuda", generator=gen)) % C)
    idx = torch.cat([torch.arange(C, device="cuda")[:, None], idx], 1).to(torch.int32).contiguous()
    indptr = torch.arange(0, C * (k + 1) + 1, k + 1, device="cuda", dtype=torch.int64)
    w = torch.full((C * (k + 1),), 1.0 / (k + 1), device="cuda", dtype=torch.float32)
    ms = timeit(lambda: dev.knn_smooth(indptr, idx.view(-1), w, S))
    report("k_knn_smooth (K5, knn_imputation SpMM, k=500)", ms, C * (k + 1) * G * 4 + C * G * 4, shape=f"{C}x{G}")
    Sx, Ux = dev.knn_smooth(indptr, idx.view(-1), w, S), dev.knn_smooth(indptr, idx.view(-1), w, U)
    # K4: fits
    ms = timeit(lambda: dev.fit_gammas(dev.FIT_SLOPE, Sx, Ux))
    report("k_gene_moments+finalize (K4, fit_slope nnls)", ms, 2 * C * G * 4, shape=f"{C}x{G}")
    ms = timeit(lambda: dev.fit_gammas(dev.FIT_SLOPE_OFFSET, Sx, Ux))
    report("K4 fit_slope_offset (OLS)", ms, 2 * C * G * 4, shape=f"{C}x{G}")
    ms_w = timeit(lambda: dev.fit_weights("maxmin_diag", Sx, Ux, Sx, Ux), iters=3, warm=1)
    report("fit_weights maxmin_diag (radix-select percentiles)", ms_w, 3 * 5 * C * G * 4 + 6 * C * G * 4, shape=f"{C}x{G}")
    W = dev.fit_weights("maxmin_diag", Sx, Ux, Sx, Ux)
    ms = timeit(lambda: dev.fit_gammas(dev.FIT_SLOPE_WEIGHTED_OFFSET, Sx, Ux, W, lo=1e-8, hi=20.0, want_r2=True))
    report("K4 fit_slope_weighted_offset (fit_gammas default)", ms, 3 * C * G * 4, shape=f"{C}x{G}")
    g, q, _, _ = dev.fit_gammas(dev.FIT_SLOPE_OFFSET, Sx, Ux)
    # K6: chain
    ms = timeit(lambda: dev.velocity_chain(Sx, Ux, g, q, transform="sqrt", psc=1.0))
    report("k_velocity_chain (K6, 5 outputs)", ms, 7 * C * G * 4, shape=f"{C}x{G}")
    ms = timeit(lambda: dev.velocity_chain(Sx, Ux, g, q, transform="sqrt", psc=1.0, want=("d",)))
    report("k_velocity_chain (K6, d only)", ms, 3 * C * G * 4, shape=f"{C}x{G}")
    ms = timeit(lambda: dev.velocity_chain(Sx, Ux, g, q, transform="sqrt", psc=1.0, eps=0.05, want=("velocity",)))
    report("velocity threshold (k_gene_max_upred) + k_velocity_chain (velocity only, eps)", ms, 4 * C * G * 4, shape=f"{C}x{G}")
    # normalize family
    sums = dev.cell_sums(S)
    fac = (sums.mean() / sums).contiguous()
    ms = timeit(lambda: dev.size_normalize(S, fac, 1.0))
    report("k_size_normalize (S_sz + S_norm in one pass)", ms, 3 * C * G * 4, shape=f"{C}x{G}")
    ms = timeit(lambda: dev.cell_sums(S))
    report("k_cell_sums", ms, C * G * 4, shape=f"{C}x{G}")
    # layout converters
    src = torch.rand((4096, C), device="cuda", dtype=torch.float64)
    dst = dev.CellMajor.empty(C, 4096)
    ms = timeit(lambda: _cabi.call("velo_dev_pack_cellmajor", src.data_ptr(), 8, 4096, C, dst.ptr, dst.ld, 0, torch.cuda.current_stream().cuda_stream))
    report("k_pack_cellmajor<double>", ms, 4096 * C * 12, shape=f"4096x{C}")
    # exact kNN on the device (replaces scikit-learn's search)
    pts = torch.randn((100_000, 2), device="cuda", dtype=torch.float64, generator=gen)
    t0 = time.perf_counter(); dev.knn(pts, 3000); torch.cuda.synchronize(); t1 = time.perf_counter()
    t0 = time.perf_counter(); dev.knn(pts, 3000); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(json.dumps(dict(kernel="k_knn_bruteforce 100000 points x 2-D, k=3000 (embedding neighbourhoods)", ms=(t1 - t0) * 1e3)))
    pcs = torch.randn((C, 20), device="cuda", dtype=torch.float64, generator=gen)
    dev.knn(pcs, k); torch.cuda.synchronize()
    t0 = time.perf_counter(); dev.knn(pcs, k); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(json.dumps(dict(kernel=f"k_knn_bruteforce {C} points x 20-D, k={k} (knn_imputation search)", ms=(t1 - t0) * 1e3)))
    del pts, pcs
    # device-side randomisation at BASELINE config 4 scale (opt-in random_backend="device")
    Cs, W, m = 100_000, 10_001, 3_000
    knn_idx = torch.randint(0, Cs, (Cs, W), device="cuda", dtype=torch.int32, generator=gen)
    p = np.linspace(0.5, 0.1, W); p /= p.sum()
    ms = timeit(lambda: dev.sample_neighbors(knn_idx, p, m, 1), iters=3, warm=1)
    print(json.dumps(dict(kernel=f"k_sample_neighbors {Cs} cells, {m} of {W} candidates (replaces the per-cell np.random.choice loop)", ms=ms)))
    del knn_idx
    Gs = 30_000
    dS = dev.CellMajor.empty(Cs // 2, Gs)
    dS.t.normal_(generator=gen)
    ms = timeit(lambda: dev.permute_rows_nsign(dS, 1), iters=3, warm=1)
    report(f"k_permute_rows_nsign {Cs // 2} cells x {Gs} genes (randomised control)", ms, 2 * (Cs // 2) * Gs * 4)
    del dS
    # full (all pairs) correlation, reduced config 3: G=30k, C=4k -> G*C^2 = 4.8e11 elements
    del S, U, Ux, W, src, dst
    Cf, Gf = 4000, 30_000
    e = dev.CellMajor.empty(Cf, Gf); d = dev.CellMajor.empty(Cf, Gf)
    e.t[:, :Gf] = torch.rand((Cf, Gf), device="cuda", generator=gen) * 3
    d.t[:, :Gf] = torch.randn((Cf, Gf), device="cuda", generator=gen)
    for tr in ("sqrt", "linear"):
        ms = timeit(lambda: dev.coldeltacor(e, d, None, tr, 1.0), iters=2, warm=1)
        print(json.dumps(dict(kernel=f"k_coldeltacor full {tr} (config 3 reduced: {Cf} cells x {Gf} genes)", ms=ms,
                              elements=Gf * Cf * Cf, gelem_per_s=Gf * Cf * Cf / (ms * 1e-3) / 1e9,
                              cells_per_s=Cf / (ms * 1e-3), l2_bytes=Cf * Cf * Gf * 4)))

if __name__ == "__main__":
    main()
