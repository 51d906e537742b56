"""One launch (after one warm-up) of every secondary kernel of the path at BASELINE config-2 shapes, for an ncu metrics pass:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
        --clock-control none --csv --log-file gpurun_out/secondary_ncu.csv python scripts/ncu_secondary.py

Numbers under ncu are not bench values (scripts/bench_secondary.py times the same kernels with CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from velocyto_b200 import _cabi, device as dev  # noqa: E402


def main():
    C, G, k = 10_000, 20_000, 500
    gen = torch.Generator(device="cuda").manual_seed(0)
    S, U = dev.CellMajor.empty(C, G), dev.CellMajor.empty(C, G)
    S.t[:, :G] = torch.poisson(torch.rand((C, G), device="cuda", generator=gen) * 2)
    U.t[:, :G] = torch.poisson(torch.rand((C, G), device="cuda", generator=gen))
    idx = ((torch.arange(C, device="cuda")[:, None] + torch.randint(1, C, (C, k), device="cuda", generator=gen)) % C)
    idx = torch.cat([torch.arange(C, device="cuda")[:, None], idx], 1).to(torch.int32).contiguous()
    indptr = torch.arange(0, C * (k + 1) + 1, k + 1, device="cuda", dtype=torch.int64)
    w = torch.full((C * (k + 1),), 1.0 / (k + 1), device="cuda", dtype=torch.float32)
    for rep in range(2):                                   # profile with `-s <launches of rep 0>` or read the second half
        sums = dev.cell_sums(S)                                                        # normalize family
        fac = (sums.mean() / sums).contiguous()
        S_sz, S_norm = dev.size_normalize(S, fac, 1.0)
        Sx = dev.knn_smooth(indptr, idx.view(-1), w, S)                                 # K5
        Ux = dev.knn_smooth(indptr, idx.view(-1), w, U)
        g, q, _, _ = dev.fit_gammas(dev.FIT_SLOPE_OFFSET, Sx, Ux)                       # K4
        W = dev.fit_weights("maxmin_diag", Sx, Ux, Sx, Ux)                              # percentile weights
        dev.fit_gammas(dev.FIT_SLOPE_WEIGHTED_OFFSET, Sx, Ux, W, lo=1e-8, hi=20.0, want_r2=True)
        chain = dev.velocity_chain(Sx, Ux, g, q, transform="sqrt", psc=1.0)            # K6
        del S_sz, S_norm, W, chain
        torch.cuda.synchronize()
    del S, U, Sx, Ux
    # all-pairs kernels at a reduced config-3 shape
    Cf, Gf = 4096, 30_000
    e, d = dev.CellMajor.empty(Cf, Gf), dev.CellMajor.empty(Cf, Gf)
    e.t[:, :Gf] = torch.rand((Cf, Gf), device="cuda", generator=gen) * 3
    d.t[:, :Gf] = torch.randn((Cf, Gf), device="cuda", generator=gen)
    for rep in range(2):
        corr = dev.coldeltacor(e, d, None, "sqrt", 1.0)                                 # K2 (MUFU bound)
        corr = dev.coldeltacor(e, d, None, "linear", 0.0)                               # K2g (tensor cores)
        dev.transition_prob(corr, None, 0.05)
        dev.permute_rows_nsign(d, 1)
        torch.cuda.synchronize()
    knn_idx = torch.randint(0, 20_000, (20_000, 10_001), device="cuda", dtype=torch.int32, generator=gen)
    p = np.linspace(0.5, 0.1, 10_001)
    for rep in range(2):
        dev.sample_neighbors(knn_idx, p / p.sum(), 3000, 1)
        torch.cuda.synchronize()
    print("launches:", _cabi.launch_count())


if __name__ == "__main__":
    main()
