"""BASELINE config 5: 500k cells x 30k genes, sparse CSR counts, knn_imputation SpMM + per-gene gamma fit,
gene-sharded across the GPUs of one box, followed by the hand-over to the cell-sharded correlation.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/bench_config5.py [--cells 500000 --genes 30000 --density 0.05 --k 500]
    python scripts/bench_config5.py --check          # reduced shape, one or more GPUs: bit-equality of the shards

Per rank (one process per GPU; torch.distributed / NCCL is only the plumbing):
  1. kNN search in PCA space (exact brute force, `velo_dev_knn_range`): QUERY-sharded, index blocks all-gathered
     (the reference runs scikit-learn on one host, neighbors.py:363-376) -> smoothing weights 1/(k+1), diag = 1
     (analysis.py:1006-1010);
  2. K5 on sparse counts, `velo_dev_knn_smooth_csr` (neighbors.py:416-423): the rank's GENE slab of S and U, CSR by
     cell in, dense cell-major slab out -- no collective (every gene is independent);
  3. K4: "maxmin_diag" weights + the default weighted slope/offset fit (analysis.py:1196-1256, estimation.py:337-366)
     and the plain nnls slope (estimation.py:267-279) on the slab -- no collective;
  4. K6: velocity chain -> transformed velocity d on the slab;
  5. gene blocks -> cell blocks (`sharding.genes_to_cells`, one all-to-all) for Sx_sz and d;
  6. all-gather of the expression blocks + K1 (`colDeltaCorSqrtpartial`, m neighbours) on the first --k1-cells cells
     of the rank's block (the full 500k-cell correlation is 5x config 4; a bounded sample is timed and labelled).

Prints one JSON line (rank 0): per-stage ms = max over ranks of CUDA-event times, algorithmic bytes by SURVEY.md 8(d)'s
formulas, achieved GB/s per GPU, peak device memory per rank, and the properties checked at full shape.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=500_000)
    ap.add_argument("--genes", type=int, default=30_000)
    ap.add_argument("--genes-per-rank", type=int, default=0, help="weak scaling: genes = this x world")
    ap.add_argument("--density", type=float, default=0.05)
    ap.add_argument("--k", type=int, default=500)
    ap.add_argument("--pcs", type=int, default=20)
    ap.add_argument("--m", type=int, default=3000)
    ap.add_argument("--k1-cells", type=int, default=2072, help="cells per rank timed through K1 (multiple of 148)")
    ap.add_argument("--no-k1", action="store_true")
    ap.add_argument("--check", action="store_true", help="reduced shape: sharded result == single-GPU result, bit for bit")
    return ap.parse_args()


class Timer:
    def __init__(self):
        self.ev, self.ms = {}, {}

    def __call__(self, name):
        t = self

        class _Ctx:
            def __enter__(self_):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t.ev[name] = (a, b)
                a.record()

            def __exit__(self_, *exc):
                t.ev[name][1].record()
        return _Ctx()

    def collect(self, world):
        torch.cuda.synchronize()
        names = list(self.ev)
        v = torch.tensor([self.ev[n][0].elapsed_time(self.ev[n][1]) for n in names], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return {n: float(x) for n, x in zip(names, v)}


def synth_pcs(C, D, device):
    """Cells along a closed latent trajectory (what makes real kNN graphs local) + isotropic noise; same on all ranks."""
    gen = torch.Generator(device=device).manual_seed(7)
    t = torch.arange(C, device=device, dtype=torch.float64) * (2 * np.pi / C)
    freqs = torch.arange(1, D + 1, device=device, dtype=torch.float64)
    X = torch.cos(t[:, None] * freqs[None, :] + freqs[None, :]) * (10.0 / freqs[None, :])
    return (X + 0.05 * torch.randn((C, D), device=device, dtype=torch.float64, generator=gen)).contiguous()


def synth_csr_slab(C, ng, density, seed, device, scale_u=1.0):
    """CSR by cell of a (C x ng) count slab: Bernoulli(density) pattern, geometric counts, per-cell size factors
    (already 'size-normalised': S_sz).  Gene ids are slab-local, sorted within a row."""
    gen = torch.Generator(device=device).manual_seed(seed)
    size = 0.5 + torch.rand(C, device=device, generator=gen)
    counts = torch.empty(C, dtype=torch.int64, device=device)
    gs, vs = [], []
    blk = max(1, (1 << 27) // max(1, ng))
    for c0 in range(0, C, blk):
        n = min(blk, C - c0)
        mask = torch.rand((n, ng), device=device, generator=gen) < density
        counts[c0:c0 + n] = mask.sum(1)
        nz = mask.nonzero()                                            # row-major: gene ids sorted within a cell
        v = (1.0 + torch.floor(-1.5 * torch.log(torch.rand(nz.shape[0], device=device, generator=gen).clamp_min(1e-7))))
        vs.append((v * size[c0 + nz[:, 0]] * scale_u).to(torch.float32))
        gs.append(nz[:, 1].to(torch.int32))
        del mask, nz, v
    indptr = torch.zeros(C + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, torch.cat(gs), torch.cat(vs)


def csr_filter_genes(indptr, genes, vals, g0, ng):
    """Column slab [g0, g0+ng) of a CSR-by-cell matrix (check mode: every rank cuts its slab out of the same matrix)."""
    keep = (genes >= g0) & (genes < g0 + ng)
    rows = torch.repeat_interleave(torch.arange(indptr.numel() - 1, device=genes.device), indptr[1:] - indptr[:-1])
    cnt = torch.zeros(indptr.numel() - 1, dtype=torch.int64, device=genes.device).index_add_(0, rows[keep], torch.ones_like(rows[keep]))
    ip = torch.zeros_like(indptr)
    ip[1:] = torch.cumsum(cnt, 0)
    return ip, (genes[keep] - g0).to(torch.int32), vals[keep]


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from velocyto_b200 import _cabi, device as dev
    from velocyto_b200.sharding import (CellShardedTransitionProb, gene_partition, genes_to_cells, partition, block_size)
    device = torch.device("cuda", lr)
    if args.check:
        args.cells, args.genes, args.k, args.m, args.k1_cells = 3001, 4100, 40, 64, 0
    C, D, k, m, psc = args.cells, args.pcs, args.k, args.m, 1.0
    G = args.genes_per_rank * world if args.genes_per_rank else args.genes
    g0, ng = gene_partition(G, world)[rank]
    c0, nc = partition(C, world)[rank]
    T = Timer()
    torch.cuda.reset_peak_memory_stats()
    launches0 = _cabi.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one-time costs that are not part of any stage: NCCL connection set-up (all-gather ring + the pairwise channels
    # of the all-to-all), first-touch growth of the library's workspace pool
    if world > 1:
        tiny = torch.zeros((world * 8, 32), device=device)
        dist.all_gather_into_tensor(tiny, tiny[rank * 8:(rank + 1) * 8].clone())
        genes_to_cells(torch.zeros((world * 4, max(1, gene_partition(64 * world, world)[rank][1])), device=device), 64 * world)
    # ------------------------------------------------------------------ 1. kNN (query-sharded) -> smoothing weights
    pcs = synth_pcs(C, D, device)
    b = block_size(C, world)
    dev.knn(pcs[:4096].contiguous(), min(k, 64), want_dist=False)              # kernel load
    barrier()
    with T("knn_search"):
        idx_blk, _ = dev.knn(pcs, k, include_self=False, q0=c0, nq=nc, want_dist=False)
    with T("knn_allgather"):
        if world > 1:
            pad = idx_blk if nc == b else torch.cat([idx_blk, torch.zeros((b - nc, k), dtype=torch.int32, device=device)])
            idx_all = torch.empty((world * b, k), dtype=torch.int32, device=device)
            dist.all_gather_into_tensor(idx_all, pad.contiguous())
            idx_all = idx_all[:C]
        else:
            idx_all = idx_blk
    w_ix = torch.cat([torch.arange(C, device=device, dtype=torch.int32)[:, None], idx_all], 1).contiguous().view(-1)
    w_ip = torch.arange(0, C * (k + 1) + 1, k + 1, device=device, dtype=torch.int64)
    w_wt = torch.full((C * (k + 1),), 1.0 / (k + 1), device=device, dtype=torch.float32)
    del idx_blk, pcs

    # ------------------------------------------------------------------ synthetic sparse counts: this rank's gene slab
    if args.check:
        full_S = synth_csr_slab(C, G, args.density, 11, device)              # same matrix on every rank
        full_U = synth_csr_slab(C, G, args.density, 12, device, 0.5)
        S_csr, U_csr = csr_filter_genes(*full_S, g0, ng), csr_filter_genes(*full_U, g0, ng)
    else:
        S_csr = synth_csr_slab(C, ng, args.density, 100 + rank, device)
        U_csr = synth_csr_slab(C, ng, args.density, 200 + rank, device, 0.5)
    nnz_S, nnz_U = int(S_csr[1].numel()), int(U_csr[1].numel())

    # ------------------------------------------------------------------ 2. K5 on sparse counts (no collective)
    smooth = lambda csr: dev.knn_smooth_csr(w_ip, w_ix, w_wt, (csr[0], csr[1], csr[2], ng), g0=0, ng=ng)
    Sx = smooth(S_csr)                                                       # warm-up + determinism reference
    tmp = smooth(U_csr)                                                      # second warm-up: leaves a cached output block,
    del tmp                                                                  # so no cudaMalloc falls inside the timed stages
    tmp = smooth(S_csr)
    del tmp
    barrier()
    with T("k5_csr_S"):
        Sx2 = smooth(S_csr)
    deterministic = bool(torch.equal(Sx.t, Sx2.t))
    del Sx2
    with T("k5_csr_U"):
        Ux = smooth(U_csr)
    # column sums: sum_c Sx[c, g] == sum_j (sum_c w[c, j]) * S[j, g]   (size-independent property of the SpMM)
    win = torch.zeros(C, dtype=torch.float64, device=device).index_add_(0, w_ix.to(torch.int64), w_wt.double())
    rows = torch.repeat_interleave(torch.arange(C, device=device), S_csr[0][1:] - S_csr[0][:-1])
    want_cols = torch.zeros(ng, dtype=torch.float64, device=device).index_add_(0, S_csr[1].to(torch.int64), win[rows] * S_csr[2].double())
    got_cols = Sx.t[:, :ng].sum(0, dtype=torch.float64)
    colsum_rel = float(((got_cols - want_cols).abs() / want_cols.abs().clamp_min(1e-30)).max())
    del rows, win
    # a few cells recomputed in fp64 from the CSR rows of their neighbours
    spot = 0.0
    for c in (0, C // 2, C - 1):
        acc = torch.zeros(ng, dtype=torch.float64, device=device)
        for j in w_ix[c * (k + 1):(c + 1) * (k + 1)].tolist():
            a, bb = int(S_csr[0][j]), int(S_csr[0][j + 1])
            acc.index_add_(0, S_csr[1][a:bb].to(torch.int64), S_csr[2][a:bb].double())
        acc /= (k + 1)
        spot = max(spot, float((Sx.t[c, :ng].double() - acc).abs().max() / acc.abs().max().clamp_min(1e-30)))

    # ------------------------------------------------------------------ 3. K4 on the slab (no collective)
    W = dev.fit_weights("maxmin_diag", Sx, Ux, Sx, Ux)                         # warm-up: grows the workspace pool (15 GB)
    del W
    barrier()
    with T("fit_weights_maxmin_diag"):
        W = dev.fit_weights("maxmin_diag", Sx, Ux, Sx, Ux)
    with T("k4_fit_weighted_offset"):
        gam, q, r2, _ = dev.fit_gammas(dev.FIT_SLOPE_WEIGHTED_OFFSET, Sx, Ux, W, lo=1e-8, hi=20.0, want_r2=True)
    with T("k4_fit_slope_nnls"):
        gam0, _, _, _ = dev.fit_gammas(dev.FIT_SLOPE, Sx, Ux)
    del W
    gam = torch.nan_to_num(gam, nan=0.0, posinf=0.0, neginf=0.0)
    # nnls slope == max(0, sum xy / sum xx) per gene (closed form, fp64 on the device)
    sxx = torch.zeros(ng, dtype=torch.float64, device=device)
    sxy = torch.zeros(ng, dtype=torch.float64, device=device)
    for r0 in range(0, C, 65536):
        xb = Sx.t[r0:r0 + 65536, :ng].double()
        sxx += (xb * xb).sum(0)
        sxy += (xb * Ux.t[r0:r0 + 65536, :ng].double()).sum(0)
        del xb
    closed = torch.clamp(sxy / sxx, min=0)
    ok = torch.isfinite(closed) & torch.isfinite(gam0.double())
    nnls_rel = float(((gam0.double() - closed).abs()[ok] / closed[ok].clamp_min(1e-12)).max())
    del sxx, sxy

    # ------------------------------------------------------------------ 4. K6 chain on the slab
    barrier()
    with T("k6_chain_d"):
        d_slab = dev.velocity_chain(Sx, Ux, gam, q, transform="sqrt", psc=psc, want=("d",))["d"]
    del Ux
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ 5. gene blocks -> cell blocks (one all-to-all each)
    def to_cm(x):
        cm = dev.CellMajor.empty(x.shape[0], G)
        cm.t[:, :G] = x
        return cm

    barrier()
    with T("genes_to_cells_Sx"):
        e_rows = genes_to_cells(Sx.t[:, :ng].contiguous(), G)
    with T("genes_to_cells_d"):
        d_rows = genes_to_cells(d_slab.t[:, :ng].contiguous(), G)
    result = {}
    if args.check:
        # the gene-sharded pipeline must reproduce the single-GPU pipeline on the whole matrices, bit for bit
        Sx_full = dev.knn_smooth_csr(w_ip, w_ix, w_wt, (full_S[0], full_S[1], full_S[2], G))
        Ux_full = dev.knn_smooth_csr(w_ip, w_ix, w_wt, (full_U[0], full_U[1], full_U[2], G))
        Wf = dev.fit_weights("maxmin_diag", Sx_full, Ux_full, Sx_full, Ux_full)
        gf, qf, _, _ = dev.fit_gammas(dev.FIT_SLOPE_WEIGHTED_OFFSET, Sx_full, Ux_full, Wf, lo=1e-8, hi=20.0, want_r2=True)
        gf = torch.nan_to_num(gf, nan=0.0, posinf=0.0, neginf=0.0)
        d_full = dev.velocity_chain(Sx_full, Ux_full, gf, qf, transform="sqrt", psc=psc, want=("d",))["d"]
        same = (torch.equal(Sx.t[:, :ng], Sx_full.t[:, g0:g0 + ng]) and torch.equal(gam, gf[g0:g0 + ng])
                and torch.equal(e_rows, Sx_full.t[c0:c0 + nc, :G]) and torch.equal(d_rows, d_full.t[c0:c0 + nc, :G]))
        flag = torch.tensor([1 if same else 0], device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        result["sharded_equals_single_gpu"] = bool(flag.item())
    del Sx, d_slab, S_csr, U_csr
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ 6. all-gather(e) + K1 on a bounded sample of cells
    k1 = None
    if not args.no_k1 and (args.k1_cells > 0 or args.check):
        n1 = nc if args.check else min(args.k1_cells, nc)
        core = CellShardedTransitionProb(G, C, "sqrt", psc, 0.05)
        e_loc, d_loc = to_cm(e_rows), to_cm(d_rows[:n1])
        del e_rows, d_rows
        torch.cuda.empty_cache()
        win_w = min(C - 1, int(round(m / 0.3)))
        gen = torch.Generator(device=device).manual_seed(99 + rank)
        pick = torch.rand((n1, win_w), device=device, generator=gen).topk(m, dim=1).indices
        off = pick - win_w // 2
        off = off + (off >= 0).to(off.dtype)
        ix = ((torch.arange(c0, c0 + n1, device=device)[:, None] + off) % C).to(torch.int32).contiguous()
        del pick, off
        barrier()
        with T("allgather_e"):
            if world > 1:
                from velocyto_b200.sharding import gather_cell_blocks
                e_all = dev.CellMajor(gather_cell_blocks(e_loc.t, core.b), G)
            else:
                e_all = e_loc
        with T("k1_sample"):
            stats = dev.cell_stats(d_loc)
            corr = dev.coldeltacor(e_all, d_loc, ix, "sqrt", psc, c0=c0, stats=stats)
            tp = dev.transition_prob(corr, ix, 0.05, c0=c0, out=corr)
        rowsum_err = float((tp.sum(1, dtype=torch.float64) - 1).abs().max())
        k1 = {"cells_per_rank": n1, "m": m, "rowsum_err": rowsum_err}

    ms = T.collect(world)
    peak_torch = torch.cuda.max_memory_allocated()
    free, total = torch.cuda.mem_get_info()
    mem = torch.tensor([peak_torch, total - free], device=device, dtype=torch.float64)
    stats_t = torch.tensor([nnz_S, nnz_U, colsum_rel, spot, nnls_rel, 0.0 if deterministic else 1.0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
        mx = stats_t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats_t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = stats_t, stats_t
    if rank == 0:
        peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
        nnz_w = C * (k + 1)
        nbar_S, nbar_U = nnz_S / C, nnz_U / C                       # non-zeros per cell INSIDE one rank's slab
        # SURVEY.md 8(d), CSR input: nnz(w) * nnz-per-cell * 8 B read + C * G * 4 B written, per rank's slab
        b5_S = nnz_w * nbar_S * 8 + C * ng * 4
        b5_U = nnz_w * nbar_U * 8 + C * ng * 4
        out = {
            "config": f"BASELINE config 5: {C} cells x {G} genes CSR ({args.density:.0%} dense), k={k}, gene-sharded x{world}"
                      + (" [--check, reduced]" if args.check else ""),
            "n_gpus": world, "genes_per_rank": ng, "nnz_total_S": int(sm[0]), "nnz_total_U": int(sm[1]),
            "stage_ms_max_over_ranks": ms,
            "k5_csr": {"algorithmic_bytes_per_rank": [b5_S, b5_U],
                       "achieved_gbs_per_gpu": [b5_S / ms["k5_csr_S"] / 1e6, b5_U / ms["k5_csr_U"] / 1e6],
                       "frac_of_measured_hbm": [b5_S / ms["k5_csr_S"] / 1e6 / hbm, b5_U / ms["k5_csr_U"] / 1e6 / hbm],
                       "cells_per_s_whole_job_both_matrices": C / ((ms["k5_csr_S"] + ms["k5_csr_U"]) * 1e-3)},
            "k4": {"algorithmic_bytes_per_rank": 3 * C * ng * 4,
                   "achieved_gbs_per_gpu_weighted_offset": 3 * C * ng * 4 / ms["k4_fit_weighted_offset"] / 1e6,
                   "achieved_gbs_per_gpu_nnls": 2 * C * ng * 4 / ms["k4_fit_slope_nnls"] / 1e6,
                   "genes_per_s_whole_job_default_fit": G / ((ms["fit_weights_maxmin_diag"] + ms["k4_fit_weighted_offset"]) * 1e-3)},
            "exchange": {"genes_to_cells_bytes_per_rank": 2 * nc * G * 4, "knn_allgather_bytes": C * k * 4,
                         "allgather_e_bytes_received_per_rank": (C - nc) * G * 4 if k1 else 0},
            "k1_sample": None if k1 is None else dict(
                k1, ms=ms["k1_sample"], cells_per_s_per_gpu=k1["cells_per_rank"] / (ms["k1_sample"] * 1e-3),
                achieved_gbs_per_gpu=k1["cells_per_rank"] * (m + 2) * G * 4 / ms["k1_sample"] / 1e6,
                note="bounded sample of each rank's cells through colDeltaCorSqrtpartial + transition_prob with the "
                     "all-gathered 500k-cell expression matrix resident; the full job is C/N cells per rank"),
            "memory": {"peak_torch_allocated_gb_max_rank": float(mem[0]) / 1e9, "device_in_use_at_end_gb_max_rank": float(mem[1]) / 1e9},
            "properties": {"k5_deterministic": bool(mx[5] == 0), "k5_column_sum_identity_max_rel": float(mx[2]),
                           "k5_spot_cells_vs_fp64_max_rel": float(mx[3]), "k4_nnls_vs_closed_form_max_rel": float(mx[4]),
                           **result},
            "gpu_launches": int(_cabi.launch_count() - launches0),
            "measured_hbm_gbs": hbm,
        }
        assert out["properties"]["k5_deterministic"], "K5 CSR is not deterministic"
        assert out["properties"]["k5_column_sum_identity_max_rel"] < 1e-5, out["properties"]
        assert out["properties"]["k5_spot_cells_vs_fp64_max_rel"] < 2e-7, out["properties"]
        assert out["properties"]["k4_nnls_vs_closed_form_max_rel"] < 1e-5, out["properties"]
        if args.check:
            assert out["properties"]["sharded_equals_single_gpu"], "gene-sharded result differs from the single-GPU result"
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
