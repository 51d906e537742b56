"""BASELINE config 3: 50k cells x 30k genes, estimate_transition_prob in full (all-pairs) mode on one B200.

    python scripts/bench_config3.py [--cells 50000] [--genes 30000] [--skip-sqrt]

linear  -> K2g (tcgen05 tensor cores, csrc/coldeltacor_tc.cu): roofline = 12 flop per pair-gene (two products, three
           fp16 MMAs each) against the measured dense bf16/fp16 GEMM peak of MEASURED_PEAKS.json;
sqrt    -> K2 (register-tiled fp32 + MUFU, csrc/coldeltacor_full.cu): one MUFU.SQRT per pair-gene, 16 lanes/SM/clk.
One JSON object per line.  Inputs are generated on the device; times are CUDA events around the whole call
(operand preparation included for K2g).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import ClockSampler  # noqa: E402
from velocyto_b200 import _cabi, device as dev  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=50_000)
    ap.add_argument("--genes", type=int, default=30_000)
    ap.add_argument("--skip-sqrt", action="store_true")
    ap.add_argument("--skip-fp32", action="store_true", help="skip the fp32 K2 kernels (11 s + 19 s at full size)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    C, G = a.cells, a.genes
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    gen = torch.Generator(device="cuda").manual_seed(3)
    e, d = dev.CellMajor.empty(C, G), dev.CellMajor.empty(C, G)
    for r0 in range(0, C, 4096):
        n = min(4096, C - r0)
        u1 = torch.rand((n, G), device="cuda", generator=gen).clamp_min_(1e-7)
        u2 = torch.rand((n, G), device="cuda", generator=gen).clamp_min_(1e-7)
        v = -(torch.log(u1) + torch.log(u2))
        v[torch.rand((n, G), device="cuda", generator=gen) < 0.3] = 0
        e.t[r0:r0 + n, :G] = v
        d.t[r0:r0 + n, :G] = torch.randn((n, G), device="cuda", generator=gen)
        del u1, u2, v
    stats = dev.cell_stats(d)
    out = torch.empty((C, C), dtype=torch.float32, device="cuda")
    pg = G * C * C

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        sampler = ClockSampler(0)
        sampler.start()
        l0 = _cabi.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps, sampler.stop(), (_cabi.launch_count() - l0) // reps

    nI, nJ = (C + 127) // 128, (C + 255) // 256
    tiles_all = nI * nJ
    tiles_upper = sum(1 for ti in range(nI) for tj in range(nJ) if tj >= ti // 2)
    tiles_lower = sum(1 for u in range((nI + 1) // 2) for tj in range(nJ) if tj < u)
    for sym in ("0", "1"):
        os.environ["VELO_TC_SYMMETRIC"] = sym
        ms, clocks, launches = timed(lambda: dev.coldeltacor_linear_tc(e, d, stats=stats, out=out), a.reps)
        tf = 12.0 * pg / (ms * 1e-3) / 1e12
        executed = 1.0 if sym == "0" else (tiles_upper + tiles_lower) / tiles_all       # MMA tile passes actually issued
        peak = peaks["bf16_tflops_sustained"] if ms > 200 else peaks["bf16_tflops"]
        print(json.dumps({
            "kernel": "K2g k_coldeltacor_tc2 (all-pairs linear, tcgen05 cta_group::2) + operand prep"
                      + (" -- symmetric-Q scheme (two launches)" if sym == "1" else " -- plain scheme (VELO_TC_SYMMETRIC=0)"),
            "workload": f"{C} cells x {G} genes, full C x C (BASELINE config 3)", "ms": ms, "cells_per_s": C / (ms * 1e-3),
            "pair_gene_per_s": pg / (ms * 1e-3), "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "flop_model": "USEFUL flops: 12 per pair-gene (P = B X^T and Q = X X^T, each hi*hi + hi*lo + lo*hi in fp16), "
                                       "whatever the scheme executes",
                         "executed_fraction_of_model": executed, "executed_tflops": tf * executed,
                         "executed_frac_of_peak": tf * executed / peak,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops%s (cuBLAS)" % ("_sustained" if ms > 200 else ""),
                         "burst_peak": peaks["bf16_tflops"], "sustained_peak": peaks["bf16_tflops_sustained"]},
            "clocks": clocks}))
    os.environ.pop("VELO_TC_SYMMETRIC", None)
    tp = torch.empty_like(out)
    ms_tp, _, _ = timed(lambda: dev.transition_prob(out, None, 0.05, out=tp), 2)
    print(json.dumps({"kernel": "k_transition_prob (full mode epilogue)", "ms": ms_tp,
                      "achieved_gbs": 2 * C * C * 4 / (ms_tp * 1e-3) / 1e9}))
    del tp
    if a.skip_fp32:
        return
    lib = _cabi.load()
    lib.velo_set_tensor_cores(0)
    try:
        ms, clocks, _ = timed(lambda: dev.coldeltacor(e, d, None, "linear", 0.0, stats=stats, out=out), 1)
        print(json.dumps({"kernel": "K2 k_coldeltacor_full<LINEAR> (fp32 FMA pipes; tensor cores switched off)", "ms": ms,
                          "cells_per_s": C / (ms * 1e-3), "pair_gene_per_s": pg / (ms * 1e-3), "clocks": clocks}))
    finally:
        lib.velo_set_tensor_cores(1)
    if not a.skip_sqrt:
        ms, clocks, _ = timed(lambda: dev.coldeltacor(e, d, None, "sqrt", 1.0, stats=stats, out=out), 1)
        sm = clocks.get("sm_mhz") or 1500.0
        mufu_peak = 148 * 16 * sm * 1e6                                  # MUFU lanes x SMs x clock under load
        print(json.dumps({"kernel": "K2 k_coldeltacor_full<SQRT> (estimate_transition_prob full mode default)", "ms": ms,
                          "cells_per_s": C / (ms * 1e-3), "pair_gene_per_s": pg / (ms * 1e-3),
                          "roofline": {"bound": "mufu", "achieved": pg / (ms * 1e-3), "peak": mufu_peak,
                                       "unit": "sqrt/s", "frac": pg / (ms * 1e-3) / mufu_peak,
                                       "peak_source": f"148 SMs x 16 MUFU lanes x {sm:.0f} MHz (median clock under load)"},
                          "clocks": clocks}))


if __name__ == "__main__":
    main()
