// Size / log normalisation on the device -- the `normalize` family of velocyto/analysis.py:535-676
// (_normalize_S :535-552, _normalize_U :554-584, _normalize_Sx :586-601, _normalize_Ux :603-633):
//   cell_size[c] = sum_g X[g, c];  norm_factor = avg_size / cell_size;  X_sz = norm_factor * X;
//   X_norm = log2(X_sz + pcount)
// In the cell-major layout a cell's total is a contiguous row sum and the rescaling is one streaming pass that
// writes both outputs (the reference makes three NumPy passes with fp64 temporaries).  HBM-bound: 4 B read and
// 4-8 B written per element.
#include "velo_common.cuh"

namespace velo {

// one warp per cell, float4 loads, fp64 accumulation (counts reach 1e5 per cell: fp32 sums would drift)
__global__ void __launch_bounds__(256) k_cell_sums(const float *__restrict__ X, int64_t ld, int64_t G, int64_t C,
                                                   double *__restrict__ sums)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const float *row = X + c * ld;
    double s = 0.0;
    const int64_t G4 = G >> 2;
    for (int64_t q = lane; q < G4; q += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + q);
        s += (static_cast<double>(v.x) + static_cast<double>(v.y)) + (static_cast<double>(v.z) + static_cast<double>(v.w));
    }
    for (int64_t g = (G4 << 2) + lane; g < G; g += 32) s += static_cast<double>(row[g]);
    s = warp_sum(s);
    if (lane == 0) sums[c] = s;
}

// out_sz[c, g] = factor[c] * X[c, g] (non-finite -> 0 when guard != 0, analysis.py:581,630);
// out_norm[c, g] = log2(out_sz + pcount).  Either output may be NULL; factor == NULL means 1 (size=False).
__global__ void __launch_bounds__(256) k_size_normalize(const float *__restrict__ X, int64_t ld, int64_t G, int64_t C,
                                                        const double *__restrict__ factor, double pcount, int guard,
                                                        float *__restrict__ out_sz, float *__restrict__ out_norm)
{
    const int64_t c = blockIdx.x;
    const double f = factor ? factor[c] : 1.0;
    const int64_t L4 = ld >> 2;                                // whole row incl. the pad columns (written as zeros)
    for (int64_t q = threadIdx.x; q < L4; q += blockDim.x) {
        const int64_t off = c * ld + (q << 2);
        const float4 v = *reinterpret_cast<const float4 *>(X + off);
        const float in[4] = {v.x, v.y, v.z, v.w};
        float sz[4], nm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool pad = (q << 2) + k >= G;                // pad columns stay zero
            double y = f * static_cast<double>(in[k]);         // the fp64 product the reference forms, rounded once
            if (guard && !isfinite(y)) y = 0.0;
            sz[k] = pad ? 0.f : static_cast<float>(y);
            // log2 in fp32 (log2f: ~1 ulp, FP32 pipe) of the fp64 sum rounded once: the fp64 log2 of round 1 was a
            // ~60-instruction fp64 sequence per element and held the kernel at 38 % of the DRAM rate
            nm[k] = pad ? 0.f : log2f(static_cast<float>(y + pcount));
        }
        if (out_sz) *reinterpret_cast<float4 *>(out_sz + off) = make_float4(sz[0], sz[1], sz[2], sz[3]);
        if (out_norm) *reinterpret_cast<float4 *>(out_norm + off) = make_float4(nm[0], nm[1], nm[2], nm[3]);
    }
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_cell_sums(const float *X_cm, int64_t ld, int64_t G, int64_t C, double *sums,
                                  velo_stream_t stream)
{
    VELO_REQUIRE(X_cm && sums && G > 0 && C >= 0 && ld >= G && (ld % 4) == 0, "cell_sums: bad arguments");
    VELO_REQUIRE((reinterpret_cast<uintptr_t>(X_cm) & 15) == 0, "cell_sums: X_cm must be 16-byte aligned");
    if (C == 0) return VELO_OK;
    const int wpb = 8;
    k_cell_sums<<<static_cast<unsigned>((C + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(X_cm, ld, G, C, sums);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_size_normalize(const float *X_cm, int64_t ld, int64_t G, int64_t C, const double *factor,
                                       double pcount, int nonfinite_to_zero, float *out_sz, float *out_norm,
                                       velo_stream_t stream)
{
    VELO_REQUIRE(X_cm && (out_sz || out_norm) && G > 0 && C >= 0 && ld >= G && (ld % 4) == 0,
                 "size_normalize: bad arguments");
    VELO_REQUIRE(((reinterpret_cast<uintptr_t>(X_cm) | reinterpret_cast<uintptr_t>(out_sz) |
                   reinterpret_cast<uintptr_t>(out_norm)) & 15) == 0,
                 "size_normalize: matrices must be 16-byte aligned");
    VELO_REQUIRE(C < (1LL << 31), "size_normalize: too many cells for one launch");
    if (C == 0) return VELO_OK;
    k_size_normalize<<<static_cast<unsigned>(C), 256, 0, as_stream(stream)>>>(X_cm, ld, G, C, factor, pcount,
                                                                              nonfinite_to_zero, out_sz, out_norm);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}
