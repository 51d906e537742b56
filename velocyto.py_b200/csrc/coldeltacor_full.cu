// K2 -- register-tiled all-pairs ("full") correlation: x_colDeltaCor{,Sqrt,Log10} (speedboosted.pyx:13-257).
//
// The full variants touch every (cell, target) pair: G*C^2 transform evaluations (7.5e13 at 50k x 30k).
// Running them through the row-gather kernel K1 re-reads every target row once per cell from L2
// (C^2*G*4 bytes); here a 64 cells x 64 targets tile of pairs shares its operands through shared memory
// and each thread keeps a 4x4 block of pairs (48 fp32 accumulators) in registers, GEMM-style, so the
// kernel is bound by the transform arithmetic (MUFU sqrt/lg2 + FP32 issue), not by memory:
//   operands per gene step and thread: 3 x LDS.128 for 16 pairs.
// Not a tensor-core shape: f(e_i - e_c) does not factor into a product (only the linear variant does).
#include "velo_common.cuh"

namespace velo {

constexpr int kTC = 64, kTI = 64, kGK = 32, kFullThreads = 256;

struct FullParams {
    const float *e_cm;     // C x ld   (targets: all cells)
    const float *d_cm;     // nc x ld  (local cells)
    const float *stats;    // nc x 2
    float *out;            // nc x out_ld
    int64_t ld, out_ld, G, C, c0, nc;
    float psc;
};

// -A for the FULL zero rule, given u = e_c - e_i (note the order; u == x - x is +0).  The reference has
// A = t > 0 ? r : -r with t = -u (speedboosted.pyx:110-114, 195-199), r = sqrt(|t|+psc) or log10(|t|+psc) -- r may be
// NEGATIVE for log10 with psc < 1 -- hence -A = (u < 0) ? -r : r = r with its sign bit flipped by the sign bit of u:
// one LOP3, and u == +0 gives -A = r, i.e. A = -r, exactly the reference's t == 0 branch.
template <int TR>
__device__ __forceinline__ float full_transform_neg(float u, float psc)   // returns -A
{
    if (TR == VELO_LINEAR) return u;
    const float a = fabsf(u);
    const float r = TR == VELO_SQRT ? sqrt_approx(a + psc) : lg2_approx(a + psc) * 0.30102999566398120f;
    return __int_as_float(__float_as_int(r) ^ (__float_as_int(u) & 0x80000000));
}

template <int TR>
__global__ void __launch_bounds__(kFullThreads, 2) k_coldeltacor_full(const FullParams p)
{
    extern __shared__ __align__(16) float sm[];
    // [stage][tile][gene][64]
    auto tile = [&](int stage, int which) { return sm + ((stage * 3 + which) * kGK) * 64; };
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;                    // 16 x 16 threads, 4 x 4 pairs each
    const int64_t i_base = static_cast<int64_t>(blockIdx.x) * kTI;   // targets
    const int64_t r_base = static_cast<int64_t>(blockIdx.y) * kTC;   // local cells
    // loader mapping: element e of 512 float4 per tile: row = e % 64, gene quad = e / 64
    const int lrow = tid & 63, lq0 = tid >> 6;                 // this thread loads quads lq0 and lq0 + 4
    const int64_t cell_r = r_base + lrow;                      // local cell row (for e_c, d)
    const int64_t targ_i = i_base + lrow;
    const bool cell_ok = cell_r < p.nc, targ_ok = targ_i < p.C;
    const float *ec_ptr = p.e_cm + (p.c0 + (cell_ok ? cell_r : 0)) * p.ld;
    const float *dc_ptr = p.d_cm + (cell_ok ? cell_r : 0) * p.ld;
    const float *ei_ptr = p.e_cm + (targ_ok ? targ_i : 0) * p.ld;
    const float mu = cell_ok ? p.stats[2 * cell_r] : 0.f;

    // sums over -A (s1), A^2 (s2) and A*b (s3) for 4 cells x 4 targets, held as fp32x2 pairs of adjacent targets:
    // the three updates per pair are packed FADD2/FFMA2 (two fp32 operations per issue slot), which leaves the
    // MUFU unit (one sqrt/lg2 per pair, 16 lanes per SM) as the only limiter instead of instruction issue
    float2 s1[4][2], s2[4][2], s3[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int h = 0; h < 2; ++h) s1[a][h] = s2[a][h] = s3[a][h] = make_float2(0.f, 0.f);

    float4 rc[2], rd[2], ri[2];
    auto gload = [&](int64_t k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t g = k0 + 4 * (lq0 + 4 * h);
            const bool in = g < p.G;                           // ld % 4 == 0: a started quad is inside the row
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            rc[h] = (in && cell_ok) ? __ldg(reinterpret_cast<const float4 *>(ec_ptr + g)) : z;
            rd[h] = (in && cell_ok) ? __ldg(reinterpret_cast<const float4 *>(dc_ptr + g)) : z;
            ri[h] = (in && targ_ok) ? __ldg(reinterpret_cast<const float4 *>(ei_ptr + g)) : z;
        }
    };
    auto sstore = [&](int stage) {
        float *tc = tile(stage, 0), *tb = tile(stage, 1), *ti = tile(stage, 2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = 4 * (lq0 + 4 * h);
            tc[(q + 0) * 64 + lrow] = rc[h].x; tc[(q + 1) * 64 + lrow] = rc[h].y;
            tc[(q + 2) * 64 + lrow] = rc[h].z; tc[(q + 3) * 64 + lrow] = rc[h].w;
            tb[(q + 0) * 64 + lrow] = rd[h].x - mu; tb[(q + 1) * 64 + lrow] = rd[h].y - mu;
            tb[(q + 2) * 64 + lrow] = rd[h].z - mu; tb[(q + 3) * 64 + lrow] = rd[h].w - mu;
            ti[(q + 0) * 64 + lrow] = ri[h].x; ti[(q + 1) * 64 + lrow] = ri[h].y;
            ti[(q + 2) * 64 + lrow] = ri[h].z; ti[(q + 3) * 64 + lrow] = ri[h].w;
        }
    };

    const int64_t nchunks = (p.G + kGK - 1) / kGK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int64_t ch = 0; ch < nchunks; ++ch) {
        const int stage = static_cast<int>(ch & 1);
        if (ch + 1 < nchunks) gload((ch + 1) * kGK);           // next chunk's loads fly during the math
        const float *tc = tile(stage, 0) + 4 * ty, *tb = tile(stage, 1) + 4 * ty, *ti = tile(stage, 2) + 4 * tx;
        const int gmax = static_cast<int>(min(static_cast<int64_t>(kGK), p.G - ch * kGK));
#pragma unroll 4
        for (int g = 0; g < gmax; ++g) {
            const float4 c4 = *reinterpret_cast<const float4 *>(tc + g * 64);
            const float4 b4 = *reinterpret_cast<const float4 *>(tb + g * 64);
            const float4 i4 = *reinterpret_cast<const float4 *>(ti + g * 64);
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
            const float2 iv[2] = {make_float2(i4.x, i4.y), make_float2(i4.z, i4.w)};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float2 ca = make_float2(cv[a], cv[a]);
                const float2 nb = make_float2(-bv[a], -bv[a]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 u = sub2(ca, iv[h]);                                      // e_c - e_i
                    const float2 nA = make_float2(full_transform_neg<TR>(u.x, p.psc),      // = -A
                                                  full_transform_neg<TR>(u.y, p.psc));
                    s1[a][h] = add2(s1[a][h], nA);
                    s2[a][h] = fma2(nA, nA, s2[a][h]);
                    s3[a][h] = fma2(nA, nb, s3[a][h]);                                     // (-A) * (-b)
                }
            }
        }
        if (ch + 1 < nchunks) {
            sstore(stage ^ 1);                                  // other buffer: last read two iterations ago
            __syncthreads();
        }
    }
    // epilogue
    const double invG = 1.0 / static_cast<double>(p.G);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t r = r_base + 4 * ty + a;
        if (r >= p.nc) continue;
        const double ssb = p.stats[2 * r + 1];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t i = i_base + 4 * tx + b;
            if (i >= p.C) continue;
            const float2 p1 = s1[a][b >> 1], p2 = s2[a][b >> 1], p3 = s3[a][b >> 1];
            const double S1 = -static_cast<double>((b & 1) ? p1.y : p1.x), S2 = (b & 1) ? p2.y : p2.x,
                         S3 = (b & 1) ? p3.y : p3.x;
            const double var = S2 - S1 * S1 * invG;
            const double corr = (var > 0.0 && ssb > 0.0) ? S3 / sqrt(var * ssb) : __longlong_as_double(0x7ff8000000000000LL);
            p.out[r * p.out_ld + i] = static_cast<float>(corr);
        }
    }
}

template <int TR>
static int launch_full(const FullParams &p, cudaStream_t st)
{
    const size_t smem = 2 * 3 * kGK * 64 * sizeof(float);      // 49152 B
    VELO_CUDA_TRY(cudaFuncSetAttribute(k_coldeltacor_full<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    dim3 grid(static_cast<unsigned>((p.C + kTI - 1) / kTI), static_cast<unsigned>((p.nc + kTC - 1) / kTC));
    k_coldeltacor_full<TR><<<grid, kFullThreads, smem, st>>>(p);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

// entry used by velo_dev_coldeltacor_ex for (ixs == NULL, rule == FULL, no residuals)
int coldeltacor_full_tiled(int transform, const float *e_cm, const float *d_cm, int64_t ld, const float *stats,
                           float *out, int64_t out_ld, int64_t G, int64_t C, int64_t c0, int64_t nc, double psc,
                           cudaStream_t st)
{
    FullParams p;
    p.e_cm = e_cm; p.d_cm = d_cm; p.stats = stats; p.out = out;
    p.ld = ld; p.out_ld = out_ld; p.G = G; p.C = C; p.c0 = c0; p.nc = nc;
    p.psc = static_cast<float>(psc);
    VELO_REQUIRE((nc + kTC - 1) / kTC <= 65535, "coldeltacor(full): too many local cells for one launch");
    switch (transform) {
    case VELO_LINEAR: return launch_full<VELO_LINEAR>(p, st);
    case VELO_SQRT: return launch_full<VELO_SQRT>(p, st);
    default: return launch_full<VELO_LOG10>(p, st);
    }
}

}  // namespace velo
