// K1/K2 -- the colDeltaCor correlation kernel (all six reference variants).
//
// Replaces x_colDeltaCor{,Sqrt,Log10}{,partial} (velocyto/speedboosted.pyx:13-538).
//
// Work decomposition (one persistent CTA per SM, static striding over cells):
//   for each local cell c:
//     for each gene slab [g0, g0+gl)  (slab = what fits in shared memory beside the accumulators)
//        TMA bulk copies (UBLKCP) stage e[c, slab] and d[c, slab] into shared memory, completion on an
//        mbarrier; d is centred in place (b = d - mean, speedboosted.pyx:46-55)
//        each WARP owns one neighbour row at a time: it streams e[i, slab] from HBM with 128-bit
//        no-L1-allocate loads (the row is used exactly once by this SM), forms A = f(e_i - e_c) and
//        accumulates the three one-pass sums  S1 = sum A, S2 = sum A^2, S3 = sum A*b  in 4-way
//        interleaved fp32 registers, then reduces them with warp shuffles.
//     finalise: corr = S3 / sqrt((S2 - S1^2/G) * ssb) in fp64, written compactly as out[c, n].
//
// The reference's seven fp64 passes over a (genes x m) scratch (speedboosted.pyx:275-346) collapse into
// one read of every neighbour row: algorithmic HBM traffic  C*(m+2)*G*4 bytes.
#include "velo_common.cuh"

namespace velo {

// 32 warps per SM (1 CTA / SM, shared-memory bound): the random-row gather is latency-bound, and bandwidth grows with the
// number of row streams in flight -- measured 6.3 / 6.8 / 7.0 / 7.5 / 7.5 TB/s at 512 / 640 / 768 / 896 / 1024 threads
// (profiles/r1_k1_ncu_v4.md).  64 registers per thread at 1024 threads: the single-buffer stream loop fits.
#ifndef VELO_K1_THREADS
#define VELO_K1_THREADS 1024
#endif
#ifndef VELO_K1_THREADS_EXACT
#define VELO_K1_THREADS_EXACT 768       // the tie-resolving variants need ~80 registers: 24 warps
#endif
template <bool EXACT>
struct K1Cfg {
    static constexpr int threads = EXACT ? VELO_K1_THREADS_EXACT : VELO_K1_THREADS;
    static constexpr int warps = threads / 32;
};
constexpr int kMaxChunk = 4096;          // neighbours handled per pass (accumulators + indices in smem)

struct CorrParams {
    const float *e_cm;      // C x ld
    const float *e_lo;      // optional C x ld residuals e64 - (double)e32 of fp64-origin data (nullptr = none)
    const float *d_cm;      // nc x ld
    const float *stats;     // nc x 2 : mean, centred sum of squares
    const int32_t *ixs;     // nc x ixs_ld, or nullptr (full)
    float *out;             // nc x out_ld
    int64_t ld, ixs_ld, out_ld;
    int64_t G, C, c0, nc, m;
    int Gs;                 // slab length (multiple of 128), H = ceil(G / Gs)
    int Mc;                 // neighbour chunk (<= kMaxChunk)
    int P;                  // power of two >= Mc: length of the sort network over the chunk's indices
    float psc;
};

template <int TR, int RULE>
__device__ __forceinline__ float transform_diff(float t, float psc)
{
    if (TR == VELO_LINEAR) {
        return t;
    } else if (TR == VELO_SQRT) {
        const float a = fabsf(t);
        if (RULE == VELO_RULE_PARTIAL) {
            // |t| < 1e-16 -> 0 ; t > 0 -> +r ; else -r          (speedboosted.pyx:372-378)
            // The zero rule is folded into the ARGUMENT of the square root: m = (|t| >= 1e-16) as 1.0 / 0.0 (one FSET.BF),
            // r = sqrt(|t| + psc * m) (FFMA + MUFU), sign transfer (LOP3) -- four instructions per element instead of
            // the five of "add, sqrt, compare, copysign, select".  t == 0 gives exactly 0.  A non-zero |t| below 1e-16
            // (fp32 values below ~1e-9: not reachable by normalised counts) yields sqrt(|t|) <= 1e-8 instead of 0.
            const float m = a < 1e-16f ? 0.0f : 1.0f;
            return copysignf(sqrt_approx(fmaf(psc, m, a)), t);
        }
        const float r = sqrt_approx(a + psc);
        {
            // t > 0 -> +r ; else (incl. t == 0) -r               (speedboosted.pyx:110-114)
            return t > 0.0f ? r : -r;
        }
    } else {
        const float a = fabsf(t);
        const float r = lg2_approx(a + psc) * 0.30102999566398120f;   // log10(x) = log2(x) * log10(2)
        if (RULE == VELO_RULE_PARTIAL) {
            return t >= 0.0f ? r : -r;                            // speedboosted.pyx:470-473
        } else {
            return t > 0.0f ? r : -r;                             // speedboosted.pyx:195-199
        }
    }
}

// interleaved partial sums kept as fp32x2 pairs: 2 pairs (4-way) for one row per warp, 1 pair (2-way) when a warp
// carries two rows at once (register budget)
#ifndef VELO_K1_TWOROWS
#define VELO_K1_TWOROWS 1
#endif
constexpr int kAccPairs = VELO_K1_TWOROWS ? 1 : 2;
struct Acc4 {
    float2 s1[kAccPairs], s2[kAccPairs], s3[kAccPairs];
};

// EXACT: the matrix came from fp64 data and carries fp32 residuals (e64 = e32 + lo).  Two DIFFERENT fp64 values
// can round to the SAME fp32 value; the reference then still sees a non-zero difference with a definite sign,
// and for psc > 0 the transforms jump by 2*f(0+) across zero.  Rounding is monotone, so the fp32 difference has
// the right sign whenever it is non-zero; only exact fp32 ties of non-zero values need the residuals, and those
// are rare (one global load pair per tie).  lo_i / lo_c point at the residuals of the same four genes.
struct Diff4 {                 // the four differences of one float4 of genes, plus the EXACT tie flag
    float2 t01, t23;
    bool tie;
};

template <bool EXACT>
__device__ __forceinline__ Diff4 diff4(const float4 v, const float4 ec)
{
    Diff4 d;
    d.t01 = sub2(make_float2(v.x, v.y), make_float2(ec.x, ec.y));
    d.t23 = sub2(make_float2(v.z, v.w), make_float2(ec.z, ec.w));
    d.tie = EXACT && (((v.x == ec.x) & (v.x != 0.0f)) | ((v.y == ec.y) & (v.y != 0.0f)) |
                      ((v.z == ec.z) & (v.z != 0.0f)) | ((v.w == ec.w) & (v.w != 0.0f)));
    return d;
}

template <int TR, int RULE, bool EXACT>
__device__ __forceinline__ void accumulate_diff(Acc4 &a, const Diff4 d, const float4 b, float psc, const float *lo_i,
                                                const float *lo_c)
{
    float t0 = d.t01.x, t1 = d.t01.y, t2 = d.t23.x, t3 = d.t23.y;
    if (EXACT) {
        if (d.tie) {
            const float4 li = __ldg(reinterpret_cast<const float4 *>(lo_i));
            const float4 lc = __ldg(reinterpret_cast<const float4 *>(lo_c));
            if (t0 == 0.0f) t0 = li.x - lc.x;
            if (t1 == 0.0f) t1 = li.y - lc.y;
            if (t2 == 0.0f) t2 = li.z - lc.z;
            if (t3 == 0.0f) t3 = li.w - lc.w;
        }
    }
    const float2 A01 = make_float2(transform_diff<TR, RULE>(t0, psc), transform_diff<TR, RULE>(t1, psc));
    const float2 A23 = make_float2(transform_diff<TR, RULE>(t2, psc), transform_diff<TR, RULE>(t3, psc));
    // packed accumulation: three FADD2/FFMA2 per element pair instead of six scalar operations
    constexpr int hi = kAccPairs - 1;
    a.s1[0] = add2(a.s1[0], A01); a.s2[0] = fma2(A01, A01, a.s2[0]); a.s3[0] = fma2(A01, make_float2(b.x, b.y), a.s3[0]);
    a.s1[hi] = add2(a.s1[hi], A23); a.s2[hi] = fma2(A23, A23, a.s2[hi]); a.s3[hi] = fma2(A23, make_float2(b.z, b.w), a.s3[hi]);
}

template <int TR, int RULE, bool EXACT>
__device__ __forceinline__ void accumulate4(Acc4 &a, const float4 v, const float4 ec, const float4 b, float psc,
                                            const float *lo_i, const float *lo_c)
{
    accumulate_diff<TR, RULE, EXACT>(a, diff4<EXACT>(v, ec), b, psc, lo_i, lo_c);
}
template <int TR, int RULE, bool EXACT>
__global__ void __launch_bounds__(K1Cfg<EXACT>::threads, 1) k_coldeltacor(const CorrParams p)
{
    constexpr int kThreads = K1Cfg<EXACT>::threads;
    constexpr int kWarps = K1Cfg<EXACT>::warps;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_e = reinterpret_cast<float *>(smem_raw);
    float *s_b = s_e + p.Gs;
    float *s_acc = s_b + p.Gs;                                   // 3 x Mc
    int32_t *s_ix = reinterpret_cast<int32_t *>(s_acc + 3 * p.Mc);   // P sorted neighbour ids (pad = INT_MAX)
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_ix + p.P);        // P % 2 == 0 -> 8-byte aligned
    uint16_t *s_pos = reinterpret_cast<uint16_t *>(bar + 1);         // P original positions within the chunk

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int Mc = p.Mc;
    const int H = static_cast<int>((p.G + p.Gs - 1) / p.Gs);
    const bool full = p.ixs == nullptr;
    const double invG = 1.0 / static_cast<double>(p.G);

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    for (int64_t r = blockIdx.x; r < p.nc; r += gridDim.x) {
        const int64_t cg = p.c0 + r;
        const float mu_b = p.stats[2 * r];
        const float ssb = p.stats[2 * r + 1];
        const float *e_row = p.e_cm + cg * p.ld;
        const float *d_row = p.d_cm + r * p.ld;

        for (int64_t n0 = 0; n0 < p.m; n0 += Mc) {
            const int mc = static_cast<int>(min(static_cast<int64_t>(Mc), p.m - n0));
            for (int t = tid; t < p.P; t += kThreads) {
                s_ix[t] = t < mc ? (full ? static_cast<int32_t>(n0 + t) : p.ixs[r * p.ixs_ld + n0 + t]) : 0x7fffffff;
                s_pos[t] = static_cast<uint16_t>(t);
                if (t < mc) {
                    s_acc[t] = 0.0f;
                    s_acc[Mc + t] = 0.0f;
                    s_acc[2 * Mc + t] = 0.0f;
                }
            }
            // Sort the neighbour list by cell id (bitonic network in shared memory, ~10 us per cell against
            // milliseconds of streaming).  The correlations do not depend on the order, results go back to the
            // original slots through s_pos -- but now every CTA sweeps the expression matrix in ascending row
            // order, so CTAs running side by side request the same rows at about the same time and all but the
            // first of them hit in the 126 MB L2 instead of HBM.
            if (!full) {
                for (int k = 2; k <= p.P; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        __syncthreads();
                        for (int t = tid; t < p.P; t += kThreads) {
                            const int l = t ^ j;
                            if (l > t) {
                                const int32_t a = s_ix[t], b = s_ix[l];
                                if ((a > b) == ((t & k) == 0)) {
                                    s_ix[t] = b;
                                    s_ix[l] = a;
                                    const uint16_t pa = s_pos[t];
                                    s_pos[t] = s_pos[l];
                                    s_pos[l] = pa;
                                }
                            }
                        }
                    }
            }

            for (int h = 0; h < H; ++h) {
                const int64_t g0 = static_cast<int64_t>(h) * p.Gs;
                const int gl = static_cast<int>(min(static_cast<int64_t>(p.Gs), p.G - g0));
                const int glp = (gl + 3) & ~3;            // ld % 4 == 0 keeps this inside the row
                fence_proxy_async_smem();                 // every thread: its generic smem accesses (centering writes,
                                                          // slab reads) are ordered before the async-proxy re-fill
                __syncthreads();                          // previous slab fully consumed; s_ix/s_acc init visible
                if (tid == 0) {
                    mbar_expect_tx(bar, 2u * glp * 4u);
                    tma_load_1d(s_e, e_row + g0, glp * 4u, bar);
                    tma_load_1d(s_b, d_row + g0, glp * 4u, bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1u;
                for (int t = tid; t < glp; t += kThreads) s_b[t] -= mu_b;
                __syncthreads();

                const int nq = gl >> 2;                   // whole float4 groups
                const float4 *s_e4 = reinterpret_cast<const float4 *>(s_e);
                const float4 *s_b4 = reinterpret_cast<const float4 *>(s_b);
                // Each warp streams whole neighbour rows, software-pipelined in ONE register buffer: the differences
                // e_i - e_c are formed first (the only consumers of the loaded values), then the next group's loads
                // are issued into the same registers and fly during the transform/accumulate math.
                // Measured alternatives (profiles/): a second buffer with 16 register copies per group (+2
                // instructions per element), and ping-pong buffers without copies, 10 % SLOWER -- two groups
                // outstanding under different consumers make the waits collapse onto the newest loads.
                // The stream is software-pipelined in registers:
                // while the 4 x 128-bit loads of group k are being consumed, those of group k+1 -- or of the
                // first group of the warp's NEXT row -- are already in flight (8 loads = 128 B per lane,
                // 64 KB per SM outstanding), which is what a random-row HBM gather needs to approach peak.
                auto load_group = [&](const float4 *row, int base, float4(&v)[4]) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int jj = base + lane + 32 * k;
                        if (jj < nq) v[k] = ldg_stream_f4(row + jj);
                    }
                };
                auto load_group_full = [&](const float4 *row, int base, float4(&v)[4]) {   // base + 128 <= nq
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = ldg_stream_f4(row + base + lane + 32 * k);
                };
                auto row_of = [&](int n) {
                    return reinterpret_cast<const float4 *>(p.e_cm + static_cast<int64_t>(s_ix[n]) * p.ld + g0);
                };
                const ptrdiff_t lo_delta =                                                 // floats, between the two matrices
                    EXACT ? (reinterpret_cast<intptr_t>(p.e_lo) - reinterpret_cast<intptr_t>(p.e_cm)) / 4 : 0;
                const float *lo_c_row = EXACT ? p.e_lo + cg * p.ld + g0 : nullptr;
#if VELO_K1_TWOROWS
                // Two neighbour rows per warp (pairs n, n + kWarps): every shared-memory read of e_c / b_c serves both
                // rows, halving the L1TEX load that v4 found at 87 % (8 B of shared memory per 4 B of HBM).  Each row
                // contributes two 128-bit loads per lane and trip, so registers and bytes in flight stay as before.
                auto finish_row = [&](int n, const Acc4 &a) {
                    const int64_t i = s_ix[n];
                    float s1 = a.s1[0].x + a.s1[0].y, s2 = a.s2[0].x + a.s2[0].y, s3 = a.s3[0].x + a.s3[0].y;
                    if (lane < (gl & 3)) {                               // ragged tail (G % 4 genes of the last slab)
                        const int k = (nq << 2) + lane;
                        const float v = __ldg(p.e_cm + i * p.ld + g0 + k);
                        float t = v - s_e[k];
                        if (EXACT && t == 0.0f && v != 0.0f)
                            t = __ldg(p.e_lo + i * p.ld + g0 + k) - __ldg(lo_c_row + k);
                        const float A = transform_diff<TR, RULE>(t, p.psc);
                        s1 += A;
                        s2 = fmaf(A, A, s2);
                        s3 = fmaf(A, s_b[k], s3);
                    }
                    s1 = warp_sum(s1);
                    s2 = warp_sum(s2);
                    s3 = warp_sum(s3);
                    if (lane == 0) {                                      // slabs are sequential -> deterministic order
                        s_acc[n] += s1;
                        s_acc[Mc + n] += s2;
                        s_acc[2 * Mc + n] += s3;
                    }
                };
                float4 cur[4];                                            // [0],[1]: row A at j, j+32; [2],[3]: row B
                auto load_pair = [&](const float4 *ra, const float4 *rb, int base) {
                    const int j0 = base + lane, j1 = j0 + 32;
                    if (base + 64 <= nq) {
                        cur[0] = ldg_stream_f4(ra + j0);
                        cur[1] = ldg_stream_f4(ra + j1);
                        cur[2] = ldg_stream_f4(rb + j0);
                        cur[3] = ldg_stream_f4(rb + j1);
                    } else {
                        if (j0 < nq) { cur[0] = ldg_stream_f4(ra + j0); cur[2] = ldg_stream_f4(rb + j0); }
                        if (j1 < nq) { cur[1] = ldg_stream_f4(ra + j1); cur[3] = ldg_stream_f4(rb + j1); }
                    }
                };
                int nA = warp;
                const float4 *rowA = nullptr, *rowB = nullptr;
                if (nA < mc) {
                    rowA = row_of(nA);
                    rowB = nA + kWarps < mc ? row_of(nA + kWarps) : rowA;   // odd tail: B mirrors A (L2 hit), result unused
                    load_pair(rowA, rowB, 0);
                }
                while (nA < mc) {
                    const int nB = nA + kWarps;
                    const bool hasB = nB < mc;
                    const int nA2 = nA + 2 * kWarps;
                    const float4 *rowA2 = nA2 < mc ? row_of(nA2) : nullptr;
                    const float4 *rowB2 = nA2 + kWarps < mc ? row_of(nA2 + kWarps) : rowA2;
                    Acc4 aA, aB;
                    aA.s1[0] = aA.s2[0] = aA.s3[0] = make_float2(0.0f, 0.0f);
                    aB.s1[0] = aB.s2[0] = aB.s3[0] = make_float2(0.0f, 0.0f);
                    // One trip = 64 float4 per row (two per lane and row).  The trips are split into a STEADY loop
                    // (every trip but the last complete one: the next group is a complete group of the same rows, so
                    // the loop carries no bounds logic at all -- round 2 counted ~2 control / address instructions
                    // per element in the single generic loop, a fifth of the issue slots), the last complete trip
                    // (whose prefetch is the ragged remainder or the first group of the warp's next rows) and the
                    // ragged remainder.
                    auto trip = [&](int base, bool prefetch_same_rows) {
                        const int j = base + lane;
                        const float *loA = EXACT ? reinterpret_cast<const float *>(rowA + j) + lo_delta : nullptr;
                        const float *loB = EXACT ? reinterpret_cast<const float *>(rowB + j) + lo_delta : nullptr;
                        const float *lo_c = EXACT ? lo_c_row + 4 * j : nullptr;
                        const float4 ec0 = s_e4[j], ec1 = s_e4[j + 32];
                        const Diff4 dA0 = diff4<EXACT>(cur[0], ec0), dA1 = diff4<EXACT>(cur[1], ec1);
                        const Diff4 dB0 = diff4<EXACT>(cur[2], ec0), dB1 = diff4<EXACT>(cur[3], ec1);
                        if (prefetch_same_rows) {                      // complete group: no predicates
                            cur[0] = ldg_stream_f4(rowA + j + 64);
                            cur[1] = ldg_stream_f4(rowA + j + 96);
                            cur[2] = ldg_stream_f4(rowB + j + 64);
                            cur[3] = ldg_stream_f4(rowB + j + 96);
                        } else if (base + 64 < nq) {
                            load_pair(rowA, rowB, base + 64);
                        } else if (rowA2) {
                            load_pair(rowA2, rowB2, 0);
                        }
                        const float4 b0 = s_b4[j], b1 = s_b4[j + 32];
                        accumulate_diff<TR, RULE, EXACT>(aA, dA0, b0, p.psc, loA, lo_c);
                        accumulate_diff<TR, RULE, EXACT>(aB, dB0, b0, p.psc, loB, lo_c);
                        accumulate_diff<TR, RULE, EXACT>(aA, dA1, b1, p.psc, loA + 128, lo_c + 128);
                        accumulate_diff<TR, RULE, EXACT>(aB, dB1, b1, p.psc, loB + 128, lo_c + 128);
                    };
                    const int full = nq >> 6;                           // complete trips (warp-uniform)
                    int base = 0;
                    for (int tr = 0; tr + 1 < full; ++tr, base += 64) trip(base, true);
                    if (full > 0) {
                        trip(base, false);
                        base += 64;
                    }
                    if (base < nq) {                                    // ragged last group of the row
                        const int j = base + lane;
                        const float *loA = EXACT ? reinterpret_cast<const float *>(rowA + j) + lo_delta : nullptr;
                        const float *loB = EXACT ? reinterpret_cast<const float *>(rowB + j) + lo_delta : nullptr;
                        const float *lo_c = EXACT ? lo_c_row + 4 * j : nullptr;
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            if (j + 32 * k < nq) {
                                const float4 ec = s_e4[j + 32 * k], bb = s_b4[j + 32 * k];
                                accumulate4<TR, RULE, EXACT>(aA, cur[k], ec, bb, p.psc, loA + 128 * k, lo_c + 128 * k);
                                accumulate4<TR, RULE, EXACT>(aB, cur[2 + k], ec, bb, p.psc, loB + 128 * k, lo_c + 128 * k);
                            }
                        if (rowA2) load_pair(rowA2, rowB2, 0);
                    }
                    finish_row(nA, aA);
                    if (hasB) finish_row(nB, aB);
                    nA = nA2;
                    rowA = rowA2;
                    rowB = rowB2;
                }
            }
#else
                int n = warp;
                const float4 *row = nullptr;
                float4 cur[4];
                if (n < mc) {
                    row = row_of(n);
                    load_group(row, 0, cur);
                }
                while (n < mc) {
                    const int n_next = n + kWarps;
                    const float4 *row_next = n_next < mc ? row_of(n_next) : nullptr;
                    Acc4 a;
#pragma unroll
                    for (int k = 0; k < 2; ++k) a.s1[k] = a.s2[k] = a.s3[k] = make_float2(0.0f, 0.0f);
                    for (int base = 0; base < nq; base += 128) {          // warp-uniform trip count
                        const int j = base + lane;
                        // residual rows (EXACT only): same offsets as the value rows, in the e_lo matrix
                        const float *lo_i = EXACT ? reinterpret_cast<const float *>(row + j) + lo_delta : nullptr;
                        const float *lo_c = EXACT ? lo_c_row + 4 * j : nullptr;
                        if (base + 128 <= nq) {                           // full group: no predicates
                            // (1) differences: the only consumers of the loaded values -> their registers are free
                            const Diff4 d0 = diff4<EXACT>(cur[0], s_e4[j]);
                            const Diff4 d1 = diff4<EXACT>(cur[1], s_e4[j + 32]);
                            const Diff4 d2 = diff4<EXACT>(cur[2], s_e4[j + 64]);
                            const Diff4 d3 = diff4<EXACT>(cur[3], s_e4[j + 96]);
                            // (2) next group (or the next row's first group) straight into the same registers:
                            //     no second buffer, no register copies, one group in flight during the math below
                            if (base + 256 <= nq) load_group_full(row, base + 128, cur);
                            else if (base + 128 < nq) load_group(row, base + 128, cur);
                            else if (row_next) load_group(row_next, 0, cur);
                            // (3) transform + accumulate
                            accumulate_diff<TR, RULE, EXACT>(a, d0, s_b4[j], p.psc, lo_i, lo_c);
                            accumulate_diff<TR, RULE, EXACT>(a, d1, s_b4[j + 32], p.psc, lo_i + 128, lo_c + 128);
                            accumulate_diff<TR, RULE, EXACT>(a, d2, s_b4[j + 64], p.psc, lo_i + 256, lo_c + 256);
                            accumulate_diff<TR, RULE, EXACT>(a, d3, s_b4[j + 96], p.psc, lo_i + 384, lo_c + 384);
                        } else {                                          // ragged last group of the row
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (j + 32 * k < nq)
                                    accumulate4<TR, RULE, EXACT>(a, cur[k], s_e4[j + 32 * k], s_b4[j + 32 * k], p.psc,
                                                                 lo_i + 128 * k, lo_c + 128 * k);
                            if (row_next) load_group(row_next, 0, cur);   // base + 128 >= nq here
                        }
                    }
                    const int64_t i = s_ix[n];
                    float s1 = (a.s1[0].x + a.s1[0].y) + (a.s1[1].x + a.s1[1].y);
                    float s2 = (a.s2[0].x + a.s2[0].y) + (a.s2[1].x + a.s2[1].y);
                    float s3 = (a.s3[0].x + a.s3[0].y) + (a.s3[1].x + a.s3[1].y);
                    // ragged tail (G % 4 genes of the last slab)
                    if (lane < (gl & 3)) {
                        const int k = (nq << 2) + lane;
                        const float v = __ldg(p.e_cm + i * p.ld + g0 + k);
                        float t = v - s_e[k];
                        if (EXACT && t == 0.0f && v != 0.0f)
                            t = __ldg(p.e_lo + i * p.ld + g0 + k) - __ldg(lo_c_row + k);
                        const float A = transform_diff<TR, RULE>(t, p.psc);
                        s1 += A;
                        s2 = fmaf(A, A, s2);
                        s3 = fmaf(A, s_b[k], s3);
                    }
                    s1 = warp_sum(s1);
                    s2 = warp_sum(s2);
                    s3 = warp_sum(s3);
                    if (lane == 0) {                      // slabs are sequential -> deterministic order
                        s_acc[n] += s1;
                        s_acc[Mc + n] += s2;
                        s_acc[2 * Mc + n] += s3;
                    }
                    n = n_next;
                    row = row_next;
                }
            }
#endif
            __syncthreads();
            // finalise this chunk of neighbours (fp64: a handful of operations per pair)
            for (int t = tid; t < mc; t += kThreads) {
                const double S1 = s_acc[t], S2 = s_acc[Mc + t], S3 = s_acc[2 * Mc + t];
                const double var = S2 - S1 * S1 * invG;
                const double den = var * static_cast<double>(ssb);
                // zero variance on either side -> NaN, as 0 * (1/sqrt(0)) in the reference
                const double corr = (var > 0.0 && ssb > 0.0f) ? S3 / sqrt(den) : __longlong_as_double(0x7ff8000000000000LL);
                p.out[r * p.out_ld + n0 + s_pos[t]] = static_cast<float>(corr);
            }
            __syncthreads();                              // s_acc is re-zeroed by the next chunk
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-cell statistics of the velocity rows (mean, centred sum of squares); one warp per cell
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_stats(const float *__restrict__ d_cm, int64_t ld, int64_t G, int64_t nc,
                                                    float *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nc) return;
    const float *row = d_cm + r * ld;
    double s = 0.0;
    for (int64_t g = lane; g < G; g += 32) s += static_cast<double>(row[g]);
    s = warp_sum(s);
    const double mu = s / static_cast<double>(G);
    const float muf = static_cast<float>(mu);
    double q = 0.0;
    for (int64_t g = lane; g < G; g += 32) {
        // centre with the fp32 mean: this is the value the correlation kernel subtracts
        const double b = static_cast<double>(row[g] - muf);
        q += b * b;
    }
    q = warp_sum(q);
    if (lane == 0) {
        stats[2 * r] = muf;
        stats[2 * r + 1] = static_cast<float>(q);
    }
}

// rm[(c0+r) * C + i] += out[r, n]
__global__ void k_scatter_dense(const float *__restrict__ out, int64_t out_ld, const int32_t *__restrict__ ixs,
                                int64_t ixs_ld, double *__restrict__ rm, int64_t C, int64_t c0, int64_t nc, int64_t m)
{
    const int64_t total = nc * m;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = t / m, n = t - r * m;
        const int64_t i = ixs ? static_cast<int64_t>(ixs[r * ixs_ld + n]) : n;
        atomicAdd(rm + (c0 + r) * C + i, static_cast<double>(out[r * out_ld + n]));
    }
}

// transition probabilities, one warp per cell row.  exp((v - vmax)/sigma): the row maximum is subtracted first (it
// cancels in the normalisation), so small sigma_corr cannot overflow fp32 -- the reference works in fp64 and stays
// finite down to sigma ~ 0.0014 (analysis.py:1697-1698).  patch_nan: NaN -> 1 as the knn_random branch does
// (analysis.py:1605-1606); without it a NaN poisons its whole row exactly like exp(NaN) / NaN-sum in the reference's
// "full" branch (analysis.py:1666-1668 only zeroes the diagonal).
__global__ void __launch_bounds__(256) k_transition_prob(const float *__restrict__ corr, int64_t ld,
                                                         const int32_t *__restrict__ ixs, int64_t ixs_ld,
                                                         float *__restrict__ p, int64_t p_ld, int64_t c0, int64_t nc,
                                                         int64_t m, double inv_sigma, int patch_nan)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nc) return;
    const float *cr = corr + r * ld;
    float *pr = p + r * p_ld;
    const int64_t self = c0 + r;
    float vmax = -INFINITY;
    bool bad = false;
    for (int64_t n = lane; n < m; n += 32) {
        float v = cr[n];
        const int64_t i = ixs ? static_cast<int64_t>(ixs[r * ixs_ld + n]) : n;
        if (i == self) v = 0.0f;            // np.fill_diagonal(corrcoef, 0)     analysis.py:1604
        else if (v != v) {
            if (patch_nan) v = 1.0f;        // NaN -> 1                           analysis.py:1605-1606
            else bad = true;
        }
        if (v == v) vmax = fmaxf(vmax, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    bad = __any_sync(0xffffffffu, bad);
    double sum = 0.0;
    for (int64_t n = lane; n < m; n += 32) {
        float v = cr[n];
        const int64_t i = ixs ? static_cast<int64_t>(ixs[r * ixs_ld + n]) : n;
        if (i == self) v = 0.0f;
        else if (v != v) v = 1.0f;
        // exponent and exp in fp64, like the reference: in fp32 the rounding of (v - vmax) / sigma alone is |x| * 6e-8
        // relative on exp(x) (2.4e-6 at x = -40, i.e. sigma = 0.05 and opposite correlations); 40 DFMA per element
        // are nothing next to the correlation kernel that produced it
        const double exd = exp(static_cast<double>(v - vmax) * inv_sigma);
        pr[n] = static_cast<float>(exd);
        sum += exd;
    }
    sum = warp_sum(sum);
    const double inv = 1.0 / sum;
    __syncwarp();
    for (int64_t n = lane; n < m; n += 32)
        pr[n] = bad ? __int_as_float(0x7fc00000) : static_cast<float>(static_cast<double>(pr[n]) * inv);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
template <int TR, int RULE, bool EXACT>
static int launch_corr_impl(const CorrParams &p, int grid, size_t smem, cudaStream_t st)
{
    VELO_CUDA_TRY(cudaFuncSetAttribute(k_coldeltacor<TR, RULE, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    k_coldeltacor<TR, RULE, EXACT><<<grid, K1Cfg<EXACT>::threads, smem, st>>>(p);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}
template <int TR, int RULE>
static int launch_corr(const CorrParams &p, int grid, size_t smem, cudaStream_t st)
{
    // the linear transform is continuous: ties need no special care
    if (TR != VELO_LINEAR && p.e_lo) return launch_corr_impl<TR, RULE, true>(p, grid, smem, st);
    return launch_corr_impl<TR, RULE, false>(p, grid, smem, st);
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_cell_stats(const float *d_cm, int64_t ld, int64_t G, int64_t nc, float *stats,
                                   velo_stream_t stream)
{
    VELO_REQUIRE(d_cm && stats && G > 0 && nc >= 0 && ld >= G, "cell_stats: bad arguments");
    if (nc == 0) return VELO_OK;
    const int wpb = 8;
    k_cell_stats<<<static_cast<unsigned>((nc + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(d_cm, ld, G, nc,
                                                                                                  stats);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_coldeltacor(int transform, int rule, const float *e_cm, const float *d_cm, int64_t ld,
                                    const float *stats, const int32_t *ixs, int64_t ixs_ld, float *out,
                                    int64_t out_ld, int64_t G, int64_t C, int64_t c0, int64_t nc, int64_t m,
                                    double psc, velo_stream_t stream)
{
    return velo_dev_coldeltacor_ex(transform, rule, e_cm, nullptr, d_cm, ld, stats, ixs, ixs_ld, out, out_ld, G, C, c0,
                                   nc, m, psc, stream);
}

extern "C" int velo_dev_coldeltacor_ex(int transform, int rule, const float *e_cm, const float *e_lo_cm,
                                       const float *d_cm, int64_t ld, const float *stats, const int32_t *ixs,
                                       int64_t ixs_ld, float *out, int64_t out_ld, int64_t G, int64_t C, int64_t c0,
                                       int64_t nc, int64_t m, double psc, velo_stream_t stream)
{
    VELO_REQUIRE(transform >= VELO_LINEAR && transform <= VELO_LOG10, "coldeltacor: unknown transform %d", transform);
    VELO_REQUIRE(rule == VELO_RULE_FULL || rule == VELO_RULE_PARTIAL, "coldeltacor: unknown rule %d", rule);
    VELO_REQUIRE(e_cm && d_cm && stats && out, "coldeltacor: null pointer");
    VELO_REQUIRE(G > 0 && C > 0 && nc >= 0 && c0 >= 0 && c0 + nc <= C && m >= 0, "coldeltacor: bad sizes");
    VELO_REQUIRE(ld >= G && (ld % 4) == 0, "coldeltacor: ld (%lld) must be >= G and a multiple of 4",
                 static_cast<long long>(ld));
    VELO_REQUIRE((reinterpret_cast<uintptr_t>(e_cm) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_cm) & 15) == 0,
                 "coldeltacor: e_cm/d_cm must be 16-byte aligned");
    VELO_REQUIRE(ixs != nullptr || m == C, "coldeltacor: full mode (ixs == NULL) requires m == C");
    VELO_REQUIRE(ixs == nullptr || ixs_ld >= m, "coldeltacor: ixs_ld < m");
    VELO_REQUIRE(out_ld >= m, "coldeltacor: out_ld < m");
    VELO_REQUIRE(C < (1LL << 31), "coldeltacor: C must fit int32 indices");
    if (nc == 0 || m == 0) return VELO_OK;
    // all-pairs linear: a pair of GEMMs over the gene axis -> tensor cores (K2g); the zero rules only concern the
    // sqrt/log10 transforms, and a linear difference has no fp32 tie to resolve
    if (ixs == nullptr && transform == VELO_LINEAR && velo_get_tensor_cores())
        return velo_dev_coldeltacor_tc(e_cm, d_cm, ld, stats, out, out_ld, G, C, c0, nc, nullptr, nullptr, stream);
    // all-pairs with the full zero rule and no tie residuals: the register-tiled kernel K2 (compute-bound)
    if (ixs == nullptr && rule == VELO_RULE_FULL && e_lo_cm == nullptr)
        return coldeltacor_full_tiled(transform, e_cm, d_cm, ld, stats, out, out_ld, G, C, c0, nc, psc,
                                      as_stream(stream));

    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc != VELO_OK) return rc;

    CorrParams p;
    VELO_REQUIRE(e_lo_cm == nullptr || (reinterpret_cast<uintptr_t>(e_lo_cm) & 15) == 0,
                 "coldeltacor: e_lo_cm must be 16-byte aligned");
    p.e_cm = e_cm; p.e_lo = e_lo_cm; p.d_cm = d_cm; p.stats = stats; p.ixs = ixs; p.out = out;
    p.ld = ld; p.ixs_ld = ixs_ld; p.out_ld = out_ld;
    p.G = G; p.C = C; p.c0 = c0; p.nc = nc; p.m = m;
    p.psc = static_cast<float>(psc);
    // shared-memory plan: [e slab | b slab | 3*Mc accumulators | P indices | mbarrier | P positions (u16)]
    int Mc = static_cast<int>(m < kMaxChunk ? round_up(m, 2) : kMaxChunk);
    int P = 2;
    while (P < Mc) P <<= 1;
    const int64_t side = 12LL * Mc + 6LL * P + 64;
    const int64_t budget = static_cast<int64_t>(dp.smem_optin) - side;
    VELO_REQUIRE(budget >= 8 * 128, "coldeltacor: not enough shared memory (%d bytes opt-in)", dp.smem_optin);
    const int64_t gs_max = (budget / 8) / 128 * 128;           // genes per slab that fit
    const int64_t H = (G + gs_max - 1) / gs_max;
    const int64_t Gs = round_up((G + H - 1) / H, 128);         // balanced slabs, 128-gene aligned
    p.Gs = static_cast<int>(Gs);
    p.Mc = Mc;
    p.P = P;
    const size_t smem = static_cast<size_t>(8 * Gs + side);
    const int grid = static_cast<int>(nc < dp.sm_count ? nc : dp.sm_count);
    cudaStream_t st = as_stream(stream);

#define VELO_DISPATCH(TR)                                                                        \
    (rule == VELO_RULE_PARTIAL ? launch_corr<TR, VELO_RULE_PARTIAL>(p, grid, smem, st)           \
                               : launch_corr<TR, VELO_RULE_FULL>(p, grid, smem, st))
    switch (transform) {
    case VELO_LINEAR: return VELO_DISPATCH(VELO_LINEAR);
    case VELO_SQRT: return VELO_DISPATCH(VELO_SQRT);
    default: return VELO_DISPATCH(VELO_LOG10);
    }
#undef VELO_DISPATCH
}

extern "C" int velo_dev_scatter_dense(const float *out, int64_t out_ld, const int32_t *ixs, int64_t ixs_ld,
                                      double *rm, int64_t C, int64_t c0, int64_t nc, int64_t m,
                                      velo_stream_t stream)
{
    VELO_REQUIRE(out && rm && C > 0 && nc >= 0 && m >= 0 && c0 >= 0 && c0 + nc <= C, "scatter_dense: bad arguments");
    VELO_REQUIRE(ixs != nullptr || m == C, "scatter_dense: full mode requires m == C");
    if (nc == 0 || m == 0) return VELO_OK;
    const int64_t total = nc * m;
    const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    k_scatter_dense<<<grid, 256, 0, as_stream(stream)>>>(out, out_ld, ixs, ixs_ld, rm, C, c0, nc, m);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_transition_prob_ex(const float *corr, int64_t ld, const int32_t *ixs, int64_t ixs_ld, float *p,
                                           int64_t p_ld, int64_t c0, int64_t nc, int64_t m, double sigma,
                                           int patch_nan, velo_stream_t stream)
{
    VELO_REQUIRE(corr && p && nc >= 0 && m > 0 && ld >= m && p_ld >= m && sigma > 0, "transition_prob: bad arguments");
    if (nc == 0) return VELO_OK;
    const int wpb = 8;
    k_transition_prob<<<static_cast<unsigned>((nc + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(
        corr, ld, ixs, ixs_ld, p, p_ld, c0, nc, m, 1.0 / sigma, patch_nan);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_transition_prob(const float *corr, int64_t ld, const int32_t *ixs, int64_t ixs_ld, float *p,
                                        int64_t p_ld, int64_t c0, int64_t nc, int64_t m, double sigma,
                                        velo_stream_t stream)
{
    return velo_dev_transition_prob_ex(corr, ld, ixs, ixs_ld, p, p_ld, c0, nc, m, sigma, 1, stream);
}
