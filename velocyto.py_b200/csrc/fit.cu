// K4 -- batched per-gene gamma fits (replaces the Python loop + one SciPy solver call per gene of
// velocyto/estimation.py:173-366).
//
// Pass 1 (k_gene_moments): for every gene the sufficient statistics of all four fit modes in ONE
// read of S, U (and W): n, sum x, sum y, sum xx, sum xy, sum yy, nnz(x), nnz(y) and their weighted
// twins.  Data is cell-major fp32 (x[cell*ld + gene]); a warp's lanes are 32 adjacent genes, so every
// load instruction is one fully-coalesced 128-byte line; the cell axis is split across blockIdx.y
// and reduced deterministically (fixed-order partials, no atomics).  Accumulation is fp64.
// Pass 2 (k_fit_finalize): closed-form solutions per gene:
//   mode 0  fit_slope                 nnls through the origin        = max(0, Sxy/Sxx)        (:173-188)
//   mode 1  fit_slope_offset          leastsq from (0,0)             = OLS slope + intercept   (:244-264)
//   mode 2  fit_slope_weighted        bounded Brent on m in (lo,hi)  = clip(Swxy/Swxx)         (:191-209)
//   mode 3  fit_slope_weighted_offset L-BFGS-B, box constraints      = exact box-constrained 2-parameter WLS (:212-241)
// with the reference's degenerate rules (x == 0 -> NaN slope, y == 0 -> 0, estimation.py:176-179) and its
// unweighted R^2 with non-finite -> -1e16 (estimation.py:325-331, 357-363).
#include "velo_common.cuh"

namespace velo {

constexpr int kMom = 14;   // moments per gene
// index: 0 n, 1 Sx, 2 Sy, 3 Sxx, 4 Sxy, 5 Syy, 6 nnzx, 7 nnzy, 8 Sw, 9 Swx, 10 Swy, 11 Swxx, 12 Swxy, 13 Swyy

// thread = GP adjacent genes (one 128-/64-bit load per matrix and cell), block = 256 threads = 1024 (512) genes
// = 4 KB (2 KB) contiguous per cell row -- whole DRAM pages instead of 512-byte fragments; 4 cells per trip are
// loaded before the fp64 accumulation consumes them.  GP = 4 unweighted (8 moments), 2 weighted (14 moments).
template <int GP> struct VecOf;
template <> struct VecOf<4> { using type = float4; };
template <> struct VecOf<2> { using type = float2; };

template <bool WEIGHTED, int GP>
__global__ void __launch_bounds__(256) k_gene_moments(const float *__restrict__ X, const float *__restrict__ Y,
                                                      const float *__restrict__ W, int64_t ld, int64_t ldw,
                                                      const uint8_t *__restrict__ cell_mask, int64_t G, int64_t C,
                                                      int64_t cells_per_part, double *__restrict__ partials)
{
    using V = typename VecOf<GP>::type;
    constexpr int NM = WEIGHTED ? kMom : 8;
    const int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * GP;
    const int64_t part = blockIdx.y;
    if (g >= G) return;                                    // ld % 4 == 0: a started group is inside the row
    const int64_t c_begin = part * cells_per_part;
    const int64_t c_end = min(C, c_begin + cells_per_part);
    double m[NM][GP];
#pragma unroll
    for (int k = 0; k < NM; ++k)
#pragma unroll
        for (int j = 0; j < GP; ++j) m[k][j] = 0.0;
    constexpr int U = 4;
    for (int64_t cb = c_begin; cb < c_end; cb += U) {
        V xs[U], ys[U], ws[U];
        bool on[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t c = cb + u;
            on[u] = c < c_end && (!cell_mask || cell_mask[c]);   // steady_state selection (analysis.py:1159-1162)
            if (on[u]) {
                xs[u] = __ldg(reinterpret_cast<const V *>(X + c * ld + g));
                ys[u] = __ldg(reinterpret_cast<const V *>(Y + c * ld + g));
                if (WEIGHTED) ws[u] = __ldg(reinterpret_cast<const V *>(W + c * ldw + g));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!on[u]) continue;
            const float *xv = reinterpret_cast<const float *>(&xs[u]);
            const float *yv = reinterpret_cast<const float *>(&ys[u]);
            const float *wv = reinterpret_cast<const float *>(&ws[u]);
#pragma unroll
            for (int j = 0; j < GP; ++j) {
                const double x = static_cast<double>(xv[j]), y = static_cast<double>(yv[j]);
                m[0][j] += 1.0;
                m[1][j] += x;
                m[2][j] += y;
                m[3][j] = fma(x, x, m[3][j]);
                m[4][j] = fma(x, y, m[4][j]);
                m[5][j] = fma(y, y, m[5][j]);
                m[6][j] += (x != 0.0);
                m[7][j] += (y != 0.0);
                if (WEIGHTED) {
                    const double w = static_cast<double>(wv[j]);
                    const double wx = w * x, wy = w * y;
                    m[NM - 6][j] += w;
                    m[NM - 5][j] += wx;
                    m[NM - 4][j] += wy;
                    m[NM - 3][j] = fma(wx, x, m[NM - 3][j]);
                    m[NM - 2][j] = fma(wx, y, m[NM - 2][j]);
                    m[NM - 1][j] = fma(wy, y, m[NM - 1][j]);
                }
            }
        }
    }
    double *out = partials + (part * kMom) * G + g;        // [part][moment][gene]
#pragma unroll
    for (int k = 0; k < kMom; ++k)
#pragma unroll
        for (int j = 0; j < GP; ++j)
            if (g + j < G) out[k * G + j] = k < NM ? m[k < NM ? k : 0][j] : 0.0;
}

struct FitOpts {
    int mode;          // 0..3, see header comment
    double lo, hi;     // slope bounds for the weighted modes
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// objective of mode 3 up to the constant Swyy: sum w (m x + q - y)^2
__device__ __forceinline__ double wls_obj(double m, double q, double sw, double swx, double swy, double swxx,
                                          double swxy, double swyy)
{
    return m * m * swxx + 2.0 * m * q * swx + q * q * sw - 2.0 * m * swxy - 2.0 * q * swy + swyy;
}

__global__ void __launch_bounds__(128) k_fit_finalize(const double *__restrict__ partials, int parts, int64_t G,
                                                      FitOpts opt, const double *__restrict__ hi_per_gene,
                                                      const double *__restrict__ q_fixed, float *__restrict__ gamma,
                                                      float *__restrict__ offset, float *__restrict__ r2,
                                                      double *__restrict__ moments_out)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double mom[kMom];
#pragma unroll
    for (int k = 0; k < kMom; ++k) mom[k] = 0.0;
    for (int p = 0; p < parts; ++p)                         // fixed order -> deterministic
#pragma unroll
        for (int k = 0; k < kMom; ++k) mom[k] += partials[(static_cast<int64_t>(p) * kMom + k) * G + g];
    if (moments_out)
#pragma unroll
        for (int k = 0; k < kMom; ++k) moments_out[k * G + g] = mom[k];
    const double n = mom[0], sx = mom[1], sy = mom[2], sxx = mom[3], sxy = mom[4], syy = mom[5];
    const double sw = mom[8], swx = mom[9], swy = mom[10], swxx = mom[11], swxy = mom[12], swyy = mom[13];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double m = 0.0, q = 0.0;
    if (hi_per_gene) opt.hi = hi_per_gene[g];            // limit_gamma: per-gene upper slope bound (estimation.py:199-204)
    if (mom[6] == 0.0) {                 // not np.any(x): "definitely not at steady state" -> NaN
        m = nan;
        q = 0.0;
    } else if (mom[7] == 0.0) {          // not np.any(y)
        m = 0.0;
        q = 0.0;
    } else if (q_fixed && (opt.mode == 1 || opt.mode == 3)) {
        // fixperc_q: offset pinned to a low percentile of y, slope by bounded 1-D least squares on (0, 20)
        // minimise sum w (x m - y + q)^2   (estimation.py:221-224, 254-257)
        q = q_fixed[g];
        m = opt.mode == 1 ? clampd((sxy - q * sx) / sxx, 0.0, 20.0) : clampd((swxy - q * swx) / swxx, 0.0, 20.0);
    } else if (opt.mode == 0) {
        m = sxy / sxx;
        if (m < 0.0) m = 0.0;            // nnls constraint
    } else if (opt.mode == 1) {
        const double den = n * sxx - sx * sx;
        m = (n * sxy - sx * sy) / den;
        q = (sy - m * sx) / n;
    } else if (opt.mode == 2) {
        m = clampd(swxy / swxx, opt.lo, opt.hi);
        if (!(swxx > 0.0)) m = 0.5 * (opt.lo + opt.hi);   // flat objective: Brent's bounded search stays mid-interval
    } else {
        // box: m in [lo, hi], q in [0, up_q], up_q = 2 sum(y w) / sum(w)        (estimation.py:235-240)
        const double qlo = 0.0, qhi = 2.0 * swy / sw;
        const double det = sw * swxx - swx * swx;
        double bm = nan, bq = nan, best = 1e300;
        bool inside = false;
        if (det > 0.0) {
            const double m0 = (sw * swxy - swx * swy) / det;
            const double q0 = (swy - m0 * swx) / sw;
            if (m0 >= opt.lo && m0 <= opt.hi && q0 >= qlo && q0 <= qhi) {
                bm = m0;
                bq = q0;
                inside = true;
            }
        }
        if (!inside) {
            // the minimum of a convex quadratic over a box lies on an edge: 1-D clipped solves
            for (int e = 0; e < 4; ++e) {
                double cm, cq;
                if (e < 2) {             // q fixed at a bound, optimise m
                    cq = e == 0 ? qlo : qhi;
                    cm = swxx > 0.0 ? clampd((swxy - cq * swx) / swxx, opt.lo, opt.hi) : opt.lo;
                } else {                 // m fixed at a bound, optimise q
                    cm = e == 2 ? opt.lo : opt.hi;
                    cq = sw > 0.0 ? clampd((swy - cm * swx) / sw, qlo, qhi) : qlo;
                }
                const double f = wls_obj(cm, cq, sw, swx, swy, swxx, swxy, swyy);
                if (f < best) {
                    best = f;
                    bm = cm;
                    bq = cq;
                }
            }
        }
        m = bm;
        q = bq;
    }
    gamma[g] = static_cast<float>(m);
    if (offset) offset[g] = static_cast<float>(q);
    if (r2) {
        // unweighted coefficient of determination of the fitted line (estimation.py:325-331, 357-363)
        // (evaluated with the fp64 m, q, as the reference does before its float32 store)
        const double ssres = m * m * sxx + 2.0 * m * q * sx + n * q * q - 2.0 * m * sxy - 2.0 * q * sy + syy;
        const double sstot = syy - sy * sy / n;
        const double v = 1.0 - ssres / sstot;
        r2[g] = isfinite(v) ? static_cast<float>(v) : -1e16f;
    }
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_fit_gammas(int mode, const float *S_cm, const float *U_cm, int64_t ld, const float *W_cm,
                                   int64_t ldw, const uint8_t *cell_mask, int64_t G, int64_t C, double lo, double hi,
                                   float *gamma, float *offset, float *r2, double *moments, velo_stream_t stream)
{
    return velo_dev_fit_gammas_ex(mode, S_cm, U_cm, ld, W_cm, ldw, cell_mask, G, C, lo, hi, nullptr, nullptr, gamma,
                                  offset, r2, moments, stream);
}

extern "C" int velo_dev_fit_gammas_ex(int mode, const float *S_cm, const float *U_cm, int64_t ld, const float *W_cm,
                                      int64_t ldw, const uint8_t *cell_mask, int64_t G, int64_t C, double lo, double hi,
                                      const double *hi_per_gene, const double *q_fixed, float *gamma, float *offset,
                                      float *r2, double *moments, velo_stream_t stream)
{
    VELO_REQUIRE(mode >= 0 && mode <= 3, "fit_gammas: unknown mode %d", mode);
    VELO_REQUIRE(S_cm && U_cm && gamma && G > 0 && C > 0 && ld >= G, "fit_gammas: bad arguments");
    const bool weighted = mode >= 2;
    VELO_REQUIRE(!weighted || (W_cm && ldw >= G), "fit_gammas: weighted mode needs W");
    VELO_REQUIRE(!weighted || hi > lo, "fit_gammas: empty slope interval");
    cudaStream_t st = as_stream(stream);
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    VELO_REQUIRE(ld % 4 == 0 && (!weighted || ldw % 4 == 0), "fit_gammas: ld must be a multiple of 4");
    const int threads = 128;                                   // finalize kernel: thread = gene
    const int64_t gblocks = (G + threads - 1) / threads;
    const int mthreads = 256;                                  // moments kernel: thread = 4 (2 if weighted) genes
    const int gp = weighted ? 2 : 4;
    const int64_t mblocks = ((G + gp - 1) / gp + mthreads - 1) / mthreads;
    // cell partitions: fill the resident CTA slots of the machine exactly ONCE (a 2.03-wave grid runs as 3 waves)
    int per_sm = 1;
    if (weighted)
        VELO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gene_moments<true, 2>, mthreads, 0));
    else
        VELO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gene_moments<false, 4>, mthreads, 0));
    const int64_t slots = static_cast<int64_t>(dp.sm_count) * (per_sm > 0 ? per_sm : 1);
    int64_t parts = slots / mblocks;                           // floor: never spill into a second wave
    if (parts > (C + 63) / 64) parts = (C + 63) / 64;          // at least ~64 cells per partition
    if (parts < 1) parts = 1;
    if (parts > 65535) parts = 65535;
    const int64_t cpp = (C + parts - 1) / parts;
    parts = (C + cpp - 1) / cpp;
    double *partials = nullptr;
    VELO_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&partials),
                                  static_cast<size_t>(parts) * kMom * G * sizeof(double), st));
    dim3 grid(static_cast<unsigned>(mblocks), static_cast<unsigned>(parts));
    if (weighted)
        k_gene_moments<true, 2><<<grid, mthreads, 0, st>>>(S_cm, U_cm, W_cm, ld, ldw, cell_mask, G, C, cpp, partials);
    else
        k_gene_moments<false, 4><<<grid, mthreads, 0, st>>>(S_cm, U_cm, nullptr, ld, 0, cell_mask, G, C, cpp, partials);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e1 = cudaGetLastError();
    FitOpts opt{mode, lo, hi};
    if (e1 == cudaSuccess) {
        k_fit_finalize<<<static_cast<unsigned>(gblocks), threads, 0, st>>>(partials, static_cast<int>(parts), G, opt,
                                                                           hi_per_gene, q_fixed, gamma, offset, r2,
                                                                           moments);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e1 = cudaGetLastError();
    }
    cudaFreeAsync(partials, st);
    VELO_CUDA_TRY(e1);
    return VELO_OK;
}
