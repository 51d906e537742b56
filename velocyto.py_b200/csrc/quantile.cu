// Per-gene order statistics and the weight matrices of fit_gammas (velocyto/analysis.py:1179-1219).
//
// The reference builds its least-squares weights from np.percentile along the cell axis (three full
// partitions per gene for the default "maxmin_diag").  Here a CTA owns one gene row (gene-major fp32 copy
// of the matrix) and finds the needed order statistics by an 8-bit-per-pass radix select on the
// order-preserving integer image of the floats (4 histogram passes + one "next larger value" pass;
// the row stays L2-resident between passes), then applies NumPy's linear interpolation rule.
#include "velo_common.cuh"

namespace velo {

__device__ __forceinline__ uint32_t f2key(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct AllOf {
    __device__ __forceinline__ bool operator()(int64_t) const { return true; }
};
// subset {i : key[i] <= thr} (le) or {i : key[i] > thr}
struct KeyCmp {
    const float *key;
    double thr;
    bool le;
    __device__ __forceinline__ bool operator()(int64_t i) const
    {
        const double k = key[i];
        return le ? (k <= thr) : (k > thr);
    }
};

// block-wide: the k-th smallest (0-based) element of {row[i] : pred(i)} and the next one (== k-th if it is the last);
// n_sub = size of the subset
template <typename Pred>
__device__ void block_select_pair(const float *__restrict__ row, int64_t n, int64_t n_sub, int64_t k, Pred pred,
                                  float *v_lo, float *v_hi, unsigned int *hist /*256*/,
                                  unsigned long long *scratch /*4*/)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    uint32_t prefix = 0, mask = 0;
    int64_t rem = k;
    unsigned int eq_count = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += nt) hist[b] = 0;
        __syncthreads();
        // (plain shared-memory atomics: the hardware already combines same-address lanes -- a __match_any_sync
        // aggregation of the skewed bins was measured and is 25-40 % SLOWER, here and in the kNN select)
        for (int64_t i = tid; i < n; i += nt) {
            if (!pred(i)) continue;
            const uint32_t u = f2key(row[i]);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int64_t cum = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] > rem) break;
                cum += hist[b];
            }
            scratch[0] = static_cast<unsigned long long>(b);
            scratch[1] = static_cast<unsigned long long>(rem - cum);
            scratch[2] = hist[b];
        }
        __syncthreads();
        prefix |= static_cast<uint32_t>(scratch[0]) << shift;
        mask |= 255u << shift;
        rem = static_cast<int64_t>(scratch[1]);
        eq_count = static_cast<unsigned int>(scratch[2]);
        __syncthreads();
    }
    const float lo = key2f(prefix);
    float hi = lo;
    if (k + 1 < n_sub && rem + 1 >= static_cast<int64_t>(eq_count)) {
        // the next order statistic is the smallest key strictly above `prefix`
        uint32_t best = 0xffffffffu;
        for (int64_t i = tid; i < n; i += nt) {
            if (!pred(i)) continue;
            const uint32_t u = f2key(row[i]);
            if (u > prefix && u < best) best = u;
        }
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (tid == 0) scratch[3] = 0xffffffffull;
        __syncthreads();
        if ((tid & 31) == 0) atomicMin(&scratch[3], static_cast<unsigned long long>(best));
        __syncthreads();
        hi = key2f(static_cast<uint32_t>(scratch[3]));
        __syncthreads();
    }
    *v_lo = lo;
    *v_hi = hi;
}

// np.percentile(subset of row, q) with the default "linear" method; every thread returns the value.
// An empty subset gives NaN (as numpy does).
template <typename Pred>
__device__ double block_percentile(const float *__restrict__ row, int64_t n, double q, Pred pred, bool masked,
                                   unsigned int *hist, unsigned long long *scratch)
{
    int64_t n_sub = n;
    if (masked) {                                        // count the subset
        unsigned int cnt = 0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) cnt += pred(i) ? 1u : 0u;
        if (threadIdx.x == 0) hist[0] = 0;
        __syncthreads();
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&hist[0], cnt);
        __syncthreads();
        n_sub = hist[0];
        __syncthreads();
    }
    if (n_sub == 0) return __longlong_as_double(0x7ff8000000000000LL);
    const double pos = q / 100.0 * static_cast<double>(n_sub - 1);
    int64_t k = static_cast<int64_t>(floor(pos));
    if (k < 0) k = 0;
    if (k > n_sub - 1) k = n_sub - 1;
    const double t = pos - static_cast<double>(k);
    float lo, hi;
    block_select_pair(row, n, n_sub, k, pred, &lo, &hi, hist, scratch);
    const double a = lo, b = hi;
    // numpy's _lerp: a + (b-a)*t, evaluated from the right end for t >= 0.5
    const double diff = b - a;
    double r = a + diff * t;
    if (t >= 0.5) r = b - diff * (1.0 - t);
    if (t == 0.0) r = a;
    return r;
}

// Stage a gene row in shared memory: every order statistic makes 5+ passes over its row, and np.percentile is asked
// for 2-3 of them per gene.  TMA bulk copies (cp.async.bulk, completion on an mbarrier) when the row is 16-byte aligned,
// a cooperative copy otherwise.  *bar must be initialised (count 1) and is used for exactly one phase.
__device__ __forceinline__ const float *stage_row(const float *__restrict__ row, float *s_row, int64_t C, uint64_t *bar)
{
    const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) && (C % 4 == 0);
    if (aligned) {
        if (threadIdx.x == 0) {
            const uint32_t total = static_cast<uint32_t>(C * 4);
            mbar_expect_tx(bar, total);
            for (uint32_t off = 0; off < total; off += 32768u) {
                const uint32_t n = total - off < 32768u ? total - off : 32768u;
                tma_load_1d(reinterpret_cast<unsigned char *>(s_row) + off, reinterpret_cast<const unsigned char *>(row) + off, n, bar);
            }
        }
        mbar_wait(bar, 0);
    } else {
        for (int64_t i = threadIdx.x; i < C; i += blockDim.x) s_row[i] = row[i];
        __syncthreads();
    }
    return s_row;
}

// out[g*nq + j] = np.percentile(rows[g, :], q[j])
__global__ void __launch_bounds__(1024) k_row_percentiles(const float *__restrict__ rows, int64_t G, int64_t C,
                                                         const double *__restrict__ q, int nq, double *__restrict__ out,
                                                         int in_smem)
{
    extern __shared__ __align__(128) unsigned char q_smem[];
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long scratch[4];
    __shared__ uint64_t bar;
    const int64_t g = blockIdx.x;
    const float *row = rows + g * C;
    if (in_smem) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        row = stage_row(row, reinterpret_cast<float *>(q_smem), C, &bar);
    }
    for (int j = 0; j < nq; ++j) {
        const double r = block_percentile(row, C, q[j], AllOf(), false, hist, scratch);
        if (threadIdx.x == 0) out[g * nq + j] = r;
        __syncthreads();
    }
}

// Per-gene constraints of the non-default fit options (x = spliced row, y = unspliced row, gene-major fp32):
//   q_fix[g]    = median( y[x <= percentile(x, 1)] )                                 estimation.py:221, 254
//   up_gamma[g] = median(y) > median(x) ? max(1.5, percentile(y[x > p90(x)], 10) / median(x[x > p90(x)])) : 1.5
//                                                                                    estimation.py:199-204, 229-234
__global__ void __launch_bounds__(256) k_fit_constraints(const float *__restrict__ xr, const float *__restrict__ yr,
                                                         int64_t G, int64_t C, double *__restrict__ q_fix,
                                                         double *__restrict__ up_gamma)
{
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long scratch[4];
    const int64_t g = blockIdx.x;
    const float *x = xr + g * C, *y = yr + g * C;
    if (q_fix) {
        const double p1 = block_percentile(x, C, 1.0, AllOf(), false, hist, scratch);
        __syncthreads();
        const double v = block_percentile(y, C, 50.0, KeyCmp{x, p1, true}, true, hist, scratch);
        if (threadIdx.x == 0) q_fix[g] = v;
        __syncthreads();
    }
    if (up_gamma) {
        const double my = block_percentile(y, C, 50.0, AllOf(), false, hist, scratch);
        __syncthreads();
        const double mx = block_percentile(x, C, 50.0, AllOf(), false, hist, scratch);
        __syncthreads();
        double up = 1.5;
        if (my > mx) {                                       // uniform across the block
            const double p90 = block_percentile(x, C, 90.0, AllOf(), false, hist, scratch);
            __syncthreads();
            const KeyCmp high{x, p90, false};
            const double yy = block_percentile(y, C, 10.0, high, true, hist, scratch);
            __syncthreads();
            const double xx = block_percentile(x, C, 50.0, high, true, hist, scratch);
            __syncthreads();
            up = fmax(1.5, yy / xx);                         // np.maximum propagates NaN; fmax does not: patch below
            if (yy != yy || xx != xx || (yy / xx) != (yy / xx)) up = __longlong_as_double(0x7ff8000000000000LL);
        }
        if (threadIdx.x == 0) up_gamma[g] = up;
    }
}

// den[g] = perc[g]; zero -> max(rowmax, 0.001)      (analysis.py:1197-1199)
__global__ void __launch_bounds__(256) k_fix_denominators(const float *__restrict__ rows, int64_t G, int64_t C,
                                                          const double *__restrict__ perc, float *__restrict__ inv_den)
{
    __shared__ float smax[8];
    const int64_t g = blockIdx.x;
    double den = perc[g];
    if (den == 0.0) {                                   // uniform per block
        float mx = -INFINITY;
        for (int64_t i = threadIdx.x; i < C; i += blockDim.x) mx = fmaxf(mx, rows[g * C + i]);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, smax[w]);
            den = fmax(static_cast<double>(mx), 0.001);
        }
    }
    if (threadIdx.x == 0) inv_den[g] = static_cast<float>(1.0 / den);
}

// X[c,g] = S[c,g]*a[g] (op 0: + , op 1: *) U[c,g]*b[g]
__global__ void __launch_bounds__(256) k_scaled_combine(const float *__restrict__ S, const float *__restrict__ U,
                                                        const float *__restrict__ a, const float *__restrict__ b,
                                                        float *__restrict__ X, int64_t ld, int64_t G, int64_t C, int op)
{
    const int64_t total = C * ld;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t g = t % ld;
        float v = 0.f;
        if (g < G) {
            const float s = S[t] * a[g], u = U[t] * b[g];
            v = op == 0 ? s + u : s * u;
        }
        X[t] = v;
    }
}

// W[c,g] (+)= (X[c,g] <= down[g]) | (X[c,g] >= up[g])
__global__ void __launch_bounds__(256) k_threshold_weights(const float *__restrict__ X, int64_t ld,
                                                           const double *__restrict__ thr /*G x 2*/,
                                                           float *__restrict__ W, int64_t ldw, int64_t G, int64_t C,
                                                           int accumulate)
{
    const int64_t total = C * G;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t c = t / G, g = t - c * G;
        const double x = X[c * ld + g];
        const float w = (x <= thr[2 * g] || x >= thr[2 * g + 1]) ? 1.f : 0.f;
        if (accumulate) W[c * ldw + g] += w;
        else W[c * ldw + g] = w;
    }
}

// "maxmin_weighted" (analysis.py:1186-1192): R = (clip(S, down, up) - down) / (up - down), W = (R^p + (1 - R)^p) / 2.
// (min / max of the clipped row are the two percentiles themselves; up == down divides 0 by 0 like the reference.)
__global__ void __launch_bounds__(256) k_smooth_weights(const float *__restrict__ X, int64_t ld,
                                                        const double *__restrict__ thr /*G x 2*/, double power,
                                                        float *__restrict__ W, int64_t ldw, int64_t G, int64_t C)
{
    const int64_t total = C * G;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t c = t / G, g = t - c * G;
        const double down = thr[2 * g], up = thr[2 * g + 1];
        const double x = fmin(fmax(static_cast<double>(X[c * ld + g]), down), up);
        const double r = (x - down) / (up - down);
        W[c * ldw + g] = static_cast<float>(0.5 * (pow(r, power) + pow(1.0 - r, power)));
    }
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_row_percentiles(const float *rows_gc, int64_t G, int64_t C, const double *q_dev, int nq,
                                        double *out, velo_stream_t stream)
{
    VELO_REQUIRE(rows_gc && q_dev && out && G > 0 && C > 0 && nq > 0, "row_percentiles: bad arguments");
    VELO_REQUIRE(G <= 2147483647LL, "row_percentiles: too many rows");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const size_t row_bytes = (static_cast<size_t>(C) * 4 + 15) / 16 * 16;
    // rows up to 72 KB (18k cells): >= 3 CTAs per SM stay resident; longer rows keep streaming from L2, where 8 CTAs
    // per SM hide the latency better than one CTA could from shared memory
    const int in_smem = row_bytes <= (72u << 10) && row_bytes + 4096 <= static_cast<size_t>(dp.smem_optin) ? 1 : 0;
    if (in_smem)
        VELO_CUDA_TRY(cudaFuncSetAttribute(k_row_percentiles, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(row_bytes)));
    // long rows (config 5: 500k cells = 2 MB per gene) are re-read ~15 times per gene from L2 / HBM: 1024 threads per
    // CTA keep four times as many loads in flight as the 256 that suffice for rows staged in shared memory
    const int threads = C > 65536 ? 1024 : 256;
    k_row_percentiles<<<static_cast<unsigned>(G), threads, in_smem ? row_bytes : 0, as_stream(stream)>>>(rows_gc, G, C, q_dev,
                                                                                                      nq, out, in_smem);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

namespace {
struct Tmp {
    void *p = nullptr;
    cudaStream_t st;
    explicit Tmp(cudaStream_t s) : st(s) {}
    ~Tmp()
    {
        if (p) cudaFreeAsync(p, st);
    }
    cudaError_t alloc(size_t n) { return cudaMallocAsync(&p, n ? n : 16, st); }
};
}  // namespace

// kind: 0 maxmin_diag, 1 maxmin, 2 maxmin_double, 3 sum, 4 prod, 5 maxmin_weighted   (analysis.py:1181-1219)
extern "C" int velo_dev_fit_weights(int kind, const float *S_cm, const float *U_cm, const float *Sx_cm,
                                    const float *Ux_cm, int64_t ld, int64_t G, int64_t C, double perc_lo,
                                    double perc_hi, float *W_cm, int64_t ldw, velo_stream_t stream)
{
    return velo_dev_fit_weights_ex(kind, S_cm, U_cm, Sx_cm, Ux_cm, ld, G, C, perc_lo, perc_hi, 15.0, W_cm, ldw, stream);
}

extern "C" int velo_dev_fit_weights_ex(int kind, const float *S_cm, const float *U_cm, const float *Sx_cm,
                                       const float *Ux_cm, int64_t ld, int64_t G, int64_t C, double perc_lo,
                                       double perc_hi, double power, float *W_cm, int64_t ldw, velo_stream_t stream)
{
    VELO_REQUIRE(kind >= 0 && kind <= 5, "fit_weights: unknown kind %d", kind);
    VELO_REQUIRE(S_cm && U_cm && W_cm && G > 0 && C > 0 && ld >= G && ldw >= G, "fit_weights: bad arguments");
    VELO_REQUIRE((kind != 0 && kind != 2) || (Sx_cm && Ux_cm), "fit_weights: maxmin_diag needs Sx and Ux");
    cudaStream_t st = as_stream(stream);
    DeviceProps dp;
    int rc0 = get_device_props(&dp);                       // also configures the workspace pool
    if (rc0) return rc0;
    Tmp rows(st), X(st), perc(st), thr(st), qd(st), invS(st), invU(st);
    VELO_CUDA_TRY(rows.alloc(static_cast<size_t>(G) * C * 4));
    VELO_CUDA_TRY(X.alloc(static_cast<size_t>(C) * ld * 4));
    VELO_CUDA_TRY(perc.alloc(static_cast<size_t>(G) * 8));
    VELO_CUDA_TRY(thr.alloc(static_cast<size_t>(G) * 2 * 8));
    VELO_CUDA_TRY(qd.alloc(4 * 8));
    VELO_CUDA_TRY(invS.alloc(static_cast<size_t>(G) * 4));
    VELO_CUDA_TRY(invU.alloc(static_cast<size_t>(G) * 4));
    float *rows_p = static_cast<float *>(rows.p), *X_p = static_cast<float *>(X.p);
    double *perc_p = static_cast<double *>(perc.p), *thr_p = static_cast<double *>(thr.p), *q_p = static_cast<double *>(qd.p);
    float *invS_p = static_cast<float *>(invS.p), *invU_p = static_cast<float *>(invU.p);
    const double qs[4] = {kind <= 2 ? 99.9 : 99.0, perc_lo, perc_hi, 0.0};
    VELO_CUDA_TRY(cudaMemcpyAsync(q_p, qs, sizeof(qs), cudaMemcpyHostToDevice, st));
    const int64_t nblk = (C * ld + 255) / 256;
    const unsigned eg = static_cast<unsigned>(nblk < 148 * 32 ? nblk : 148 * 32);
    int rc;
    auto denominators = [&](const float *M, float *inv) -> int {   // 1 / percentile(M, q0, axis=cells), zero-guarded
        if ((rc = velo_dev_unpack_genemajor(M, ld, G, C, rows_p, 4, stream))) return rc;
        if ((rc = velo_dev_row_percentiles(rows_p, G, C, q_p, 1, perc_p, stream))) return rc;
        k_fix_denominators<<<static_cast<unsigned>(G), 256, 0, st>>>(rows_p, G, C, perc_p, inv);
        VELO_LAUNCH_CHECK();
        return VELO_OK;
    };
    auto thresholds = [&](const float *M, int accumulate) -> int {   // W (+)= (M <= p_lo) | (M >= p_hi)
        if ((rc = velo_dev_unpack_genemajor(M, ld, G, C, rows_p, 4, stream))) return rc;
        if ((rc = velo_dev_row_percentiles(rows_p, G, C, q_p + 1, 2, thr_p, stream))) return rc;
        k_threshold_weights<<<eg, 256, 0, st>>>(M, ld, thr_p, W_cm, ldw, G, C, accumulate);
        VELO_LAUNCH_CHECK();
        return VELO_OK;
    };
    if (kind == 1) return thresholds(S_cm, 0);                                       // "maxmin"  :1193-1195
    if (kind == 5) {                                                                 // "maxmin_weighted" :1186-1192
        if ((rc = velo_dev_unpack_genemajor(S_cm, ld, G, C, rows_p, 4, stream))) return rc;
        if ((rc = velo_dev_row_percentiles(rows_p, G, C, q_p + 1, 2, thr_p, stream))) return rc;
        k_smooth_weights<<<eg, 256, 0, st>>>(S_cm, ld, thr_p, power, W_cm, ldw, G, C);
        VELO_LAUNCH_CHECK();
        return VELO_OK;
    }
    if (kind == 0 || kind == 2) {                                                    // "maxmin_diag" :1196-1207
        if ((rc = denominators(Sx_cm, invS_p))) return rc;
        if ((rc = denominators(Ux_cm, invU_p))) return rc;
        k_scaled_combine<<<eg, 256, 0, st>>>(Sx_cm, Ux_cm, invS_p, invU_p, X_p, ld, G, C, 0);
        VELO_LAUNCH_CHECK();
        if ((rc = thresholds(X_p, 0))) return rc;
        if (kind == 2) return thresholds(Sx_cm, 1);                                   // "maxmin_double" :1217-1218
        return VELO_OK;
    }
    // "sum" / "prod": S / p99(S) (+|*) U / p99(U)   :1182-1185.  (A zero 99th percentile divides by zero in the
    // reference; here the denominator falls back to max(row max, 0.001) as in the maxmin_diag branch.)
    VELO_REQUIRE(ldw == ld, "fit_weights: sum/prod need ldw == ld");
    if ((rc = denominators(S_cm, invS_p))) return rc;
    if ((rc = denominators(U_cm, invU_p))) return rc;
    k_scaled_combine<<<eg, 256, 0, st>>>(S_cm, U_cm, invS_p, invU_p, W_cm, ldw, G, C, kind == 3 ? 0 : 1);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_fit_constraints(const float *S_cm, const float *U_cm, int64_t ld, int64_t G, int64_t C,
                                        double *q_fix, double *up_gamma, velo_stream_t stream)
{
    VELO_REQUIRE(S_cm && U_cm && G > 0 && C > 0 && ld >= G && (q_fix || up_gamma), "fit_constraints: bad arguments");
    cudaStream_t st = as_stream(stream);
    DeviceProps dp;
    int rc0 = get_device_props(&dp);
    if (rc0) return rc0;
    Tmp xr(st), yr(st);
    VELO_CUDA_TRY(xr.alloc(static_cast<size_t>(G) * C * 4));
    VELO_CUDA_TRY(yr.alloc(static_cast<size_t>(G) * C * 4));
    int rc;
    if ((rc = velo_dev_unpack_genemajor(S_cm, ld, G, C, xr.p, 4, stream))) return rc;
    if ((rc = velo_dev_unpack_genemajor(U_cm, ld, G, C, yr.p, 4, stream))) return rc;
    k_fit_constraints<<<static_cast<unsigned>(G), 256, 0, st>>>(static_cast<const float *>(xr.p),
                                                               static_cast<const float *>(yr.p), G, C, q_fix, up_gamma);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}
