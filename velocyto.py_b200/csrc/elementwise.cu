// K6 -- the elementwise chain predict_U -> calculate_velocity -> calculate_shift ->
// extrapolate_cell_at_t -> velocity transform (velocyto/analysis.py:1343-1346, 1369, 1398-1406,
// 1428-1431, 1577/1597), fused into one pass: reads S and U once, writes whichever outputs are
// requested.  The reference makes >= 12 NumPy passes with fp64 temporaries.
// K5 -- kNN smoothing Sx[c, :] = sum_n w[c, n] * S[idx[c, n], :] (velocyto/neighbors.py:416-423 on the
// weights of :385-390): a CSR-by-rows gather over the cell-major matrix, fp64 accumulation.
#include "velo_common.cuh"

namespace velo {

struct ChainParams {
    const float *S, *U;        // cell-major, C x ld
    const float *gamma, *q;    // per gene (q may be null)
    const float *vel_thr;      // per gene threshold eps * max_c Upred, or null   (analysis.py:1377-1379)
    float *Upred, *vel, *dS, *St, *dtr;   // outputs, each may be null
    int64_t ld, G, C;
    float dt_shift, dt_extrap, psc;
    int assumption;            // 0 constant_velocity, 1 constant_unspliced
    int transform;             // VELO_LINEAR / VELO_SQRT / VELO_LOG10 for dtr
    int clip;
};

__device__ __forceinline__ float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// grid.x = cell, threads sweep the row in float4 steps: no index division, gamma/q come as (L1-resident) float4
__global__ void __launch_bounds__(256) k_velocity_chain(const ChainParams p)
{
    const int64_t c = blockIdx.x;
    const int64_t G4 = (p.G + 3) >> 2;
    const bool vec_ok = (p.G & 3) == 0;                   // per-gene vectors are exactly G long
    for (int64_t g4 = threadIdx.x; g4 < G4; g4 += blockDim.x) {
        const int64_t g = g4 << 2;
        const int64_t off = c * p.ld + g;                 // ld % 4 == 0 -> 16-byte aligned
        const float4 s4 = *reinterpret_cast<const float4 *>(p.S + off);
        const float4 u4 = *reinterpret_cast<const float4 *>(p.U + off);
        const float s[4] = {s4.x, s4.y, s4.z, s4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w};
        float gam4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f}, thr4[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec_ok || g + 3 < p.G) {
            if (vec_ok) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(p.gamma + g));
                gam4[0] = t.x; gam4[1] = t.y; gam4[2] = t.z; gam4[3] = t.w;
                if (p.q) {
                    const float4 r = __ldg(reinterpret_cast<const float4 *>(p.q + g));
                    q4[0] = r.x; q4[1] = r.y; q4[2] = r.z; q4[3] = r.w;
                }
                if (p.vel_thr) {
                    const float4 r = __ldg(reinterpret_cast<const float4 *>(p.vel_thr + g));
                    thr4[0] = r.x; thr4[1] = r.y; thr4[2] = r.z; thr4[3] = r.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    gam4[k] = __ldg(p.gamma + g + k);
                    if (p.q) q4[k] = __ldg(p.q + g + k);
                    if (p.vel_thr) thr4[k] = __ldg(p.vel_thr + g + k);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (g + k < p.G) {
                    gam4[k] = __ldg(p.gamma + g + k);
                    if (p.q) q4[k] = __ldg(p.q + g + k);
                    if (p.vel_thr) thr4[k] = __ldg(p.vel_thr + g + k);
                }
        }
        float up[4], v[4], ds[4], st[4], dt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t gg = g + k;
            const bool valid = gg < p.G;
            const float gam = gam4[k];
            const float qq = q4[k];
            up[k] = fmaf(gam, s[k], qq);                                  // analysis.py:1343-1346
            v[k] = u[k] - up[k];                                          // analysis.py:1369
            if (p.vel_thr && valid && fabsf(v[k]) < thr4[k]) v[k] = 0.f;
            if (p.assumption == 0) {
                ds[k] = p.dt_shift * v[k];                                // analysis.py:1399
            } else {                                                      // analysis.py:1403-1406
                const float uo = fmaxf(u[k] - qq, 0.f);
                const float egt = expf(-gam * p.dt_shift);
                ds[k] = s[k] * egt + (1.f - egt) * uo / gam - s[k];
            }
            const float step = p.dt_extrap * ds[k];
            st[k] = s[k] + step;                                          // analysis.py:1429
            if (p.clip) st[k] = fmaxf(st[k], 0.f);                        // analysis.py:1431
            // hi_dim_t - hi_dim with hi_dim_t = hi_dim + used_delta_t * delta_S (not clipped), analysis.py:1538
            if (p.transform == VELO_SQRT) dt[k] = sqrtf(fabsf(step) + p.psc) * signf(step);          // :1597
            else if (p.transform == VELO_LOG10) dt[k] = log10f(fabsf(step) + p.psc) * signf(step);   // :1577
            else dt[k] = step;                                                                        // :1594
            if (!valid) up[k] = v[k] = ds[k] = st[k] = dt[k] = 0.f;       // keep pad columns zero
        }
        if (p.Upred) *reinterpret_cast<float4 *>(p.Upred + off) = make_float4(up[0], up[1], up[2], up[3]);
        if (p.vel) *reinterpret_cast<float4 *>(p.vel + off) = make_float4(v[0], v[1], v[2], v[3]);
        if (p.dS) *reinterpret_cast<float4 *>(p.dS + off) = make_float4(ds[0], ds[1], ds[2], ds[3]);
        if (p.St) *reinterpret_cast<float4 *>(p.St + off) = make_float4(st[0], st[1], st[2], st[3]);
        if (p.dtr) *reinterpret_cast<float4 *>(p.dtr + off) = make_float4(dt[0], dt[1], dt[2], dt[3]);
    }
}

// per-gene max over cells of gamma*S + q (for the eps threshold of calculate_velocity, analysis.py:1377-1378).
// grid = (gene blocks, cell slices): lanes = adjacent genes (coalesced 128-byte lines), every CTA scans one slice of the
// cell axis with 4 independent running maxima, slices merge through an order-preserving integer atomicMax.
__device__ __forceinline__ unsigned int f2ord(float f)
{
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(128) k_gene_max_upred(const float *__restrict__ S, int64_t ld, const float *__restrict__ gamma,
                                                        const float *__restrict__ q, int64_t G, int64_t C,
                                                        unsigned int *__restrict__ mx_ord)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const float gam = gamma[g], qq = q ? q[g] : 0.f;
    const int64_t per = (C + gridDim.y - 1) / gridDim.y;
    const int64_t c0 = static_cast<int64_t>(blockIdx.y) * per, c1 = min(C, c0 + per);
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    int64_t c = c0;
    for (; c + 4 <= c1; c += 4) {
        m0 = fmaxf(m0, fmaf(gam, S[c * ld + g], qq));
        m1 = fmaxf(m1, fmaf(gam, S[(c + 1) * ld + g], qq));
        m2 = fmaxf(m2, fmaf(gam, S[(c + 2) * ld + g], qq));
        m3 = fmaxf(m3, fmaf(gam, S[(c + 3) * ld + g], qq));
    }
    for (; c < c1; ++c) m0 = fmaxf(m0, fmaf(gam, S[c * ld + g], qq));
    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    if (c1 > c0) atomicMax(mx_ord + g, f2ord(mx));      // NaN (gamma = NaN) orders above +inf: propagates like np.max
}
__global__ void k_gene_max_finalize(const unsigned int *__restrict__ mx_ord, int64_t G, float eps, float *__restrict__ thr)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g < G) thr[g] = ord2f(mx_ord[g]) * eps;
}

// one CTA row-segment: cell c, 1024 genes; thread = one float4 of genes, loops over the cell's neighbours.
// Accumulation: the four products of a group of 4 neighbours are summed pairwise in fp32 and the group sum is folded
// into an fp64 accumulator -- one fp32->fp64 conversion per 4 neighbours instead of five (round 1 converted every
// gathered value: the conversions issue on the 16-lane XU pipe, which ncu showed 87 % busy).  Error of a group sum
// <= 3 fp32 roundings of same-sign terms (~1e-7 relative, unbiased); the fp64 fold keeps it from growing with k.
__global__ void __launch_bounds__(256) k_knn_smooth(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                                    const float *__restrict__ weights, const float *__restrict__ S,
                                                    float *__restrict__ out, int64_t ld, int64_t G, int maximum)
{
    const int64_t c = blockIdx.x;
    const int64_t g = (static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x) << 2;
    if (g >= G) return;                                   // G padded to ld (multiple of 4): whole float4 is in-row
    const int64_t p0 = indptr[c], p1 = indptr[c + 1];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t p = p0;
    for (; p + 8 <= p1; p += 8) {                         // 8 independent row loads in flight, two groups of 4
        float4 v[8];
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            w[k] = weights[p + k];
            v[k] = __ldg(reinterpret_cast<const float4 *>(S + static_cast<int64_t>(indices[p + k]) * ld + g));
        }
#pragma unroll
        for (int h = 0; h < 8; h += 4) {
            a0 += static_cast<double>(fmaf(w[h], v[h].x, w[h + 1] * v[h + 1].x) + fmaf(w[h + 2], v[h + 2].x, w[h + 3] * v[h + 3].x));
            a1 += static_cast<double>(fmaf(w[h], v[h].y, w[h + 1] * v[h + 1].y) + fmaf(w[h + 2], v[h + 2].y, w[h + 3] * v[h + 3].y));
            a2 += static_cast<double>(fmaf(w[h], v[h].z, w[h + 1] * v[h + 1].z) + fmaf(w[h + 2], v[h + 2].z, w[h + 3] * v[h + 3].z));
            a3 += static_cast<double>(fmaf(w[h], v[h].w, w[h + 1] * v[h + 1].w) + fmaf(w[h + 2], v[h + 2].w, w[h + 3] * v[h + 3].w));
        }
    }
    for (; p < p1; ++p) {
        const float w = weights[p];
        const float4 v = __ldg(reinterpret_cast<const float4 *>(S + static_cast<int64_t>(indices[p]) * ld + g));
        a0 += (double)(w * v.x); a1 += (double)(w * v.y); a2 += (double)(w * v.z); a3 += (double)(w * v.w);
    }
    float4 r = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
    if (maximum) {                                        // np.maximum(S_sz, Sx), analysis.py:1017-1019
        const float4 s = *reinterpret_cast<const float4 *>(S + c * ld + g);
        r.x = fmaxf(r.x, s.x); r.y = fmaxf(r.y, s.y); r.z = fmaxf(r.z, s.z); r.w = fmaxf(r.w, s.w);
    }
    *reinterpret_cast<float4 *>(out + c * ld + g) = r;
}

// out = f(dt * dS) (mode 0..2 = VELO_LINEAR/SQRT/LOG10, analysis.py:1577/1594/1597); mode 3: out = clip?(S + dt * dS)
// (extrapolate_cell_at_t, analysis.py:1429-1431); modes 4/5: the "logratio" pair log2(S + psc) and
// log2(|S + dt*dS| + psc) - log2(S + psc) (analysis.py:1582-1583).  Pad columns: S = dS = 0 -> mode 4 writes
// log2(psc) there, which no kernel reads (all sums stop at G).
__global__ void __launch_bounds__(256) k_delta_ops(const float *__restrict__ S, const float *__restrict__ dS,
                                                   float *__restrict__ out, int64_t n4, float dt, float psc, int mode,
                                                   int clip)
{
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n4;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 d4 = reinterpret_cast<const float4 *>(dS)[t];
        float d[4] = {d4.x, d4.y, d4.z, d4.w}, r[4];
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mode >= 3) s4 = reinterpret_cast<const float4 *>(S)[t];
        const float s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float step = dt * d[k];
            if (mode == VELO_SQRT) r[k] = sqrtf(fabsf(step) + psc) * signf(step);
            else if (mode == VELO_LOG10) r[k] = log10f(fabsf(step) + psc) * signf(step);
            else if (mode == VELO_LINEAR) r[k] = step;
            else if (mode == 3) {
                r[k] = s[k] + step;
                if (clip) r[k] = fmaxf(r[k], 0.f);
            } else if (mode == 4) r[k] = log2f(s[k] + psc);                                   // analysis.py:1582
            else r[k] = log2f(fabsf(s[k] + step) + psc) - log2f(s[k] + psc);                  // analysis.py:1583
        }
        reinterpret_cast<float4 *>(out)[t] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

// corr[r, n]: self pair -> 0, NaN -> 1 (optional); counts the NaNs it replaced (analysis.py:1604-1612)
__global__ void k_patch_corr(float *__restrict__ corr, int64_t ld, const int32_t *__restrict__ ixs, int64_t ixs_ld,
                             int64_t c0, int64_t nc, int64_t m, int patch_nan, unsigned long long *__restrict__ nan_count)
{
    const int64_t total = nc * m;
    unsigned long long local = 0;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = t / m, n = t - r * m;
        const int64_t i = ixs ? static_cast<int64_t>(ixs[r * ixs_ld + n]) : n;
        float v = corr[r * ld + n];
        if (i == c0 + r) v = 0.f;
        else if (patch_nan && v != v) {
            v = 1.f;
            ++local;
        }
        corr[r * ld + n] = v;
    }
    if (local && nan_count) atomicAdd(nan_count, local);
}

// delta_embedding[r, :] = sum_n (P[r, n] - 1/m) * unit(emb[ixs[r, n]] - emb[c0 + r])   (analysis.py:1704-1712)
__global__ void __launch_bounds__(256) k_embedding_shift(const float *__restrict__ P, int64_t ld,
                                                         const int32_t *__restrict__ ixs, int64_t ixs_ld,
                                                         const double *__restrict__ emb, int dims, int64_t c0,
                                                         int64_t nc, int64_t m, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nc) return;
    const double ex = emb[(c0 + r) * dims], ey = emb[(c0 + r) * dims + 1];
    const double inv_m = 1.0 / static_cast<double>(m);
    double ax = 0.0, ay = 0.0;
    for (int64_t n = lane; n < m; n += 32) {
        const int64_t i = ixs[r * ixs_ld + n];
        const double dx = emb[i * dims] - ex, dy = emb[i * dims + 1] - ey;
        const double nrm = sqrt(dx * dx + dy * dy);
        if (nrm > 0.0) {                                   // the reference zeroes the 0/0 diagonal entries
            const double w = (static_cast<double>(P[r * ld + n]) - inv_m) / nrm;
            ax += w * dx;
            ay += w * dy;
        }
    }
    ax = warp_sum(ax);
    ay = warp_sum(ay);
    if (lane == 0) {
        out[r * 2] = ax;
        out[r * 2 + 1] = ay;
    }
}

// calculate_grid_arrows (analysis.py:1794-1803): one warp per grid point.  w = N(0, sigma).pdf(dist) over the point's
// n_neighbors nearest cells, mass = sum w, flow = sum w * delta[neigh] / max(1, mass).  fp64 like the reference.
__global__ void __launch_bounds__(256) k_grid_flow(const int32_t *__restrict__ neighs, const double *__restrict__ dists,
                                                   int64_t npts, int k, const double *__restrict__ delta, int dims,
                                                   double sigma, double *__restrict__ mass, double *__restrict__ flow)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= npts) return;
    const double norm = 1.0 / (sigma * 2.5066282746310002);       // 1 / (sigma * sqrt(2 pi)): scipy.stats.norm.pdf
    double ms = 0.0;
    for (int n = lane; n < k; n += 32) {
        const double z = dists[p * k + n] / sigma;
        ms += exp(-0.5 * z * z) * norm;
    }
    ms = warp_sum(ms);
    const double inv = 1.0 / fmax(1.0, ms);
    for (int d = 0; d < dims; ++d) {
        double a = 0.0;
        for (int n = lane; n < k; n += 32) {
            const double z = dists[p * k + n] / sigma;
            a += exp(-0.5 * z * z) * norm * delta[static_cast<int64_t>(neighs[p * k + n]) * dims + d];
        }
        a = warp_sum(a);
        if (lane == 0) flow[p * dims + d] = a * inv;
    }
    if (lane == 0) mass[p] = ms;
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_grid_flow(const int32_t *neighs, const double *dists, int64_t npts, int k, const double *delta,
                                  int dims, double sigma, double *mass, double *flow, velo_stream_t stream)
{
    VELO_REQUIRE(neighs && dists && delta && mass && flow && npts >= 0 && k > 0 && dims > 0 && sigma > 0, "grid_flow: bad arguments");
    if (npts == 0) return VELO_OK;
    k_grid_flow<<<static_cast<unsigned>((npts + 7) / 8), 256, 0, as_stream(stream)>>>(neighs, dists, npts, k, delta, dims, sigma,
                                                                                    mass, flow);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_delta_transform(const float *delta_S_cm, float *out_cm, int64_t ld, int64_t C, double dt,
                                        int transform, double psc, velo_stream_t stream)
{
    VELO_REQUIRE(delta_S_cm && out_cm && ld > 0 && ld % 4 == 0 && C > 0, "delta_transform: bad arguments");
    VELO_REQUIRE(transform >= VELO_LINEAR && transform <= VELO_LOG10, "delta_transform: unknown transform");
    const int64_t n4 = C * ld / 4, blocks = (n4 + 255) / 256;
    k_delta_ops<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, as_stream(stream)>>>(
        nullptr, delta_S_cm, out_cm, n4, static_cast<float>(dt), static_cast<float>(psc), transform, 0);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_extrapolate(const float *S_cm, const float *delta_S_cm, float *out_cm, int64_t ld, int64_t C,
                                    double dt, int clip, velo_stream_t stream)
{
    VELO_REQUIRE(S_cm && delta_S_cm && out_cm && ld > 0 && ld % 4 == 0 && C > 0, "extrapolate: bad arguments");
    const int64_t n4 = C * ld / 4, blocks = (n4 + 255) / 256;
    k_delta_ops<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, as_stream(stream)>>>(
        S_cm, delta_S_cm, out_cm, n4, static_cast<float>(dt), 0.f, 3, clip);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_patch_corr(float *corr, int64_t ld, const int32_t *ixs, int64_t ixs_ld, int64_t c0, int64_t nc,
                                   int64_t m, int patch_nan, unsigned long long *nan_count, velo_stream_t stream)
{
    VELO_REQUIRE(corr && nc >= 0 && m >= 0 && ld >= m, "patch_corr: bad arguments");
    if (nc == 0 || m == 0) return VELO_OK;
    const int64_t blocks = (nc * m + 255) / 256;
    k_patch_corr<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, as_stream(stream)>>>(
        corr, ld, ixs, ixs_ld, c0, nc, m, patch_nan, nan_count);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_embedding_shift(const float *P, int64_t ld, const int32_t *ixs, int64_t ixs_ld,
                                        const double *embedding, int dims, int64_t c0, int64_t nc, int64_t m,
                                        double *out, velo_stream_t stream)
{
    VELO_REQUIRE(P && ixs && embedding && out && dims >= 2 && nc >= 0 && m > 0, "embedding_shift: bad arguments");
    if (nc == 0) return VELO_OK;
    k_embedding_shift<<<static_cast<unsigned>((nc + 7) / 8), 256, 0, as_stream(stream)>>>(P, ld, ixs, ixs_ld, embedding,
                                                                                          dims, c0, nc, m, out);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_velocity_chain(const float *S_cm, const float *U_cm, int64_t ld, const float *gamma,
                                       const float *q, const float *vel_thr, int64_t G, int64_t C, int assumption,
                                       double dt_shift, double dt_extrap, int clip, int transform, double psc,
                                       float *Upred, float *vel, float *delta_S, float *S_t, float *d_transformed,
                                       velo_stream_t stream)
{
    VELO_REQUIRE(S_cm && U_cm && gamma && G > 0 && C > 0 && ld >= G && ld % 4 == 0, "velocity_chain: bad arguments");
    VELO_REQUIRE(assumption == 0 || assumption == 1, "velocity_chain: unknown assumption %d", assumption);
    VELO_REQUIRE(assumption == 0 || q, "velocity_chain: constant_unspliced needs the offsets q");
    VELO_REQUIRE(transform >= VELO_LINEAR && transform <= VELO_LOG10, "velocity_chain: unknown transform");
    ChainParams p;
    p.S = S_cm; p.U = U_cm; p.gamma = gamma; p.q = q; p.vel_thr = vel_thr;
    p.Upred = Upred; p.vel = vel; p.dS = delta_S; p.St = S_t; p.dtr = d_transformed;
    p.ld = ld; p.G = G; p.C = C;
    p.dt_shift = static_cast<float>(dt_shift); p.dt_extrap = static_cast<float>(dt_extrap);
    p.psc = static_cast<float>(psc); p.assumption = assumption; p.transform = transform; p.clip = clip;
    VELO_REQUIRE(C <= 2147483647LL, "velocity_chain: too many cells");
    k_velocity_chain<<<static_cast<unsigned>(C), 256, 0, as_stream(stream)>>>(p);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_velocity_threshold(const float *S_cm, int64_t ld, const float *gamma, const float *q,
                                           int64_t G, int64_t C, double eps, float *thr, velo_stream_t stream)
{
    VELO_REQUIRE(S_cm && gamma && thr && G > 0 && C > 0 && ld >= G, "velocity_threshold: bad arguments");
    cudaStream_t st = as_stream(stream);
    unsigned int *mx = nullptr;
    VELO_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&mx), static_cast<size_t>(G) * 4, st));
    cudaError_t e = cudaMemsetAsync(mx, 0, static_cast<size_t>(G) * 4, st);         // 0 orders below every float
    if (e == cudaSuccess) {
        const int64_t gx = (G + 127) / 128;
        int64_t gy = (148 * 16 + gx - 1) / gx;                                      // ~16 CTAs per SM in total
        if (gy > (C + 63) / 64) gy = (C + 63) / 64;
        if (gy < 1) gy = 1;
        dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
        k_gene_max_upred<<<grid, 128, 0, st>>>(S_cm, ld, gamma, q, G, C, mx);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        k_gene_max_finalize<<<static_cast<unsigned>(gx), 128, 0, st>>>(mx, G, static_cast<float>(eps), thr);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    cudaFreeAsync(mx, st);
    VELO_CUDA_TRY(e);
    return VELO_OK;
}

extern "C" int velo_dev_knn_smooth(const int64_t *indptr, const int32_t *indices, const float *weights,
                                   const float *S_cm, float *out_cm, int64_t ld, int64_t G, int64_t C, int maximum,
                                   velo_stream_t stream)
{
    VELO_REQUIRE(indptr && indices && weights && S_cm && out_cm, "knn_smooth: null pointer");
    VELO_REQUIRE(G > 0 && C > 0 && ld >= G && ld % 4 == 0 && S_cm != out_cm, "knn_smooth: bad arguments");
    VELO_REQUIRE(C <= 2147483647LL, "knn_smooth: too many cells");
    const int64_t gy = (((G + 3) / 4) + 255) / 256;
    VELO_REQUIRE(gy <= 65535, "knn_smooth: too many genes");
    dim3 grid(static_cast<unsigned>(C), static_cast<unsigned>(gy));
    k_knn_smooth<<<grid, 256, 0, as_stream(stream)>>>(indptr, indices, weights, S_cm, out_cm, ld, G, maximum);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_logratio(const float *S_cm, const float *delta_S_cm, float *out_cm, int64_t ld, int64_t C,
                                 double dt, double psc, int which, velo_stream_t stream)
{
    VELO_REQUIRE(S_cm && out_cm && ld > 0 && ld % 4 == 0 && C > 0 && (which == 0 || which == 1), "logratio: bad arguments");
    VELO_REQUIRE(which == 0 || delta_S_cm, "logratio: the delta needs delta_S");
    const int64_t n4 = C * ld / 4, blocks = (n4 + 255) / 256;
    k_delta_ops<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, as_stream(stream)>>>(
        S_cm, which == 0 ? S_cm : delta_S_cm, out_cm, n4, static_cast<float>(dt), static_cast<float>(psc), 4 + which, 0);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

// cos_proj[c] = sum_g a[c,g]*b[c,g] / sqrt(sum_g b[c,g]^2), clipped to [0,1] after / penalty  (analysis.py:1718-1719)
namespace velo {
__global__ void __launch_bounds__(256) k_row_cosine_scale(const float *__restrict__ a, const float *__restrict__ b,
                                                          int64_t ld, int64_t G, int64_t C, double penalty,
                                                          double *__restrict__ scale)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double dot = 0.0, nn = 0.0;
    for (int64_t g = lane; g < G; g += 32) {
        const double x = a[c * ld + g], y = b[c * ld + g];
        dot = fma(x, y, dot);
        nn = fma(y, y, nn);
    }
    dot = warp_sum(dot);
    nn = warp_sum(nn);
    if (lane == 0) {
        double v = dot / sqrt(nn) / penalty;              // 0/0 -> NaN propagates like np.clip(nan)
        if (v < 0.0) v = 0.0;
        if (v > 1.0) v = 1.0;
        scale[c] = v;
    }
}
}  // namespace velo

extern "C" int velo_dev_row_cosine_scale(const float *delta_S_cm, const float *estim_cm, int64_t ld, int64_t G, int64_t C,
                                         double penalty, double *scale, velo_stream_t stream)
{
    VELO_REQUIRE(delta_S_cm && estim_cm && scale && G > 0 && C > 0 && ld >= G && penalty > 0, "row_cosine_scale: bad arguments");
    velo::k_row_cosine_scale<<<static_cast<unsigned>((C + 7) / 8), 256, 0, as_stream(stream)>>>(delta_S_cm, estim_cm, ld, G, C,
                                                                                                penalty, scale);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}
