// K5s -- kNN smoothing with SPARSE counts: out[c, g] = sum_p w[p] * S[idx[p], g], S in CSR by cell.
//
// BASELINE config 5 (500k cells x 30k genes, ~5 % dense counts, k = 500): the dense input of K5 would be 60 GB per
// matrix, and the reference (dense float64 S, scipy coo_matmat_dense, neighbors.py:416-423) cannot hold it at all.
// Here the counts stay CSR (cells x genes, sorted gene ids) and only the OUTPUT gene slab [g0, g0+ng) of this rank is
// dense (gene-sharded across GPUs).
//
// One CTA per (cell, tile of <= 4096 genes): a shared-memory accumulator over the tile; each warp walks one
// neighbour's CSR row at a time (binary search for the tile's sub-range, then coalesced entries) and adds
// w * value into the accumulator.  Accumulation is 64-bit FIXED POINT (2^-32 resolution, native shared-memory
// integer atomics): exact integer addition makes the result independent of the order in which warps arrive --
// bit-reproducible, unlike fp32 atomics -- with 2.3e-10 absolute resolution on sums of normalised counts.
#include "velo_common.cuh"

namespace velo {

constexpr int kTileGenes = 4096;
constexpr double kFixScale = 4294967296.0;       // 2^32

__device__ __forceinline__ int64_t lower_bound_i32(const int32_t *__restrict__ a, int64_t lo, int64_t hi, int32_t key)
{
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_knn_smooth_csr(const int64_t *__restrict__ w_indptr,
                                                        const int32_t *__restrict__ w_indices,
                                                        const float *__restrict__ w_weights,
                                                        const int64_t *__restrict__ s_indptr,
                                                        const int32_t *__restrict__ s_genes,
                                                        const float *__restrict__ s_values, float *__restrict__ out,
                                                        int64_t ld_out, int64_t g0, int64_t ng, int maximum)
{
    // 64-bit fixed-point accumulators kept as two 32-bit halves: a 64-bit shared-memory atomicAdd compiles to a
    // compare-and-swap spin loop (ATOMS.CAST.SPIN.64), two 32-bit adds are native (ATOMS.ADD).  The low half wraps
    // mod 2^32; the add that causes a wrap sees it in the value it gets back and forwards the carry to the high half.
    // Integer additions commute, every carry is counted exactly once: the pair is the exact 64-bit sum whatever the
    // order in which warps arrive -- bit-reproducible, like the single 64-bit accumulator it replaces.
    __shared__ unsigned int acc_lo[kTileGenes];
    __shared__ unsigned int acc_hi[kTileGenes];
    const int64_t c = blockIdx.x;
    const int64_t t0 = static_cast<int64_t>(blockIdx.y) * kTileGenes;          // tile start within the slab
    const int tile = static_cast<int>(min(static_cast<int64_t>(kTileGenes), ng - t0));
    const int32_t gene_lo = static_cast<int32_t>(g0 + t0), gene_hi = gene_lo + tile;
    for (int i = threadIdx.x; i < tile; i += blockDim.x) {
        acc_lo[i] = 0u;
        acc_hi[i] = 0u;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t p0 = w_indptr[c], p1 = w_indptr[c + 1];
    // Every warp walks the CSR rows of its share of the neighbours.  The walk is a chain of dependent loads
    // (neighbour id -> row bounds -> entries), so a warp keeps TWO neighbours in flight, and the two binary searches
    // for the tile's sub-range are skipped when the row's first and last gene already lie inside the tile (always,
    // when a rank holds just its own gene slab -- the gene-sharded configuration): round 2 measured the searched
    // version at 0.33 of the HBM rate by SURVEY 8(d)'s byte count, latency-bound.
    auto bounds = [&](int64_t j, int64_t &a, int64_t &b) {
        const int64_t r0 = s_indptr[j], r1 = s_indptr[j + 1];
        a = r0;
        b = r1;
        if (r1 > r0) {
            const int32_t first = s_genes[r0], last = s_genes[r1 - 1];
            if (first < gene_lo) a = lower_bound_i32(s_genes, r0, r1, gene_lo);           // uniform across the warp
            if (last >= gene_hi) b = lower_bound_i32(s_genes, a, r1, gene_hi);
        }
    };
    auto scatter = [&](int64_t a, int64_t b, double w) {
        for (int64_t q = a + lane; q < b; q += 32) {
            const long long v = __double2ll_rn(w * static_cast<double>(s_values[q]));
            const unsigned int vlo = static_cast<unsigned int>(v), vhi = static_cast<unsigned int>(v >> 32);
            const int g = s_genes[q] - gene_lo;
            const unsigned int old = atomicAdd(&acc_lo[g], vlo);
            const unsigned int carry = (old + vlo) < old ? 1u : 0u;
            if (vhi + carry) atomicAdd(&acc_hi[g], vhi + carry);
        }
    };
    for (int64_t p = p0 + warp; p < p1; p += 2 * nwarps) {
        const int64_t pB = p + nwarps;
        const bool hasB = pB < p1;
        const int64_t jA = w_indices[p], jB = hasB ? w_indices[pB] : jA;
        const double wA = static_cast<double>(w_weights[p]) * kFixScale;
        const double wB = hasB ? static_cast<double>(w_weights[pB]) * kFixScale : 0.0;
        int64_t aA, bA, aB, bB;
        bounds(jA, aA, bA);
        bounds(jB, aB, bB);
        scatter(aA, bA, wA);
        if (hasB) scatter(aB, bB, wB);
    }
    __syncthreads();
    float *orow = out + c * ld_out + t0;
    for (int i = threadIdx.x; i < tile; i += blockDim.x)
        orow[i] = static_cast<float>(static_cast<double>(static_cast<long long>((static_cast<unsigned long long>(acc_hi[i]) << 32) |
                                                                                 acc_lo[i])) * (1.0 / kFixScale));
    if (maximum) {                                                               // np.maximum(S, Sx): own row, sparse
        __syncthreads();
        const int64_t r0 = s_indptr[c], r1 = s_indptr[c + 1];
        const int64_t a = lower_bound_i32(s_genes, r0, r1, gene_lo), b = lower_bound_i32(s_genes, a, r1, gene_hi);
        for (int64_t q = a + threadIdx.x; q < b; q += blockDim.x) {
            float *o = orow + (s_genes[q] - gene_lo);
            *o = fmaxf(*o, s_values[q]);
        }
    }
}

// CSR by cell -> dense cell-major slab [g0, g0 + ng): one warp per cell; the row is zero-filled with 128-bit stores,
// then the cell's entries inside the slab are scattered (sorted gene ids: binary search for the sub-range).
__global__ void __launch_bounds__(256) k_csr_to_cellmajor(const int64_t *__restrict__ indptr, const int32_t *__restrict__ genes,
                                                          const float *__restrict__ values, int64_t C, int64_t g0, int64_t ng,
                                                          float *__restrict__ out, int64_t ld)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    float4 *row4 = reinterpret_cast<float4 *>(out + c * ld);
    for (int64_t q = lane; q < (ld >> 2); q += 32) row4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const int64_t r0 = indptr[c], r1 = indptr[c + 1];
    const int64_t a = lower_bound_i32(genes, r0, r1, static_cast<int32_t>(g0));
    const int64_t b = lower_bound_i32(genes, a, r1, static_cast<int32_t>(g0 + ng));
    for (int64_t q = a + lane; q < b; q += 32) out[c * ld + (genes[q] - g0)] = values[q];
}

// per-cell totals of a CSR-by-cell matrix (= X.sum(0) of the reference's genes x cells matrix), fp64; optional
// in-place rescaling values[q] *= factor[cell] (size normalisation keeps the sparsity pattern)
__global__ void __launch_bounds__(256) k_csr_cell_sums_scale(const int64_t *__restrict__ indptr, float *__restrict__ values,
                                                             int64_t C, const double *__restrict__ factor,
                                                             double *__restrict__ sums)
{
    const int lane = threadIdx.x & 31;
    const int64_t c = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int64_t r0 = indptr[c], r1 = indptr[c + 1];
    if (factor) {
        const double f = factor[c];
        for (int64_t q = r0 + lane; q < r1; q += 32) {
            const double y = f * static_cast<double>(values[q]);
            values[q] = isfinite(y) ? static_cast<float>(y) : 0.f;
        }
    }
    if (sums) {
        double s = 0.0;
        for (int64_t q = r0 + lane; q < r1; q += 32) s += static_cast<double>(values[q]);
        s = warp_sum(s);
        if (lane == 0) sums[c] = s;
    }
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_csr_to_cellmajor(const int64_t *indptr, const int32_t *genes, const float *values, int64_t C,
                                         int64_t g0, int64_t ng, float *out_cm, int64_t ld, velo_stream_t stream)
{
    VELO_REQUIRE(indptr && genes && values && out_cm && C >= 0 && g0 >= 0 && ng > 0 && ld >= ng && (ld % 4) == 0,
                 "csr_to_cellmajor: bad arguments");
    VELO_REQUIRE((reinterpret_cast<uintptr_t>(out_cm) & 15) == 0, "csr_to_cellmajor: out_cm must be 16-byte aligned");
    if (C == 0) return VELO_OK;
    k_csr_to_cellmajor<<<static_cast<unsigned>((C + 7) / 8), 256, 0, as_stream(stream)>>>(indptr, genes, values, C, g0, ng,
                                                                                        out_cm, ld);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_csr_cell_sums_scale(const int64_t *indptr, float *values, int64_t C, const double *factor,
                                            double *sums, velo_stream_t stream)
{
    VELO_REQUIRE(indptr && values && C >= 0 && (factor || sums), "csr_cell_sums_scale: bad arguments");
    if (C == 0) return VELO_OK;
    k_csr_cell_sums_scale<<<static_cast<unsigned>((C + 7) / 8), 256, 0, as_stream(stream)>>>(indptr, values, C, factor, sums);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_knn_smooth_csr(const int64_t *w_indptr, const int32_t *w_indices, const float *w_weights,
                                       const int64_t *s_indptr, const int32_t *s_genes, const float *s_values,
                                       float *out_cm, int64_t ld_out, int64_t C, int64_t g0, int64_t ng, int maximum,
                                       velo_stream_t stream)
{
    VELO_REQUIRE(w_indptr && w_indices && w_weights && s_indptr && s_genes && s_values && out_cm, "knn_smooth_csr: null pointer");
    VELO_REQUIRE(C > 0 && C <= 2147483647LL && g0 >= 0 && ng > 0 && ld_out >= ng, "knn_smooth_csr: bad sizes");
    const int64_t tiles = (ng + kTileGenes - 1) / kTileGenes;
    VELO_REQUIRE(tiles <= 65535, "knn_smooth_csr: too many gene tiles");
    dim3 grid(static_cast<unsigned>(C), static_cast<unsigned>(tiles));
    k_knn_smooth_csr<<<grid, 256, 0, as_stream(stream)>>>(w_indptr, w_indices, w_weights, s_indptr, s_genes, s_values,
                                                          out_cm, ld_out, g0, ng, maximum);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}
