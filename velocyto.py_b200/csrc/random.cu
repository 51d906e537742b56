// Device-side randomisation for estimate_transition_prob (opt-in, `random_backend="device"`):
//   * the per-cell weighted neighbour sampler of analysis.py:1552-1566 -- one legacy
//     `np.random.choice(n+1, size, replace=False, p=p)` per cell in a Python loop (0.3-0.5 ms per cell: tens of seconds
//     at 100k cells, an order of magnitude more than the correlation kernel it feeds);
//   * the randomised control `permute_rows_nsign` (analysis.py:2413-2420): every gene row shuffled and sign-flipped.
// NumPy's algorithm keeps the first occurrences of an i.i.d. stream drawn from p and renormalises -- i.e. successive
// sampling without replacement (Plackett-Luce).  The same distribution, order included, is obtained by giving every
// candidate an exponential clock key = -log(u) / p and taking the `size` smallest keys in ascending order
// (Efraimidis-Spirakis); that is what the kernel does, with a counter-based Philox4x32-10 stream per (seed, cell,
// candidate).  The RANDOM STREAM differs from NumPy's MT19937 -- which is why the default backend stays the host one
// (bit-equal `sampling_ixs` / `delta_S_rndm`, tested against the reference's golden vectors).
#include "velo_common.cuh"

namespace velo {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x)   // murmur3 finaliser
{
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

// one CTA per cell: keys for the W candidates, bitonic sort of (key, position) in shared memory, first m win
__global__ void __launch_bounds__(1024) k_sample_neighbors(const int32_t *__restrict__ knn_idx, int64_t C, int W,
                                                           const float *__restrict__ inv_p, int m, uint64_t seed, int Np,
                                                           int32_t *__restrict__ neigh_ixs, int32_t *__restrict__ sampling_ixs)
{
    extern __shared__ unsigned long long keys[];
    const int64_t cell = blockIdx.x;
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    for (int j4 = threadIdx.x; 4 * j4 < Np; j4 += blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(cell), static_cast<uint32_t>(cell >> 32),
                                                 static_cast<uint32_t>(j4), 0x53414d50u /* "SAMP" */), key);
        const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = 4 * j4 + k;
            unsigned long long packed = ~0ull;
            if (j < W) {
                const float u = (static_cast<float>(rv[k] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
                const float t = -__logf(u) * inv_p[j];                                           // exponential clock
                packed = (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | static_cast<uint32_t>(j);
            }
            keys[j] = packed;
        }
    }
    __syncthreads();
    for (int k = 2; k <= Np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (Np >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned long long a = keys[i], b = keys[l];
                if ((a > b) == ((i & k) == 0)) {
                    keys[i] = b;
                    keys[l] = a;
                }
            }
            __syncthreads();
        }
    for (int n = threadIdx.x; n < m; n += blockDim.x) {
        const int pos = static_cast<int>(keys[n] & 0xffffffffu);
        sampling_ixs[cell * m + n] = pos;
        neigh_ixs[cell * m + n] = knn_idx[cell * W + pos];
    }
}

// out[c, g] = +-in[pi_g(c), g]: pi_g = 4-round Feistel bijection on [0, 4^h) >= C keyed by (seed, g), cycle-walked into
// [0, C); the sign is an independent hash bit of (seed, g, c).
// Round function: the top h bits of an odd-constant product of (R + round key) -- every bit of R reaches them; one
// add, one multiply, one shift per round (round 1 hashed with a full murmur finaliser per round: ~100 instructions
// per element made the permutation kernel compute-bound at a tenth of the HBM rate).  Any round function gives a
// bijection; uniformity / independence of the resulting permutations is tested statistically (tests/test_host_logic.py).
__device__ __forceinline__ uint32_t feistel_walk(uint32_t x, const uint32_t (&rk)[4], int h, uint32_t mask, uint32_t C)
{
    do {
        uint32_t L = x >> h, R = x & mask;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t f = ((R + rk[r]) * 0x9E3779B1u) >> (32 - h);
            const uint32_t nl = R;
            R = L ^ f;
            L = nl;
        }
        x = (L << h) | R;
    } while (x >= C);
    return x;
}

// A per-gene permutation of the CELL axis is a 4-byte gather from a random row in the cell-major layout: one useful
// float per 32-byte sector (round 1: 338 GB/s, 5 % of the HBM rate at 50k x 30k).  It is done in the gene-major
// orientation instead: a block of genes is transposed to gene-major scratch (coalesced tile transpose), every gene row
// is permuted by ONE CTA that stages the whole row in shared memory with TMA bulk copies (200 KB at 50k cells; rows
// beyond the shared-memory limit gather from the L2-resident row), and the block is transposed back.  Three
// streaming passes (6 x the matrix in traffic) instead of one pass that wastes 7/8 of every sector.
__global__ void __launch_bounds__(1024) k_permute_gene_rows(const float *__restrict__ in_gm, float *__restrict__ out_gm,
                                                           int64_t C, int64_t g_first, uint64_t seed, int h, int in_smem)
{
    extern __shared__ __align__(128) unsigned char perm_smem[];
    float *s_row = reinterpret_cast<float *>(perm_smem);
    __shared__ uint64_t bar;
    const int64_t gl = blockIdx.x;
    const float *row = in_gm + gl * C;
    if (in_smem) {
        const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) && (C % 4 == 0);
        if (aligned) {
            if (threadIdx.x == 0) {
                mbar_init(&bar, 1);
                mbar_fence_init();
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const uint32_t total = static_cast<uint32_t>(C * 4);
                mbar_expect_tx(&bar, total);
                for (uint32_t off = 0; off < total; off += 32768u) {
                    const uint32_t n = total - off < 32768u ? total - off : 32768u;
                    tma_load_1d(reinterpret_cast<unsigned char *>(s_row) + off, reinterpret_cast<const unsigned char *>(row) + off, n, &bar);
                }
            }
            mbar_wait(&bar, 0);
        } else {
            for (int64_t c = threadIdx.x; c < C; c += blockDim.x) s_row[c] = row[c];
            __syncthreads();
        }
        row = s_row;
    }
    const int64_t g = g_first + gl;
    const uint32_t mask = (1u << h) - 1u;
    const uint32_t s0 = static_cast<uint32_t>(seed), s1 = static_cast<uint32_t>(seed >> 32);
    const uint32_t kg = mix32(static_cast<uint32_t>(g) ^ s0) + s1;
    const uint32_t sk = mix32(kg ^ 0x5bd1e995u);
    uint32_t rk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rk[r] = mix32(kg + 0x9E3779B9u * (r + 1u));          // per-gene round keys, hoisted
    float *orow = out_gm + gl * C;
    for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
        const uint32_t x = feistel_walk(static_cast<uint32_t>(c), rk, h, mask, static_cast<uint32_t>(C));
        const uint32_t sbit = ((static_cast<uint32_t>(c) * 0x9E3779B1u ^ sk) * 0x85EBCA6Bu) >> 31;
        const float v = row[x];
        orow[c] = sbit ? -v : v;
    }
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_sample_neighbors(const int32_t *knn_idx, int64_t C, int W, const float *inv_p, int m,
                                         uint64_t seed, int32_t *neigh_ixs, int32_t *sampling_ixs, velo_stream_t stream)
{
    VELO_REQUIRE(knn_idx && inv_p && neigh_ixs && sampling_ixs, "sample_neighbors: null pointer");
    VELO_REQUIRE(C >= 0 && W > 0 && m >= 0 && m <= W, "sample_neighbors: need 0 <= m <= W");
    VELO_REQUIRE(C < (1LL << 31), "sample_neighbors: too many cells for one launch");
    if (C == 0 || m == 0) return VELO_OK;
    int Np = 64;
    while (Np < W) Np <<= 1;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const size_t smem = static_cast<size_t>(Np) * 8;
    VELO_REQUIRE(smem <= static_cast<size_t>(dp.smem_optin), "sample_neighbors: at most %d candidates per cell (got %d)",
                 dp.smem_optin / 8, W);
    VELO_CUDA_TRY(cudaFuncSetAttribute(k_sample_neighbors, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    const int threads = Np / 2 >= 1024 ? 1024 : (Np / 2 < 64 ? 64 : Np / 2);
    k_sample_neighbors<<<static_cast<unsigned>(C), threads, smem, as_stream(stream)>>>(knn_idx, C, W, inv_p, m, seed, Np,
                                                                                     neigh_ixs, sampling_ixs);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_permute_rows_nsign(const float *in_cm, float *out_cm, int64_t ld, int64_t G, int64_t C,
                                           uint64_t seed, velo_stream_t stream)
{
    VELO_REQUIRE(in_cm && out_cm && in_cm != out_cm && G > 0 && C > 0 && ld >= G, "permute_rows_nsign: bad arguments");
    VELO_REQUIRE(C < (1LL << 30), "permute_rows_nsign: too many cells");
    int h = 1;
    while ((1LL << (2 * h)) < C) ++h;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    // gene blocks sized so that the two gene-major scratch copies stay around 1 GB (L2-friendly, bounded workspace)
    int64_t gb = (1LL << 27) / C;
    gb = gb < 32 ? 32 : gb / 32 * 32;
    if (gb > G) gb = G;
    float *a = nullptr, *b = nullptr;
    VELO_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&a), static_cast<size_t>(gb * C) * 4, st));
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&b), static_cast<size_t>(gb * C) * 4, st);
    if (e != cudaSuccess) {
        cudaFreeAsync(a, st);
        VELO_CUDA_TRY(e);
    }
    const size_t row_bytes = static_cast<size_t>(C) * 4;
    const int in_smem = row_bytes + 1024 <= static_cast<size_t>(dp.smem_optin) ? 1 : 0;
    const size_t smem = in_smem ? (row_bytes + 15) / 16 * 16 : 0;
    if (in_smem) e = cudaFuncSetAttribute(k_permute_gene_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    for (int64_t g0 = 0; g0 < G && e == cudaSuccess && rc == VELO_OK; g0 += gb) {
        const int64_t ng = G - g0 < gb ? G - g0 : gb;
        if ((rc = velo_dev_unpack_genemajor(in_cm + g0, ld, ng, C, a, 4, stream))) break;
        k_permute_gene_rows<<<static_cast<unsigned>(ng), 1024, smem, st>>>(a, b, C, g0, seed, h, in_smem);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        rc = velo_dev_pack_cellmajor(b, 4, ng, C, out_cm, ld, g0, stream);
    }
    cudaFreeAsync(a, st);
    cudaFreeAsync(b, st);
    if (rc) return rc;
    VELO_CUDA_TRY(e);
    return VELO_OK;
}
