// Exact brute-force k nearest neighbours in a low-dimensional space (PCA coordinates for knn_imputation,
// analysis.py:989-1005 / neighbors.py:239-243,363-376; the 2-D embedding for estimate_transition_prob,
// analysis.py:1547-1549).  Replaces scikit-learn's KD-tree search, which at 100k cells and k = 10^4 takes minutes
// on the host.  SURVEY.md 8f item 2.
//
// One CTA per query cell (persistent grid).  The points (C x D fp64) stay L2 resident; per query:
//   1. squared Euclidean distances in fp64, their order-preserving fp32 keys to a per-CTA scratch row;
//   2. 8-bit-per-pass radix select of the k-th smallest key (shared-memory histograms);
//   3. compaction of the candidates (key <= threshold; a handful more than k when fp32 keys tie), with their exact
//      fp64 distances, into shared memory;
//   4. bitonic sort by (distance, index) and write-out of the k nearest in ascending order.
// Distances are evaluated directly as sum (x_q - x_j)^2 (no |x|^2+|y|^2-2xy cancellation), in fp64 like the reference.
#include "velo_common.cuh"

namespace velo {

__device__ __forceinline__ uint32_t knn_f2key(float f)   // f >= 0
{
    return __float_as_uint(f);                            // non-negative floats order like their bit patterns
}

struct KnnParams {
    const double *X;       // C x D
    const double *Q;       // query points (nq x D) when they are not points of X, else nullptr
    int64_t C, q0, nq;     // points; query range [q0, q0 + nq)
    int D, k, include_self, P;
    int32_t *out_idx;      // C x k
    double *out_dist;      // C x k or null
    uint32_t *scratch;     // gridDim.x x C
    int *overflow;         // internal consistency flag of the tie-resolving path (stays 0)
};

// QT queries per CTA pass: the distance sweep reads every point ONCE for the whole tile of queries (at 500k points x
// 20-D the sweep is 80 MB of L2 traffic per query -- 40 TB for all queries, the whole cost of the search -- so a tile
// of 4 cuts the search time almost 4x); selection, compaction and the sort then run per query on its own key row.
template <int QT>
__global__ void __launch_bounds__(512) k_knn_bruteforce(const KnnParams p)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double *c_key = reinterpret_cast<double *>(sm_raw);                  // P
    int32_t *c_idx = reinterpret_cast<int32_t *>(c_key + p.P);           // P
    double *xq_all = reinterpret_cast<double *>(c_idx + p.P);            // QT x D (P even -> 8-byte aligned)
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long sel[3];
    __shared__ unsigned int n_cand;
    const int tid = threadIdx.x, nt = blockDim.x;
    uint32_t *keys_all = p.scratch + static_cast<int64_t>(blockIdx.x) * p.C * QT;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const int64_t q_end = p.q0 + p.nq;

    for (int64_t qb = p.q0 + static_cast<int64_t>(blockIdx.x) * QT; qb < q_end; qb += static_cast<int64_t>(gridDim.x) * QT) {
        for (int t = tid; t < QT * p.D; t += nt) {
            const int qi = t / p.D, d = t - qi * p.D;
            const int64_t q = min(qb + qi, q_end - 1);                    // ragged last tile: repeat the last query
            xq_all[t] = p.Q ? p.Q[(q - p.q0) * p.D + d] : p.X[q * p.D + d];
        }
        __syncthreads();
        // 1. distances -> keys, one read of every point for the QT queries of the tile
        for (int64_t j = tid; j < p.C; j += nt) {
            double d2[QT];
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) d2[qi] = 0.0;
            const double *xj = p.X + j * p.D;
            for (int d = 0; d < p.D; ++d) {
                const double x = xj[d];
#pragma unroll
                for (int qi = 0; qi < QT; ++qi) {
                    const double t = x - xq_all[qi * p.D + d];
                    d2[qi] = fma(t, t, d2[qi]);
                }
            }
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) {
                const bool skip = !p.include_self && j == qb + qi;
                keys_all[static_cast<int64_t>(qi) * p.C + j] = skip ? 0x7f800000u : knn_f2key(static_cast<float>(d2[qi]));
            }
        }
        __syncthreads();
      for (int qi = 0; qi < QT && qb + qi < q_end; ++qi) {
        const int64_t q = qb + qi;
        const uint32_t *keys = keys_all + static_cast<int64_t>(qi) * p.C;
        const double *xq = xq_all + qi * p.D;
        // 2. radix select of the k-th smallest key (0-based rank k-1)
        uint32_t prefix = 0, mask = 0;
        int64_t rem = p.k - 1;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int b = tid; b < 256; b += nt) hist[b] = 0;
            __syncthreads();
            for (int64_t j = tid; j < p.C; j += nt) {
                const uint32_t u = keys[j];
                if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                int64_t cum = 0;
                int b = 0;
                for (; b < 255; ++b) {
                    if (cum + hist[b] > rem) break;
                    cum += hist[b];
                }
                sel[0] = static_cast<unsigned long long>(b);
                sel[1] = static_cast<unsigned long long>(rem - cum);
            }
            __syncthreads();
            prefix |= static_cast<uint32_t>(sel[0]) << shift;
            mask |= 255u << shift;
            rem = static_cast<int64_t>(sel[1]);
            __syncthreads();
        }
        // 3. candidates: every key <= threshold, with exact fp64 distances
        if (tid == 0) n_cand = 0;
        for (int i = tid; i < p.P; i += nt) {
            c_key[i] = inf;
            c_idx[i] = 0x7fffffff;
        }
        __syncthreads();
        for (int64_t j = tid; j < p.C; j += nt) {
            if (keys[j] <= prefix && (p.include_self || j != q)) {
                const unsigned int slot = atomicAdd(&n_cand, 1u);
                if (slot < static_cast<unsigned int>(p.P)) {
                    double d2 = 0.0;
                    const double *xj = p.X + j * p.D;
                    for (int d = 0; d < p.D; ++d) {
                        const double t = xj[d] - xq[d];
                        d2 = fma(t, t, d2);
                    }
                    c_key[slot] = d2;
                    c_idx[slot] = static_cast<int32_t>(j);
                }
            }
        }
        __syncthreads();
        if (n_cand > static_cast<unsigned int>(p.P)) {
            // More points tie with the k-th distance (same fp32 key) than the candidate buffer has room for: in
            // practice exact duplicates (identical cells in PCA / embedding space).  Such ties are interchangeable,
            // so they are resolved by LOWEST INDEX: a second radix select, over the point index among the tied keys,
            // finds the index bound T below which exactly rem + 1 ties lie (rem = rank of the k-th neighbour inside
            // the tie group, left over from the key select); the candidate set is then exactly k points.
            __syncthreads();
            uint32_t ipre = 0, imask = 0;
            int64_t irem = rem;
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                for (int b = tid; b < 256; b += nt) hist[b] = 0;
                __syncthreads();
                for (int64_t j = tid; j < p.C; j += nt)
                    if (keys[j] == prefix && (static_cast<uint32_t>(j) & imask) == ipre)
                        atomicAdd(&hist[(static_cast<uint32_t>(j) >> shift) & 255u], 1u);
                __syncthreads();
                if (tid == 0) {
                    int64_t cum = 0;
                    int b = 0;
                    for (; b < 255; ++b) {
                        if (cum + hist[b] > irem) break;
                        cum += hist[b];
                    }
                    sel[0] = static_cast<unsigned long long>(b);
                    sel[1] = static_cast<unsigned long long>(irem - cum);
                }
                __syncthreads();
                ipre |= static_cast<uint32_t>(sel[0]) << shift;
                imask |= 255u << shift;
                irem = static_cast<int64_t>(sel[1]);
                __syncthreads();
            }
            if (tid == 0) n_cand = 0;
            for (int i = tid; i < p.P; i += nt) {
                c_key[i] = inf;
                c_idx[i] = 0x7fffffff;
            }
            __syncthreads();
            for (int64_t j = tid; j < p.C; j += nt) {
                const uint32_t u = keys[j];
                if (u < prefix || (u == prefix && static_cast<uint32_t>(j) <= ipre)) {
                    const unsigned int slot = atomicAdd(&n_cand, 1u);
                    if (slot < static_cast<unsigned int>(p.P)) {
                        double d2 = 0.0;
                        const double *xj = p.X + j * p.D;
                        for (int d = 0; d < p.D; ++d) {
                            const double t = xj[d] - xq[d];
                            d2 = fma(t, t, d2);
                        }
                        c_key[slot] = d2;
                        c_idx[slot] = static_cast<int32_t>(j);
                    }
                }
            }
            __syncthreads();
            if (tid == 0 && n_cand != static_cast<unsigned int>(p.k)) atomicExch(p.overflow, 1);   // cannot happen
        }
        // 4. bitonic sort by (distance, index)
        for (int kk = 2; kk <= p.P; kk <<= 1)
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                for (int t = tid; t < p.P; t += nt) {
                    const int l = t ^ jj;
                    if (l > t) {
                        const double a = c_key[t], b = c_key[l];
                        const int32_t ia = c_idx[t], ib = c_idx[l];
                        const bool gt = a > b || (a == b && ia > ib);
                        if (gt == ((t & kk) == 0)) {
                            c_key[t] = b; c_key[l] = a;
                            c_idx[t] = ib; c_idx[l] = ia;
                        }
                    }
                }
                __syncthreads();
            }
        for (int r = tid; r < p.k; r += nt) {
            p.out_idx[(q - p.q0) * p.k + r] = c_idx[r];
            if (p.out_dist) p.out_dist[(q - p.q0) * p.k + r] = sqrt(c_key[r]);
        }
        __syncthreads();
      }
    }
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_knn(const double *X, int64_t C, int D, int k, int include_self, int32_t *out_idx,
                            double *out_dist, velo_stream_t stream)
{
    return velo_dev_knn_range(X, C, D, k, include_self, 0, C, out_idx, out_dist, stream);
}

static int knn_launch(const double *X, const double *Q, int64_t C, int D, int k, int include_self, int64_t q0, int64_t nq,
                      int32_t *out_idx, double *out_dist, velo_stream_t stream);

extern "C" int velo_dev_knn_range(const double *X, int64_t C, int D, int k, int include_self, int64_t q0, int64_t nq,
                                  int32_t *out_idx, double *out_dist, velo_stream_t stream)
{
    VELO_REQUIRE(q0 >= 0 && nq >= 0 && q0 + nq <= C, "knn: query range outside [0, %lld)", static_cast<long long>(C));
    return knn_launch(X, nullptr, C, D, k, include_self, q0, nq, out_idx, out_dist, stream);
}

extern "C" int velo_dev_knn_query(const double *X, int64_t C, int D, const double *Q, int64_t nq, int k,
                                  int32_t *out_idx, double *out_dist, velo_stream_t stream)
{
    VELO_REQUIRE(Q && nq >= 0, "knn_query: bad arguments");
    return knn_launch(X, Q, C, D, k, 1, 0, nq, out_idx, out_dist, stream);
}

static int knn_launch(const double *X, const double *Q, int64_t C, int D, int k, int include_self, int64_t q0, int64_t nq,
                      int32_t *out_idx, double *out_dist, velo_stream_t stream)
{
    VELO_REQUIRE(X && out_idx && C > 0 && D > 0 && D <= 4096, "knn: bad arguments");
    if (nq == 0) return VELO_OK;
    VELO_REQUIRE(k > 0 && k <= (include_self ? C : C - 1), "knn: k = %d out of range for %lld points", k,
                 static_cast<long long>(C));
    VELO_REQUIRE(C < 2147483647LL, "knn: too many points");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    int P = 64;
    while (P < k + (k >> 3) + 32) P <<= 1;                 // room for fp32-key ties at the threshold
    // query tile: 4 queries share one sweep over the points when their coordinates fit next to the candidate buffers
    // (a tile of 8 was measured and is SLOWER: 3.96 s vs 3.77 s at 500k x 20-D, 154 vs 96 ms at 100k x 2-D with k = 3000 --
    // the per-CTA key scratch doubles and the select passes, not the sweep, dominate by then)
    const int QT = (D <= 64 && nq >= 4) ? 4 : 1;
    const size_t smem = static_cast<size_t>(P) * 12 + static_cast<size_t>(D) * 8 * QT + 16;
    VELO_REQUIRE(smem + 2048 <= static_cast<size_t>(dp.smem_optin), "knn: k = %d too large for shared memory (max ~14000)", k);
    cudaStream_t st = as_stream(stream);
    auto kern = QT == 4 ? k_knn_bruteforce<4> : k_knn_bruteforce<1>;
    VELO_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 1;
    VELO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 512, smem));
    int64_t grid = static_cast<int64_t>(dp.sm_count) * (per_sm > 0 ? per_sm : 1);
    const int64_t tiles = (nq + QT - 1) / QT;
    if (grid > tiles) grid = tiles;
    uint32_t *scratch = nullptr;
    int *overflow = nullptr;
    VELO_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&scratch), static_cast<size_t>(grid) * C * 4 * QT, st));
    VELO_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&overflow), sizeof(int), st));
    VELO_CUDA_TRY(cudaMemsetAsync(overflow, 0, sizeof(int), st));
    KnnParams p;
    p.X = X; p.Q = Q; p.C = C; p.q0 = q0; p.nq = nq; p.D = D; p.k = k; p.include_self = include_self; p.P = P;
    p.out_idx = out_idx; p.out_dist = out_dist; p.scratch = scratch; p.overflow = overflow;
    kern<<<static_cast<unsigned>(grid), 512, smem, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    int flag = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, overflow, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(scratch, st);
    cudaFreeAsync(overflow, st);
    VELO_CUDA_TRY(e);
    VELO_REQUIRE(flag == 0, "knn: internal error in the tie-resolving selection (k = %d)", k);
    return VELO_OK;
}
