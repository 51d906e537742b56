// Host-side neighbour sampler with NumPy's legacy random stream, bit for bit.
//
// estimate_transition_prob (velocyto/analysis.py:1529, 1552-1566) seeds NumPy's global MT19937 and then calls
//     np.random.choice(n_neighbors + 1, size=int(frac * (n_neighbors + 1)), replace=False, p=p)
// once per cell in a Python loop.  `sampling_ixs` -- and with it every correlation downstream -- is reproducible against
// the reference only if that exact stream is consumed in that exact way, so the default backend keeps it on the host.
// The loop costs ~1 ms of interpreter and NumPy overhead per cell (12 s at 10k cells, minutes at 100k); this file
// restates the algorithm of RandomState.choice(replace=False, p=...) (numpy/random/mtrand.pyx, legacy stream) in C++:
//     repeat: draw (size - found) doubles; zero the probabilities of what was found; cdf = cumsum(p) / cdf[-1];
//             new = searchsorted(cdf, x, side="right"); keep first occurrences in draw order
// The draw count of a cell depends on its collisions, hence one sequential stream over all cells; the work per
// iteration is a sequential fp64 prefix sum (same operation order as np.cumsum) and a vectorisable division.
// No device code here: this is part of the host logic the reference also runs on the host.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/velo_b200.h"

namespace {

struct MT19937 {
    uint32_t key[624];
    int pos;
    void seed(uint32_t s)
    {   // numpy/random/src/mt19937/mt19937.c: mt19937_seed
        for (int i = 0; i < 624; ++i) {
            key[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + static_cast<uint32_t>(i) + 1u;
        }
        pos = 624;
    }
    void gen()
    {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; ++i) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        }
        for (; i < 623; ++i) {
            y = (key[i] & UPPER) | (key[i + 1] & LOWER);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        }
        y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        pos = 0;
    }
    inline uint32_t next32()
    {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline double next_double()
    {   // random_double of the legacy generator: 53 bits from two draws
        const int32_t a = static_cast<int32_t>(next32() >> 5), b = static_cast<int32_t>(next32() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

}  // namespace

extern "C" int velo_host_sample_neighbors_numpy(uint32_t seed, int64_t n_cells, int W, const double *p, int size,
                                                int64_t *sampling_ixs, uint32_t *mt_key_out, int *mt_pos_out)
{
    if (!p || !sampling_ixs || n_cells < 0 || W <= 0 || size < 0 || size > W) return VELO_E_INVALID;
    int positive = 0;
    for (int i = 0; i < W; ++i) positive += p[i] > 0;
    if (positive < size) return VELO_E_INVALID;            // "Fewer non-zero entries in p than size"
    MT19937 rng;
    rng.seed(seed);
    std::vector<double> pw(W), cdf(W), cdf0(W), x(size);
    std::vector<int64_t> stamp(W, -1);
    // the first iteration of every cell sees the untouched p: one shared cdf
    {
        double s = 0.0;
        for (int i = 0; i < W; ++i) {
            s += p[i];
            cdf0[i] = s;
        }
        const double tot = cdf0[W - 1];
        for (int i = 0; i < W; ++i) cdf0[i] /= tot;
    }
    // guide table over the shared first-iteration cdf: x in [b/B, (b+1)/B) can only land in [guide[b], guide[b+1]]
    // (the search inside that range is the same exact comparison, just without ~10 cache-missing bisection steps)
    const int B = 1 << 14;
    std::vector<int32_t> guide(B + 1);
    for (int b = 0; b <= B; ++b)
        guide[b] = static_cast<int32_t>(std::upper_bound(cdf0.begin(), cdf0.end(), static_cast<double>(b) / B) - cdf0.begin());
    int64_t batch = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
        int64_t *found = sampling_ixs + c * size;
        int n_uniq = 0;
        bool dirty = false;
        while (n_uniq < size) {
            const int k = size - n_uniq;
            for (int j = 0; j < k; ++j) x[j] = rng.next_double();
            const double *cd = cdf0.data();
            if (n_uniq > 0) {
                if (!dirty) {
                    memcpy(pw.data(), p, sizeof(double) * W);
                    dirty = true;
                }
                for (int j = 0; j < n_uniq; ++j) pw[found[j]] = 0.0;
                double s = 0.0;
                for (int i = 0; i < W; ++i) {                 // np.cumsum: sequential fp64 adds
                    s += pw[i];
                    cdf[i] = s;
                }
                const double tot = cdf[W - 1];
                for (int i = 0; i < W; ++i) cdf[i] /= tot;
                cd = cdf.data();
            }
            ++batch;
            for (int j = 0; j < k; ++j) {
                int64_t idx;
                if (cd == cdf0.data()) {
                    const int b = static_cast<int>(x[j] * B);                     // x in [0, 1): b in [0, B)
                    int lo = guide[b], hi = guide[b + 1] < W ? guide[b + 1] + 1 : W;
                    idx = std::upper_bound(cd + lo, cd + hi, x[j]) - cd;
                } else {
                    idx = std::upper_bound(cd, cd + W, x[j]) - cd;                // searchsorted(side="right")
                }
                if (idx < W && stamp[idx] != batch) {         // np.unique(return_index) + sort: first occurrences, draw order
                    stamp[idx] = batch;
                    found[n_uniq++] = idx;
                }
            }
        }
    }
    if (mt_key_out) memcpy(mt_key_out, rng.key, sizeof(rng.key));
    if (mt_pos_out) *mt_pos_out = rng.pos;
    return VELO_OK;
}
