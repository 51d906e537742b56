/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the colDeltaCor family.
 *
 * Plain-C (fp64, OpenMP) restatement of the algorithm of the reference's only
 * native component, velocyto/speedboosted.pyx (six `cdef ... nogil` loops,
 * speedboosted.pyx:13-538).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py may load this library; the product
 * (velocyto.py_b200 / libvelo_b200.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_pinning.py checks this file against
 * (a) golden vectors produced by the unmodified reference (compiled from the
 *     .pyx by oracle/build_ref.sh, driven by tests/golden/make_golden.py) and
 * (b) the compiled reference itself (oracle/_ref) when it is present.
 *
 * What the reference computes, for every cell (column) c of the gene x cell
 * matrices e, d (row-major, rows = genes):
 *
 *   for every target cell i  (all cells: "full"; i = ixs[c, n]: "partial")
 *       A[g]   = f(e[g, i] - e[g, c])                      pyx:27-29, 279-282
 *       muA    = mean_g A[g]                               pyx:32-38
 *       b[g]   = d[g, c] - mean_g d[g, c]                  pyx:46-55
 *       rm[c, i] += sum_g ((A[g]-muA) * 1/sqrt(sum (A-muA)^2))
 *                        * (b[g]      * 1/sqrt(sum b^2))   pyx:57-79
 *
 * with the per-variant transform f and its zero rule:
 *   full    linear : f(t) = t                                        pyx:29
 *   full    sqrt   : t > 0 ? sqrt(t+psc)  : -sqrt(-t+psc)            pyx:110-114
 *   full    log10  : t > 0 ? log10(t+psc) : -log10(-t+psc)           pyx:195-199
 *   partial linear : f(t) = t                                        pyx:282
 *   partial sqrt   : |t|<1e-16 ? 0 : t>0 ? sqrt(t+psc):-sqrt(-t+psc) pyx:372-378
 *   partial log10  : t >= 0 ? log10(t+psc) : -log10(-t+psc)          pyx:470-473
 *
 * The output is ACCUMULATED (+=) into caller-zeroed rm (pyx:78, 336), so a
 * duplicated index in a row of ixs counts twice, exactly as in the reference.
 * Degenerate columns (zero variance) give 1/sqrt(0)=inf and 0*inf = NaN, as in
 * the reference; callers patch NaN->1 and the diagonal->0 (analysis.py:1604-1612).
 *
 * Deliberate differences from the reference (none change the arithmetic):
 *   - 64-bit indexing (the reference's C `int` products overflow once
 *     genes*cells >= 2^31, speedboosted.pyx:24,29,78);
 *   - scratch is one gene-long column per thread instead of a genes x targets
 *     block per thread (the per-(c,i) operation order over genes is the same);
 *   - no -ffast-math, so results are reproducible run to run.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { VO_LINEAR = 0, VO_SQRT = 1, VO_LOG10 = 2 };

static inline double vo_transform(double t, int transform, int partial, double psc)
{
    switch (transform) {
    case VO_SQRT:
        if (partial && fabs(t) < 1e-16) return 0.0;
        return t > 0 ? sqrt(t + psc) : -sqrt(-t + psc);
    case VO_LOG10:
        if (partial) return t >= 0 ? log10(t + psc) : -log10(-t + psc);
        return t > 0 ? log10(t + psc) : -log10(-t + psc);
    default:
        return t;
    }
}

/*
 * e, d : rows x cols, row-major (genes x cells), fp64
 * rm   : cols x cols, row-major, caller-zeroed, accumulated into
 * ixs  : cols x nrndm (int64) or NULL for the full variants
 * returns 0, or 1 on allocation failure.
 */
int velo_oracle_coldeltacor(const double *e, const double *d, double *rm,
                            const int64_t *ixs, int64_t rows, int64_t cols,
                            int64_t nrndm, int transform, int num_threads,
                            double psc)
{
    const int partial = ixs != NULL;
    const int64_t ntargets = partial ? nrndm : cols;
    int failed = 0;
#ifdef _OPENMP
    if (num_threads < 1) num_threads = 1;
#pragma omp parallel num_threads(num_threads)
#endif
    {
        double *col = (double *)malloc((size_t)rows * sizeof(double));
        double *b = (double *)malloc((size_t)rows * sizeof(double));
        if (!col || !b) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            failed = 1;
        } else {
            int64_t c;
#ifdef _OPENMP
#pragma omp for schedule(guided)
#endif
            for (c = 0; c < cols; ++c) {
                /* velocity column: centre and inverse norm (pyx:46-55, 67-72) */
                double mub = 0.0, ssb = 0.0;
                for (int64_t g = 0; g < rows; ++g) mub += d[g * cols + c];
                mub /= (double)rows;
                for (int64_t g = 0; g < rows; ++g) {
                    b[g] = d[g * cols + c] - mub;
                    ssb += b[g] * b[g];
                }
                ssb = 1.0 / sqrt(ssb);

                for (int64_t n = 0; n < ntargets; ++n) {
                    const int64_t i = partial ? ixs[c * nrndm + n] : n;
                    double mu = 0.0, ss = 0.0, acc;
                    for (int64_t g = 0; g < rows; ++g) {
                        col[g] = vo_transform(e[g * cols + i] - e[g * cols + c],
                                              transform, partial, psc);
                        mu += col[g];
                    }
                    mu /= (double)rows;
                    for (int64_t g = 0; g < rows; ++g) {
                        col[g] -= mu;
                        ss += col[g] * col[g];
                    }
                    ss = 1.0 / sqrt(ss);
                    acc = rm[c * cols + i];
                    for (int64_t g = 0; g < rows; ++g)
                        acc += (col[g] * ss) * (b[g] * ssb);
                    rm[c * cols + i] = acc;
                }
            }
        }
        free(col);
        free(b);
    }
    return failed;
}

/*
 * The same arithmetic for a FEW selected cells of a large problem (the bench-shape spot checks: 30 000 genes,
 * 3 000 neighbours, a handful of cells) -- identical per-pair operation order (pyx:275-336), compact output, and
 * the thread loop over the (cell, neighbour) PAIRS instead of over cells.
 *
 * e        : rows x cols, row-major (the cells involved and all their neighbours, gathered by the caller)
 * d_sel    : rows x ncells, row-major -- velocity columns of the selected cells only
 * cells    : ncells column ids (into e) of the selected cells
 * ixs_sel  : ncells x nrndm neighbour column ids (into e)
 * out      : ncells x nrndm, out[k*nrndm + n] = corr(cells[k], ixs_sel[k, n])   (assigned, not accumulated)
 * partial  : zero rule of the partial (1) or full (0) loops
 * Pinned against velo_oracle_coldeltacor in tests/test_oracle_pinning.py.
 */
int velo_oracle_coldeltacor_cells(const double *e, const double *d_sel, double *out,
                                  const int64_t *ixs_sel, const int64_t *cells, int64_t rows,
                                  int64_t cols, int64_t ncells, int64_t nrndm, int transform,
                                  int partial, int num_threads, double psc)
{
    int failed = 0;
    double *ball = (double *)malloc((size_t)rows * (size_t)ncells * sizeof(double));
    double *ssb = (double *)malloc((size_t)ncells * sizeof(double));
    if (!ball || !ssb) {
        free(ball);
        free(ssb);
        return 1;
    }
    for (int64_t k = 0; k < ncells; ++k) {          /* pyx:300-309, 323-329 */
        double mub = 0.0, q = 0.0;
        for (int64_t g = 0; g < rows; ++g) mub += d_sel[g * ncells + k];
        mub /= (double)rows;
        for (int64_t g = 0; g < rows; ++g) {
            const double b = d_sel[g * ncells + k] - mub;
            ball[k * rows + g] = b;
            q += b * b;
        }
        ssb[k] = 1.0 / sqrt(q);
    }
#ifdef _OPENMP
    if (num_threads < 1) num_threads = 1;
#pragma omp parallel num_threads(num_threads)
#endif
    {
        double *col = (double *)malloc((size_t)rows * sizeof(double));
        if (!col) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            failed = 1;
        } else {
            int64_t t;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (t = 0; t < ncells * nrndm; ++t) {
                const int64_t k = t / nrndm, c = cells[k], i = ixs_sel[t];
                const double *b = ball + k * rows;
                double mu = 0.0, ss = 0.0, acc = 0.0;
                for (int64_t g = 0; g < rows; ++g) {
                    col[g] = vo_transform(e[g * cols + i] - e[g * cols + c], transform, partial, psc);
                    mu += col[g];
                }
                mu /= (double)rows;
                for (int64_t g = 0; g < rows; ++g) {
                    col[g] -= mu;
                    ss += col[g] * col[g];
                }
                ss = 1.0 / sqrt(ss);
                for (int64_t g = 0; g < rows; ++g) acc += (col[g] * ss) * (b[g] * ssb[k]);
                out[t] = acc;
            }
        }
        free(col);
    }
    free(ball);
    free(ssb);
    return failed;
}

int velo_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
