/*
 * velo_b200.h -- C ABI of libvelo_b200.so, the B200-native (sm_100a) numerical
 * core behind VelocytoLoom.{knn_imputation, fit_gammas, predict_U,
 * estimate_transition_prob} of velocyto.py 0.17.16.
 *
 * Two tiers of entry points:
 *
 *  (1) HOST drop-ins  (velo_colDeltaCor*): same argument lists as the reference's
 *      native functions `x_colDeltaCor*` (velocyto/speedboosted.pyx:13-538, reached
 *      through the Python-callable `_colDeltaCor*` at speedboosted.pyx:542-610) --
 *      host pointers, gene-major (rows = genes) row-major fp64 matrices, output
 *      ACCUMULATED into the caller-zeroed dense cells x cells fp64 `rm`.  Sizes are
 *      64-bit (the reference's C `int` overflows at genes*cells >= 2^31).
 *      `num_threads` is accepted and ignored.  These are what a maintainer binds in
 *      place of `velocyto.speedboosted`.
 *
 *  (2) DEVICE tier (velo_dev_*): device pointers + a CUDA stream; matrices are
 *      CELL-MAJOR fp32 (`x[cell * ld + gene]`, ld % 4 == 0, pad columns zero), the
 *      layout the kernels are designed around (a cell's expression profile is one
 *      contiguous, 16-byte aligned row).  Outputs are compact (cells x m) instead of
 *      dense cells x cells.  The Python host layer (velocyto.py_b200) keeps data
 *      resident in HBM between the four methods through this tier.
 *
 * Every function returns 0 on success, or a negative VELO_E_* code; the message of
 * the last failure on the calling thread is available from velo_last_error().
 * There is no CPU fallback anywhere in this library.
 */
#ifndef VELO_B200_H
#define VELO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VELO_ABI_VERSION 1

/* status codes */
#define VELO_OK            0
#define VELO_E_INVALID    -1   /* bad argument (shape, alignment, enum) */
#define VELO_E_CUDA       -2   /* CUDA runtime error (see velo_last_error) */
#define VELO_E_NODEVICE   -3   /* no sm_100-class device visible */
#define VELO_E_NOMEM      -4   /* device allocation failed */

/* transform applied to the expression difference (the three reference families) */
#define VELO_LINEAR 0          /* colDeltaCor       / colDeltaCorpartial        */
#define VELO_SQRT   1          /* colDeltaCorSqrt   / colDeltaCorSqrtpartial    */
#define VELO_LOG10  2          /* colDeltaCorLog10  / colDeltaCorLog10partial   */

/* zero rule: the reference's full and partial loops treat t == 0 differently
 * (speedboosted.pyx:110-114,195-199 vs :372-378,470-473); see DESIGN.md. */
#define VELO_RULE_FULL    0
#define VELO_RULE_PARTIAL 1

typedef void *velo_stream_t;   /* a cudaStream_t; NULL = legacy default stream */

/* ---------------------------------------------------------------- library --- */
int         velo_abi_version(void);
const char *velo_last_error(void);
/* sm count, opt-in shared memory per block, total HBM bytes of the current device */
int         velo_device_info(int *sm_count, int *smem_optin_bytes, size_t *hbm_bytes,
                             int *cc_major, int *cc_minor);
/* give cached workspace (the stream-ordered pool this library allocates its scratch from) back to the driver */
int         velo_release_workspace(void);
/* number of kernel launches issued by this library in this process (bench evidence) */
uint64_t    velo_launch_count(void);

/* ------------------------------------------------- (1) host drop-in tier --- */
/* replaces x_colDeltaCor            speedboosted.pyx:13-87   (_colDeltaCor :542-550) */
int velo_colDeltaCor(const double *e, const double *d, double *rm,
                     int64_t rows, int64_t cols, int num_threads);
/* replaces x_colDeltaCorSqrt        speedboosted.pyx:93-172  (_colDeltaCorSqrt :552-561) */
int velo_colDeltaCorSqrt(const double *e, const double *d, double *rm,
                         int64_t rows, int64_t cols, int num_threads, double psc);
/* replaces x_colDeltaCorLog10       speedboosted.pyx:178-257 (_colDeltaCorLog10 :563-572) */
int velo_colDeltaCorLog10(const double *e, const double *d, double *rm,
                          int64_t rows, int64_t cols, int num_threads, double psc);
/* replaces x_colDeltaCorpartial     speedboosted.pyx:263-346 (_colDeltaCorpartial :574-584) */
int velo_colDeltaCorpartial(const double *e, const double *d, double *rm, const int64_t *ixs,
                            int64_t rows, int64_t cols, int64_t nrndm, int num_threads);
/* replaces x_colDeltaCorSqrtpartial speedboosted.pyx:352-443 (_colDeltaCorSqrtpartial :586-597) */
int velo_colDeltaCorSqrtpartial(const double *e, const double *d, double *rm, const int64_t *ixs,
                                int64_t rows, int64_t cols, int64_t nrndm, int num_threads,
                                double psc);
/* replaces x_colDeltaCorLog10partial speedboosted.pyx:449-538 (_colDeltaCorLog10partial :599-610) */
int velo_colDeltaCorLog10partial(const double *e, const double *d, double *rm, const int64_t *ixs,
                                 int64_t rows, int64_t cols, int64_t nrndm, int num_threads,
                                 double psc);

/* Same computation as the partial drop-ins, compact output: out[c*nrndm + n] =
 * corr(cell c, its n-th sampled neighbour), fp32, NaN where the reference gives NaN.
 * The dense cells x cells result of the reference does not exist at 100k cells (80 GB).
 * e, d: host, gene-major rows x cols, fp64 (elem_bytes = 8) or fp32 (elem_bytes = 4). */
int velo_colDeltaCorpartial_compact(int transform, const void *e, const void *d, int elem_bytes,
                                    const int64_t *ixs, float *out,
                                    int64_t rows, int64_t cols, int64_t nrndm, double psc);
/* The whole estimate_transition_prob numeric core in one host call: the compact correlations
 * above followed by the transition-probability epilogue (analysis.py:1604-1612, 1697-1698):
 * out[c*nrndm + n] = P(c -> ixs[c, n]).  sigma = sigma_corr of calculate_embedding_shift. */
int velo_transition_prob_partial(int transform, const void *e, const void *d, int elem_bytes,
                                 const int64_t *ixs, float *out,
                                 int64_t rows, int64_t cols, int64_t nrndm, double psc, double sigma);

/* Host-side neighbour sampler of estimate_transition_prob with NumPy's legacy random stream, bit for bit
 * (analysis.py:1529, 1552-1566): equivalent to
 *     np.random.seed(seed); np.stack([np.random.choice(W, size, replace=False, p=p) for _ in range(n_cells)])
 * sampling_ixs: n_cells x size int64 (host).  mt_key_out (624 words) / mt_pos_out: the generator state afterwards, so the
 * caller can leave np.random where the reference would have left it (either may be NULL).  No device work. */
int velo_host_sample_neighbors_numpy(uint32_t seed, int64_t n_cells, int W, const double *p, int size,
                                     int64_t *sampling_ixs, uint32_t *mt_key_out, int *mt_pos_out);

/* ---- cell-sharded host tier (one process per GPU) ----
 * The reference parallelises colDeltaCor*partial with an OpenMP prange over cells inside one process
 * (speedboosted.pyx:22-23); at box scale every rank owns a contiguous block of cells [c0, c0 + nc).
 *
 * velo_upload_cellmajor: host gene-major block -- `rows` rows of nc values, row pitch src_cols values (a column
 * block of the reference's rows x cols matrix) -- to nc cell-major fp32 DEVICE rows dst_cm (row stride ld), in
 * gene-row chunks through PCIe with the transpose/convert on the device.  Pageable host memory is staged through
 * the library's pinned ring by several host threads.  dst_lo_cm != NULL also writes the fp32 residuals of fp64 data
 * and sets *lo_nonzero (host int) to 1 when any is non-zero (then the call synchronises `stream`).
 * Otherwise the call is asynchronous on `stream`: a PAGE-LOCKED source is read by the copy engine until the work
 * queued here completes (do not modify or free it before synchronising `stream`); a pageable source has been
 * consumed when the call returns. */
int velo_upload_cellmajor(const void *src_gc, int elem_bytes, int64_t rows, int64_t nc, int64_t src_cols,
                          float *dst_cm, float *dst_lo_cm, int *lo_nonzero, int64_t ld, velo_stream_t stream);
/* velo_transition_prob_partial for the local cells of one rank, once the expression rows of ALL cells are resident on
 * this GPU (e_all_cm: cols x ld cell-major fp32 DEVICE matrix -- the all-gathered blocks; e_lo_all_cm: its residual
 * matrix or NULL).  Work already queued on e_ready_stream (the all-gather) is waited for on the device, so the
 * upload of the first velocity chunk overlaps it.  d_block: HOST gene-major, `rows` rows of nc values at a row pitch
 * of d_cols values; ixs_block / out_block: HOST nc x nrndm (GLOBAL neighbour ids / P(c -> ixs[c, n]), or the
 * correlations when sigma <= 0).  d, ixs and out move in cell chunks underneath the correlation kernel. */
int velo_transition_prob_partial_sharded(int transform, const float *e_all_cm, const float *e_lo_all_cm, int64_t ld,
                                         velo_stream_t e_ready_stream, const void *d_block, int elem_bytes,
                                         int64_t d_cols, const int64_t *ixs_block, float *out_block,
                                         int64_t rows, int64_t cols, int64_t c0, int64_t nc, int64_t nrndm,
                                         double psc, double sigma);

/* ---------------------------------------------------- (2) device tier ------ */
/* gene-major (G x C, row-major, host layout of the reference) -> cell-major fp32
 * dst[c * ld + g], c in [0,C), g in [g_off, g_off+G); pad columns are NOT touched.
 * src is a DEVICE pointer; elem_bytes 8 (fp64) or 4 (fp32). */
int velo_dev_pack_cellmajor(const void *src_gc, int elem_bytes, int64_t G, int64_t C,
                            float *dst_cg, int64_t ld, int64_t g_off, velo_stream_t stream);
/* As above for fp64 sources, additionally writing the fp32 residuals dst_lo = (float)(x - (double)(float)x)
 * (same layout) and setting *nonzero_flag (device int, caller-zeroed) to 1 when any residual is non-zero.
 * The residual matrix lets the correlation kernel reproduce the reference's fp64 sign of e[i]-e[c] for values
 * that tie in fp32.  dst_lo_cg == NULL behaves like velo_dev_pack_cellmajor. */
int velo_dev_pack_cellmajor_split(const void *src_gc, int elem_bytes, int64_t G, int64_t C,
                                  float *dst_cg, float *dst_lo_cg, int *nonzero_flag,
                                  int64_t ld, int64_t g_off, velo_stream_t stream);
/* cell-major fp32 -> gene-major fp64/fp32 (inverse of the above, for results that
 * go back to the reference's attribute layout) */
int velo_dev_unpack_genemajor(const float *src_cg, int64_t ld, int64_t G, int64_t C,
                              void *dst_gc, int elem_bytes, velo_stream_t stream);
int velo_dev_i64_to_i32(const int64_t *src, int32_t *dst, int64_t n, velo_stream_t stream);

/* per-cell mean and centred sum of squares of the velocity rows:
 * stats[2*r] = mean_g d[r, g], stats[2*r+1] = sum_g (d[r, g] - mean)^2
 * (speedboosted.pyx:46-55, 67-72).  d_cm: nc x ld. */
int velo_dev_cell_stats(const float *d_cm, int64_t ld, int64_t G, int64_t nc,
                        float *stats, velo_stream_t stream);

/* The correlation kernel (K1).  For local cells r in [0, nc) (global id c0 + r):
 *   out[r * out_ld + n] = pearson_g( f(e[i, g] - e[c0 + r, g]), d[r, g] ),
 *   i = ixs[r * ixs_ld + n]   (ixs != NULL, n < m)      -- "partial"
 *   i = n                      (ixs == NULL, m == C)     -- "full"
 * e_cm : C x ld cell-major fp32 (ALL cells: neighbours may be any cell)
 * d_cm : nc x ld (rows of the local cells only), stats from velo_dev_cell_stats
 * rule : VELO_RULE_PARTIAL or VELO_RULE_FULL (zero rule of the variant)
 * Workspace-free, one launch -- except the all-pairs linear case, which runs on the tensor cores (K2g below). */
int velo_dev_coldeltacor(int transform, int rule,
                         const float *e_cm, const float *d_cm, int64_t ld,
                         const float *stats, const int32_t *ixs, int64_t ixs_ld,
                         float *out, int64_t out_ld,
                         int64_t G, int64_t C, int64_t c0, int64_t nc, int64_t m,
                         double psc, velo_stream_t stream);

/* Same with the optional residual matrix e_lo_cm (C x ld, from velo_dev_pack_cellmajor_split; NULL = none):
 * exact fp32 ties of non-zero values are resolved with the residuals, as the fp64 reference would see them. */
int velo_dev_coldeltacor_ex(int transform, int rule,
                            const float *e_cm, const float *e_lo_cm, const float *d_cm, int64_t ld,
                            const float *stats, const int32_t *ixs, int64_t ixs_ld,
                            float *out, int64_t out_ld,
                            int64_t G, int64_t C, int64_t c0, int64_t nc, int64_t m,
                            double psc, velo_stream_t stream);

/* K2g: the all-pairs LINEAR variant (x_colDeltaCor, speedboosted.pyx:13-87; `_colDeltaCor` :542-550) on the tensor
 * cores: out[r * out_ld + i] = pearson_g(e[i, g] - e[c0 + r, g], d[r, g]) for every target i in [0, C), from the two
 * products P = B X^T and Q = X X^T over the gene axis (fp16 hi/lo split operands -- bf16 pairs are 6 bits short, DESIGN.md -- tcgen05.mma, fp32 block sums drained
 * from TMEM every 64 genes; DESIGN.md "K2g").  Self pairs and coincident cells give NaN like the reference.
 * Scratch (velo_coldeltacor_tc_workspace_bytes) comes from the stream-ordered pool.  dbgP/dbgQ: optional nc x out_ld
 * raw products for diagnostics (both NULL in production). */
int    velo_dev_coldeltacor_tc(const float *e_cm, const float *d_cm, int64_t ld, const float *stats,
                               float *out, int64_t out_ld, int64_t G, int64_t C, int64_t c0, int64_t nc,
                               float *dbgP, float *dbgQ, velo_stream_t stream);
size_t velo_coldeltacor_tc_workspace_bytes(int64_t G, int64_t C, int64_t nc);
/* velo_dev_coldeltacor{,_ex} route (ixs == NULL, VELO_LINEAR) to K2g; 0 switches back to the fp32 kernel K2
 * (A/B measurements, or when the K2g scratch of 8 bytes per matrix element does not fit).  Default 1. */
void   velo_set_tensor_cores(int enable);
int    velo_get_tensor_cores(void);

/* rm[(c0 + r) * C + i] += out[r, n]  (dense adapter for small C; fp64 atomics so that
 * duplicated indices accumulate as in the reference, speedboosted.pyx:336) */
int velo_dev_scatter_dense(const float *out, int64_t out_ld, const int32_t *ixs, int64_t ixs_ld,
                           double *rm, int64_t C, int64_t c0, int64_t nc, int64_t m,
                           velo_stream_t stream);

/* transition probabilities, compact form of analysis.py:1604-1612 + 1697-1698:
 * corr patched (self -> 0, NaN -> 1), p[r, n] = exp(corr/sigma) / sum_n exp(corr/sigma).
 * ixs == NULL means neighbour n is cell n (full mode). In place allowed (p == corr). */
int velo_dev_transition_prob(const float *corr, int64_t ld, const int32_t *ixs, int64_t ixs_ld,
                             float *p, int64_t p_ld, int64_t c0, int64_t nc, int64_t m,
                             double sigma, velo_stream_t stream);
/* As above with the NaN rule selectable: patch_nan != 0 is the knn_random branch (NaN -> 1, analysis.py:1605-1606);
 * patch_nan == 0 is the "full" branch (analysis.py:1666-1668 only zeroes the diagonal): a NaN correlation makes its
 * whole row of probabilities NaN, as exp(NaN) and the NaN row sum do in the reference.  The row maximum is
 * subtracted before exponentiating, so any sigma > 0 is safe in fp32. */
int velo_dev_transition_prob_ex(const float *corr, int64_t ld, const int32_t *ixs, int64_t ixs_ld,
                                float *p, int64_t p_ld, int64_t c0, int64_t nc, int64_t m,
                                double sigma, int patch_nan, velo_stream_t stream);

/* ---- gamma fits (K4): replaces the per-gene SciPy loop of velocyto/estimation.py:173-366 ----
 * S_cm = X (spliced, independent), U_cm = Y (unspliced, dependent), W_cm weights; all C x ld(w)
 * cell-major fp32.  cell_mask: optional C bytes (the steady_state selection, analysis.py:1159-1162).
 * mode 0 fit_slope (nnls, estimation.py:173-188,267-279)      -> gamma
 * mode 1 fit_slope_offset (leastsq/OLS, :244-264,282-297)       -> gamma, offset
 * mode 2 fit_slope_weighted (bounded (lo,hi), :191-209,300-334) -> gamma [, r2]
 * mode 3 fit_slope_weighted_offset (box m in [lo,hi], q in [0, 2*sum(yw)/sum(w)], :212-241,337-366)
 *                                                               -> gamma, offset [, r2]
 * gamma/offset/r2: G floats (offset, r2 may be NULL); moments: optional 14 x G fp64 sufficient statistics. */
int velo_dev_fit_gammas(int mode, const float *S_cm, const float *U_cm, int64_t ld,
                        const float *W_cm, int64_t ldw, const uint8_t *cell_mask,
                        int64_t G, int64_t C, double lo, double hi,
                        float *gamma, float *offset, float *r2, double *moments, velo_stream_t stream);

/* np.percentile(rows[g, :], q[j]) ("linear" interpolation) for every row of a gene-major fp32 matrix
 * (rows_gc: G x C; q_dev: nq percentiles in [0,100], DEVICE array; out: G x nq fp64, device) */
int velo_dev_row_percentiles(const float *rows_gc, int64_t G, int64_t C, const double *q_dev, int nq,
                             double *out, velo_stream_t stream);
/* weight matrix of fit_gammas (analysis.py:1179-1219), cell-major fp32 C x ldw.
 * kind 0 "maxmin_diag" (default), 1 "maxmin", 2 "maxmin_double", 3 "sum", 4 "prod";
 * S/U = the matrices being fitted (tmpS/tmpU), Sx/Ux = the smoothed matrices the diag modes use. */
int velo_dev_fit_weights(int kind, const float *S_cm, const float *U_cm, const float *Sx_cm,
                         const float *Ux_cm, int64_t ld, int64_t G, int64_t C, double perc_lo,
                         double perc_hi, float *W_cm, int64_t ldw, velo_stream_t stream);

/* As velo_dev_fit_weights plus kind 5 "maxmin_weighted" (analysis.py:1186-1192): W = (R^power + (1-R)^power) / 2 with
 * R = S clipped to its [perc_lo, perc_hi] percentiles and rescaled to [0, 1] (power = maxmin_weighted_pow). */
int velo_dev_fit_weights_ex(int kind, const float *S_cm, const float *U_cm, const float *Sx_cm,
                            const float *Ux_cm, int64_t ld, int64_t G, int64_t C, double perc_lo,
                            double perc_hi, double power, float *W_cm, int64_t ldw, velo_stream_t stream);

/* As velo_dev_fit_gammas with the non-default options: hi_per_gene (G fp64, device; limit_gamma's per-gene upper
 * slope bound, estimation.py:199-204/229-234) and q_fixed (G fp64, device; fixperc_q's pinned offset,
 * estimation.py:221-224/254-257); either may be NULL.  velo_dev_fit_constraints computes both from the data. */
int velo_dev_fit_gammas_ex(int mode, const float *S_cm, const float *U_cm, int64_t ld,
                           const float *W_cm, int64_t ldw, const uint8_t *cell_mask,
                           int64_t G, int64_t C, double lo, double hi,
                           const double *hi_per_gene, const double *q_fixed,
                           float *gamma, float *offset, float *r2, double *moments, velo_stream_t stream);
/* q_fix[g] = median(U[g, S[g,:] <= percentile(S[g,:], 1)]);  up_gamma[g] = median(U) > median(S) ?
 * max(1.5, percentile(U[S > p90(S)], 10) / median(S[S > p90(S)])) : 1.5.  Outputs: G fp64 (device); either may be NULL. */
int velo_dev_fit_constraints(const float *S_cm, const float *U_cm, int64_t ld, int64_t G, int64_t C,
                             double *q_fix, double *up_gamma, velo_stream_t stream);

/* ---- elementwise chain (K6): predict_U -> calculate_velocity -> calculate_shift ->
 * extrapolate_cell_at_t -> velocity transform (analysis.py:1343-1346,1369,1398-1406,1428-1431,1577/1597).
 * gamma, q (q may be NULL), vel_thr (NULL = no eps threshold): G floats.  assumption 0 = constant_velocity,
 * 1 = constant_unspliced.  Any of the five C x ld outputs may be NULL. */
int velo_dev_velocity_chain(const float *S_cm, const float *U_cm, int64_t ld, const float *gamma,
                            const float *q, const float *vel_thr, int64_t G, int64_t C, int assumption,
                            double dt_shift, double dt_extrap, int clip, int transform, double psc,
                            float *Upred, float *vel, float *delta_S, float *S_t, float *d_transformed,
                            velo_stream_t stream);
/* thr[g] = eps * max_c (gamma[g]*S[c,g] + q[g])   (analysis.py:1377-1378) */
int velo_dev_velocity_threshold(const float *S_cm, int64_t ld, const float *gamma, const float *q,
                                int64_t G, int64_t C, double eps, float *thr, velo_stream_t stream);

/* out = f(dt * delta_S): the `d` argument of the correlation kernel (analysis.py:1577/1594/1597) */
int velo_dev_delta_transform(const float *delta_S_cm, float *out_cm, int64_t ld, int64_t C, double dt,
                             int transform, double psc, velo_stream_t stream);
/* out = S + dt * delta_S, clipped at 0 when clip != 0 (extrapolate_cell_at_t, analysis.py:1429-1431) */
int velo_dev_extrapolate(const float *S_cm, const float *delta_S_cm, float *out_cm, int64_t ld, int64_t C,
                         double dt, int clip, velo_stream_t stream);
/* transform="logratio" operands (analysis.py:1582-1583): which 0: out = log2(S + psc);
 * which 1: out = log2(|S + dt*delta_S| + psc) - log2(S + psc) */
int velo_dev_logratio(const float *S_cm, const float *delta_S_cm, float *out_cm, int64_t ld, int64_t C,
                      double dt, double psc, int which, velo_stream_t stream);
/* expression scaling of calculate_embedding_shift (analysis.py:1714-1719): scale[c] =
 * clip( sum_g delta_S[c,g]*estim[c,g] / sqrt(sum_g estim[c,g]^2) / penalty, 0, 1 ), fp64 (device, C values) */
int velo_dev_row_cosine_scale(const float *delta_S_cm, const float *estim_cm, int64_t ld, int64_t G, int64_t C,
                              double penalty, double *scale, velo_stream_t stream);
/* in place: self pair -> 0, and NaN -> 1 when patch_nan != 0 (analysis.py:1604-1612); *nan_count
 * (device, may be NULL) is incremented by the number of NaNs replaced */
int velo_dev_patch_corr(float *corr, int64_t ld, const int32_t *ixs, int64_t ixs_ld, int64_t c0, int64_t nc,
                        int64_t m, int patch_nan, unsigned long long *nan_count, velo_stream_t stream);
/* delta_embedding[r,:] = sum_n (P[r,n] - 1/m) * unit(emb[ixs[r,n]] - emb[c0+r])  (analysis.py:1704-1712);
 * embedding: C x dims fp64 (first two coordinates used), out: nc x 2 fp64 */
int velo_dev_embedding_shift(const float *P, int64_t ld, const int32_t *ixs, int64_t ixs_ld,
                             const double *embedding, int dims, int64_t c0, int64_t nc, int64_t m,
                             double *out, velo_stream_t stream);

/* ---- size / log normalisation (the `normalize` family, analysis.py:535-676) ----
 * sums[c] = sum_g X[c, g] (= X.sum(0) of the reference's gene-major matrix), fp64 (device, C values) */
int velo_dev_cell_sums(const float *X_cm, int64_t ld, int64_t G, int64_t C, double *sums, velo_stream_t stream);
/* out_sz = factor[c] * X (factor: C fp64 on the device, NULL = 1; non-finite -> 0 when nonfinite_to_zero != 0, the guard
 * of _normalize_U/_normalize_Ux analysis.py:581,630); out_norm = log2(out_sz + pcount).  Either output may be NULL. */
int velo_dev_size_normalize(const float *X_cm, int64_t ld, int64_t G, int64_t C, const double *factor,
                            double pcount, int nonfinite_to_zero, float *out_sz, float *out_norm,
                            velo_stream_t stream);

/* ---- kNN smoothing (K5): out[c,:] = sum_p weights[p] * S[indices[p],:], p in [indptr[c], indptr[c+1])
 * = convolve_by_sparse_weights(data, w) with w in CSR by rows (neighbors.py:416-423, weights from
 * :385-390); maximum != 0 applies np.maximum(S, Sx) (analysis.py:1017-1019).  fp64 accumulation. */
int velo_dev_knn_smooth(const int64_t *indptr, const int32_t *indices, const float *weights,
                        const float *S_cm, float *out_cm, int64_t ld, int64_t G, int64_t C,
                        int maximum, velo_stream_t stream);

/* kNN smoothing of SPARSE counts (BASELINE config 5): S is CSR by cell (s_indptr C+1, s_genes sorted gene ids,
 * s_values), the smoothing weights CSR by cell as above; writes the dense gene slab [g0, g0+ng) of the smoothed
 * matrix: out_cm[c * ld_out + (g - g0)].  Deterministic (64-bit fixed-point shared-memory accumulation). */
int velo_dev_knn_smooth_csr(const int64_t *w_indptr, const int32_t *w_indices, const float *w_weights,
                            const int64_t *s_indptr, const int32_t *s_genes, const float *s_values,
                            float *out_cm, int64_t ld_out, int64_t C, int64_t g0, int64_t ng, int maximum,
                            velo_stream_t stream);

/* Sparse ingest: counts arrive as CSR by cell -- what 10x / AnnData / sparse-loom HDF5 files store
 * (indptr over cells, gene ids, values) -- and never exist as a dense float64 host matrix (the reference's loader,
 * analysis.py:56-64, reads `ds.layer[...][:, :]` densely: 120 GB per matrix at 500k x 30k).
 * velo_dev_csr_to_cellmajor: dense cell-major fp32 slab [g0, g0 + ng) of the matrix: out_cm[c * ld + (g - g0)]
 * (pad columns zeroed).  Gene ids sorted within a cell.
 * velo_dev_csr_cell_sums_scale: optional in-place values[q] *= factor[cell] (non-finite -> 0), then optional per-cell
 * totals (fp64) of the (rescaled) values: the size normalisation of analysis.py:535-584 on the sparse form. */
int velo_dev_csr_to_cellmajor(const int64_t *indptr, const int32_t *genes, const float *values, int64_t C,
                              int64_t g0, int64_t ng, float *out_cm, int64_t ld, velo_stream_t stream);
int velo_dev_csr_cell_sums_scale(const int64_t *indptr, float *values, int64_t C, const double *factor,
                                 double *sums, velo_stream_t stream);

/* Exact brute-force kNN (Euclidean) in a low-dimensional space: X is C x D fp64 row-major (device); writes the k
 * nearest points of every point in ascending distance: out_idx (C x k int32), out_dist (C x k fp64 or NULL).
 * include_self = 0 excludes the query point itself (scikit-learn's kneighbors with X=None, neighbors.py:370-376,
 * analysis.py:1549); 1 keeps it (normally at rank 0; BalancedKNN's candidate lists, neighbors.py:282). k <= ~14000. */
int velo_dev_knn(const double *X, int64_t C, int D, int k, int include_self,
                 int32_t *out_idx, double *out_dist, velo_stream_t stream);
/* The same search for the query points [q0, q0 + nq) only (neighbours are still drawn from all C points): out_idx /
 * out_dist are nq x k.  Queries are independent, so a multi-GPU caller gives every rank a block of queries and
 * all-gathers the index blocks ("kNN search").  More than k/8+32 points tying at the k-th distance
 * (duplicated points) are resolved by lowest index. */
int velo_dev_knn_range(const double *X, int64_t C, int D, int k, int include_self, int64_t q0, int64_t nq,
                       int32_t *out_idx, double *out_dist, velo_stream_t stream);

/* Neighbours of nq separate QUERY points Q (nq x D fp64, device) among the C points of X -- the grid-point search of
 * calculate_grid_arrows (analysis.py:1788-1790: NearestNeighbors.fit(embedding).kneighbors(gridpoints)). */
int velo_dev_knn_query(const double *X, int64_t C, int D, const double *Q, int64_t nq, int k,
                       int32_t *out_idx, double *out_dist, velo_stream_t stream);
/* calculate_grid_arrows core (analysis.py:1792-1797): gaussian kernel weights w = normal.pdf(dists, scale = sigma) of
 * every grid point's k nearest cells, mass[p] = sum_n w, flow[p, :] = sum_n w * delta[neighs[p, n], :] / max(1, mass[p]).
 * neighs / dists: npts x k (from velo_dev_knn_query); delta: cells x dims fp64 (delta_embedding); outputs fp64. */
int velo_dev_grid_flow(const int32_t *neighs, const double *dists, int64_t npts, int k, const double *delta,
                       int dims, double sigma, double *mass, double *flow, velo_stream_t stream);

/* ---- device-side randomisation (opt-in; the default keeps the reference's NumPy / numba streams on the host) ----
 * Weighted sampling without replacement of m of the W candidate neighbours of every cell (analysis.py:1552-1566, one
 * np.random.choice(W, m, replace=False, p) per cell): same distribution (successive sampling, order included) through
 * exponential-clock keys -log(u) / p and a per-cell sort; Philox4x32-10 stream keyed by (seed, cell, candidate).
 * knn_idx: C x W int32 candidate cells (row = kNN order); inv_p: W floats 1 / p_j (p_j > 0);
 * outputs C x m int32: sampling_ixs = chosen positions, neigh_ixs = knn_idx[c, sampling_ixs[c, :]]. */
int velo_dev_sample_neighbors(const int32_t *knn_idx, int64_t C, int W, const float *inv_p, int m, uint64_t seed,
                              int32_t *neigh_ixs, int32_t *sampling_ixs, velo_stream_t stream);
/* Randomised control permute_rows_nsign (analysis.py:2413-2420) in the cell-major layout: out[c, g] = +-in[pi_g(c), g],
 * an independent pseudo-random permutation of the cells and independent signs for every gene (Feistel bijection, no
 * sort).  in != out. */
int velo_dev_permute_rows_nsign(const float *in_cm, float *out_cm, int64_t ld, int64_t G, int64_t C, uint64_t seed,
                                velo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VELO_B200_H */
