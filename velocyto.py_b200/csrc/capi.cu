// Library plumbing + the HOST drop-in tier of the C ABI (include/velo_b200.h).
//
// The host entry points take exactly what the reference's native functions take
// (x_colDeltaCor*, velocyto/speedboosted.pyx:13-538): host pointers to gene-major fp64
// matrices, an index matrix, and a dense cells x cells fp64 output that is accumulated into.
// They stage the inputs to HBM in gene-row chunks, convert them on the device to the
// cell-major fp32 layout of the kernels, run the same device tier the Python layer uses,
// and bring the result back.  No host arithmetic on the data path.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "velo_common.cuh"

namespace velo {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

int get_device_props(DeviceProps *out)
{
    static std::mutex mu;
    static DeviceProps cache[64];
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device visible: libvelo_b200 has no CPU fallback");
        return VELO_E_NODEVICE;
    }
    int dev = 0;
    VELO_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && cache[dev].device == dev) {
        *out = cache[dev];
        return VELO_OK;
    }
    cudaDeviceProp prop;
    VELO_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    DeviceProps dp;
    dp.device = dev;
    dp.sm_count = prop.multiProcessorCount;
    dp.smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    dp.cc_major = prop.major;
    dp.cc_minor = prop.minor;
    dp.hbm_bytes = prop.totalGlobalMem;
    if (dp.cc_major != 10) {
        set_error("device %d is sm_%d%d; libvelo_b200 is built for sm_100a only", dev, prop.major, prop.minor);
        return VELO_E_NODEVICE;
    }
    // Workspace comes from the device's stream-ordered pool (cudaMallocAsync).  By default the pool hands memory
    // back to the driver at every synchronisation, which turns each call's scratch (tens of MB for the fits, tens
    // of GB for the host tier) into fresh driver allocations; keep it cached instead (velo_release_workspace trims).
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    (void)cudaGetLastError();
    if (dev < 64) cache[dev] = dp;
    *out = dp;
    return VELO_OK;
}

// RAII for stream-ordered device allocations inside the host tier
struct DevBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf()
    {
        if (p) cudaFreeAsync(p, st);
    }
    int alloc(size_t bytes, cudaStream_t s)
    {
        st = s;
        VELO_CUDA_TRY(cudaMallocAsync(&p, bytes ? bytes : 16, s));
        return VELO_OK;
    }
    template <typename T>
    T *as() const
    {
        return static_cast<T *>(p);
    }
};

// upload a host gene-major matrix into a cell-major fp32 device matrix, gene-row chunk by chunk
static int upload_cellmajor(const void *host_gc, int elem_bytes, int64_t G, int64_t C, float *dst_cg, int64_t ld,
                            cudaStream_t st, float *dst_lo = nullptr, int *nz_flag = nullptr)
{
    const int64_t row_bytes = C * elem_bytes;
    int64_t chunk_rows = (256LL << 20) / (row_bytes > 0 ? row_bytes : 1);
    if (chunk_rows < 32) chunk_rows = 32;
    if (chunk_rows > G) chunk_rows = G;
    DevBuf stage;
    int rc = stage.alloc(static_cast<size_t>(chunk_rows * row_bytes), st);
    if (rc) return rc;
    for (int64_t g0 = 0; g0 < G; g0 += chunk_rows) {
        const int64_t rows = (G - g0 < chunk_rows) ? G - g0 : chunk_rows;
        VELO_CUDA_TRY(cudaMemcpyAsync(stage.p, static_cast<const char *>(host_gc) + g0 * row_bytes,
                                      static_cast<size_t>(rows * row_bytes), cudaMemcpyHostToDevice, st));
        rc = velo_dev_pack_cellmajor_split(stage.p, elem_bytes, rows, C, dst_cg, dst_lo, nz_flag, ld, g0, st);
        if (rc) return rc;
    }
    return VELO_OK;
}

// Compact-output partial path, pipelined: the whole expression matrix e has to be resident before the first cell can be
// processed (neighbours are arbitrary cells), but a cell's velocity row d_c, its neighbour list and its output row are
// only needed while that cell is being processed.  So e goes up first, and d / ixs / out move in cell chunks on their own
// streams underneath the correlation kernel of the previous chunk (at 100k x 30k: 24 GB of the 50 GB of H2D traffic and
// all of the D2H traffic leave the critical path).
static int host_partial_compact_pipelined(int transform, int rule, const void *e, const void *d, int elem_bytes,
                                          const int64_t *ixs, int64_t G, int64_t C, int64_t m, double psc,
                                          float *out_compact, double sigma, const DeviceProps &dp)
{
    const int64_t ld = round_up(G, 32);
    const int64_t eb = elem_bytes;
    // chunking: a multiple of the SM count (K1 strides cells statically over one persistent CTA per SM)
    int64_t nchunks = (C * G * eb >= (2LL << 30)) ? 8 : 1;
    int64_t chunk = round_up((C + nchunks - 1) / nchunks, dp.sm_count > 0 ? dp.sm_count : 148);
    if (const char *env = getenv("VELO_HOST_CHUNK_CELLS")) {      // test hook: force the multi-chunk pipeline on small inputs
        const long long v = atoll(env);
        if (v > 0) chunk = v;
    }
    if (chunk > C) chunk = C;
    nchunks = (C + chunk - 1) / chunk;

    struct Streams {                 // declared FIRST: destroyed after every buffer has been handed back (cudaFreeAsync)
        cudaStream_t st = nullptr, cp = nullptr, dd = nullptr;
        std::vector<cudaEvent_t> ev;
        void sync()
        {
            if (st) cudaStreamSynchronize(st);
            if (cp) cudaStreamSynchronize(cp);
            if (dd) cudaStreamSynchronize(dd);
        }
        ~Streams()
        {
            sync();
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
            if (st) cudaStreamDestroy(st);
            if (cp) cudaStreamDestroy(cp);
            if (dd) cudaStreamDestroy(dd);
        }
    };
    struct PinnedInts {
        int *p = nullptr;
        ~PinnedInts()
        {
            if (p) cudaFreeHost(p);
        }
    };
    struct Quiesce {                 // declared LAST: on every exit path nothing is still using the buffers when they go
        Streams &s;
        ~Quiesce() { s.sync(); }
    };
    Streams S;
    DevBuf e_cm, d_cm, stats, out, ix32, e_lo, lo_flag, stage_d, stage_ix, flags;
    PinnedInts bad;
    Quiesce quiesce{S};
    VELO_CUDA_TRY(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
    VELO_CUDA_TRY(cudaStreamCreateWithFlags(&S.cp, cudaStreamNonBlocking));
    VELO_CUDA_TRY(cudaStreamCreateWithFlags(&S.dd, cudaStreamNonBlocking));
    cudaStream_t st = S.st, cp = S.cp, dd = S.dd;
    auto new_event = [&](cudaEvent_t *ev) -> int {
        VELO_CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        S.ev.push_back(*ev);
        return VELO_OK;
    };
    int rc;
    if ((rc = e_cm.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
    if ((rc = d_cm.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
    if ((rc = stats.alloc(static_cast<size_t>(C) * 2 * 4, st))) return rc;
    if ((rc = out.alloc(static_cast<size_t>(C * m) * 4, st))) return rc;
    if ((rc = ix32.alloc(static_cast<size_t>(C * m) * 4, st))) return rc;
    if ((rc = stage_d.alloc(static_cast<size_t>(chunk * G * eb), st))) return rc;
    if ((rc = stage_ix.alloc(static_cast<size_t>(chunk * m) * 8, st))) return rc;
    if ((rc = flags.alloc(static_cast<size_t>(nchunks) * sizeof(int), st))) return rc;
    VELO_CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&bad.p), static_cast<size_t>(nchunks) * sizeof(int)));
    VELO_CUDA_TRY(cudaMemsetAsync(flags.p, 0, static_cast<size_t>(nchunks) * sizeof(int), st));
    if (ld != G) {
        VELO_CUDA_TRY(cudaMemsetAsync(e_cm.p, 0, static_cast<size_t>(C * ld) * 4, st));
        VELO_CUDA_TRY(cudaMemsetAsync(d_cm.p, 0, static_cast<size_t>(C * ld) * 4, st));
    }
    // fp64 inputs + a transform that jumps at zero difference: keep the fp32 residuals of e
    const double jump = transform == VELO_SQRT ? 2.0 * sqrt(psc > 0 ? psc : 0.0)
                        : transform == VELO_LOG10 ? 2.0 * fabs(log10(psc > 0 ? psc : 1e-300)) : 0.0;
    const bool want_lo = elem_bytes == 8 && jump > 1e-4;
    int lo_nonzero = 0;
    if (want_lo) {
        if ((rc = e_lo.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
        if ((rc = lo_flag.alloc(sizeof(int), st))) return rc;
        VELO_CUDA_TRY(cudaMemsetAsync(e_lo.p, 0, static_cast<size_t>(C * ld) * 4, st));
        VELO_CUDA_TRY(cudaMemsetAsync(lo_flag.p, 0, sizeof(int), st));
    }
    cudaEvent_t ev_alloc;
    if ((rc = new_event(&ev_alloc))) return rc;
    VELO_CUDA_TRY(cudaEventRecord(ev_alloc, st));
    VELO_CUDA_TRY(cudaStreamWaitEvent(cp, ev_alloc, 0));          // the copy stream may touch the buffers from here on

    // e: all of it, on the copy stream (the first d chunk queues up right behind it)
    if ((rc = upload_cellmajor(e, elem_bytes, G, C, e_cm.as<float>(), ld, cp, want_lo ? e_lo.as<float>() : nullptr,
                               want_lo ? lo_flag.as<int>() : nullptr)))
        return rc;
    cudaEvent_t ev_e;
    if ((rc = new_event(&ev_e))) return rc;
    if (want_lo) VELO_CUDA_TRY(cudaMemcpyAsync(&lo_nonzero, lo_flag.p, sizeof(int), cudaMemcpyDeviceToHost, cp));
    VELO_CUDA_TRY(cudaEventRecord(ev_e, cp));

    std::vector<cudaEvent_t> ev_ready(nchunks), ev_out(nchunks);
    auto enqueue_copy = [&](int64_t j) -> int {                   // H2D + layout conversion of chunk j, copy stream
        const int64_t c0 = j * chunk, nc = (C - c0 < chunk) ? C - c0 : chunk;
        VELO_CUDA_TRY(cudaMemcpy2DAsync(stage_d.p, static_cast<size_t>(nc * eb), static_cast<const char *>(d) + c0 * eb,
                                        static_cast<size_t>(C * eb), static_cast<size_t>(nc * eb), static_cast<size_t>(G),
                                        cudaMemcpyHostToDevice, cp));
        int r = velo_dev_pack_cellmajor(stage_d.p, elem_bytes, G, nc, d_cm.as<float>() + c0 * ld, ld, 0, cp);
        if (r) return r;
        VELO_CUDA_TRY(cudaMemcpyAsync(stage_ix.p, ixs + c0 * m, static_cast<size_t>(nc * m) * 8, cudaMemcpyHostToDevice, cp));
        if ((r = i64_to_i32_checked(stage_ix.as<int64_t>(), ix32.as<int32_t>() + c0 * m, nc * m, C, flags.as<int>() + j, cp)))
            return r;
        VELO_CUDA_TRY(cudaMemcpyAsync(bad.p + j, flags.as<int>() + j, sizeof(int), cudaMemcpyDeviceToHost, cp));
        if ((r = new_event(&ev_ready[j]))) return r;
        VELO_CUDA_TRY(cudaEventRecord(ev_ready[j], cp));
        return VELO_OK;
    };
    if ((rc = enqueue_copy(0))) return rc;
    VELO_CUDA_TRY(cudaEventSynchronize(ev_e));                    // lo_nonzero is known; e is resident
    VELO_CUDA_TRY(cudaStreamWaitEvent(st, ev_e, 0));
    for (int64_t j = 0; j < nchunks; ++j) {
        const int64_t c0 = j * chunk, nc = (C - c0 < chunk) ? C - c0 : chunk;
        if (j + 1 < nchunks && (rc = enqueue_copy(j + 1))) return rc;      // flies under this chunk's kernel
        VELO_CUDA_TRY(cudaEventSynchronize(ev_ready[j]));
        // the kernel must not gather through an out-of-range index: check before launching it
        VELO_REQUIRE(bad.p[j] == 0, "colDeltaCor: ixs holds an index outside [0, %lld)", static_cast<long long>(C));
        VELO_CUDA_TRY(cudaStreamWaitEvent(st, ev_ready[j], 0));
        if ((rc = velo_dev_cell_stats(d_cm.as<float>() + c0 * ld, ld, G, nc, stats.as<float>() + 2 * c0, st))) return rc;
        if ((rc = velo_dev_coldeltacor_ex(transform, rule, e_cm.as<float>(), lo_nonzero ? e_lo.as<float>() : nullptr,
                                          d_cm.as<float>() + c0 * ld, ld, stats.as<float>() + 2 * c0,
                                          ix32.as<int32_t>() + c0 * m, m, out.as<float>() + c0 * m, m, G, C, c0, nc, m, psc,
                                          st)))
            return rc;
        if (sigma > 0.0 && (rc = velo_dev_transition_prob(out.as<float>() + c0 * m, m, ix32.as<int32_t>() + c0 * m, m,
                                                          out.as<float>() + c0 * m, m, c0, nc, m, sigma, st)))
            return rc;
        if ((rc = new_event(&ev_out[j]))) return rc;
        VELO_CUDA_TRY(cudaEventRecord(ev_out[j], st));
        VELO_CUDA_TRY(cudaStreamWaitEvent(dd, ev_out[j], 0));
        VELO_CUDA_TRY(cudaMemcpyAsync(out_compact + c0 * m, out.as<float>() + c0 * m, static_cast<size_t>(nc * m) * 4,
                                      cudaMemcpyDeviceToHost, dd));
    }
    VELO_CUDA_TRY(cudaStreamSynchronize(st));
    VELO_CUDA_TRY(cudaStreamSynchronize(dd));
    VELO_CUDA_TRY(cudaStreamSynchronize(cp));
    return VELO_OK;
}

// Shared body of the host tier.  Exactly one of (rm, out_compact) is non-null.
static int host_coldeltacor(int transform, int rule, const void *e, const void *d, int elem_bytes,
                            const int64_t *ixs, int64_t rows, int64_t cols, int64_t nrndm, double psc, double *rm,
                            float *out_compact, double sigma = 0.0)
{
    VELO_REQUIRE(e && d && (rm || out_compact), "colDeltaCor: null pointer");
    VELO_REQUIRE(rows > 0 && cols > 0, "colDeltaCor: empty matrix (%lld x %lld)", static_cast<long long>(rows),
                 static_cast<long long>(cols));
    VELO_REQUIRE(ixs == nullptr || nrndm >= 0, "colDeltaCor: negative neighbour count");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    const int64_t G = rows, C = cols;
    const int64_t m = ixs ? nrndm : C;
    if (m == 0) return VELO_OK;
    if (out_compact && ixs)
        return host_partial_compact_pipelined(transform, rule, e, d, elem_bytes, ixs, G, C, m, psc, out_compact, sigma, dp);
    const int64_t ld = round_up(G, 32);

    cudaStream_t st;
    VELO_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } guard{st};

    DevBuf e_cm, d_cm, stats, out, ix32, ix64, rm_dev, flag, e_lo, lo_flag;
    if ((rc = e_cm.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
    if ((rc = d_cm.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
    if ((rc = stats.alloc(static_cast<size_t>(C) * 2 * 4, st))) return rc;
    if ((rc = out.alloc(static_cast<size_t>(C * m) * 4, st))) return rc;
    if (ld != G) {   // pad columns: defined values (never enter the sums)
        VELO_CUDA_TRY(cudaMemsetAsync(e_cm.p, 0, static_cast<size_t>(C * ld) * 4, st));
        VELO_CUDA_TRY(cudaMemsetAsync(d_cm.p, 0, static_cast<size_t>(C * ld) * 4, st));
    }
    // fp64 inputs + a transform that jumps at zero difference (sqrt with psc > 0, log10 with psc != 1): keep the
    // fp32 residuals of e so that fp32 ties resolve to the sign the fp64 reference sees
    const double jump = transform == VELO_SQRT ? 2.0 * sqrt(psc > 0 ? psc : 0.0)
                        : transform == VELO_LOG10 ? 2.0 * fabs(log10(psc > 0 ? psc : 1e-300)) : 0.0;
    const bool want_lo = elem_bytes == 8 && jump > 1e-4;
    int lo_nonzero = 0;
    if (want_lo) {
        if ((rc = e_lo.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
        if ((rc = lo_flag.alloc(sizeof(int), st))) return rc;
        VELO_CUDA_TRY(cudaMemsetAsync(e_lo.p, 0, static_cast<size_t>(C * ld) * 4, st));
        VELO_CUDA_TRY(cudaMemsetAsync(lo_flag.p, 0, sizeof(int), st));
    }
    if ((rc = upload_cellmajor(e, elem_bytes, G, C, e_cm.as<float>(), ld, st, want_lo ? e_lo.as<float>() : nullptr,
                               want_lo ? lo_flag.as<int>() : nullptr)))
        return rc;
    if ((rc = upload_cellmajor(d, elem_bytes, G, C, d_cm.as<float>(), ld, st))) return rc;
    if (want_lo) {
        VELO_CUDA_TRY(cudaMemcpyAsync(&lo_nonzero, lo_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        VELO_CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (ixs) {
        if ((rc = ix64.alloc(static_cast<size_t>(C * m) * 8, st))) return rc;
        if ((rc = ix32.alloc(static_cast<size_t>(C * m) * 4, st))) return rc;
        if ((rc = flag.alloc(sizeof(int), st))) return rc;
        VELO_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        VELO_CUDA_TRY(cudaMemcpyAsync(ix64.p, ixs, static_cast<size_t>(C * m) * 8, cudaMemcpyHostToDevice, st));
        if ((rc = i64_to_i32_checked(ix64.as<int64_t>(), ix32.as<int32_t>(), C * m, C, flag.as<int>(), st))) return rc;
        int bad = 0;   // the kernel must not gather through an out-of-range index: check before launching it
        VELO_CUDA_TRY(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        VELO_CUDA_TRY(cudaStreamSynchronize(st));
        VELO_REQUIRE(bad == 0, "colDeltaCor: ixs holds an index outside [0, %lld)", static_cast<long long>(C));
    }
    if ((rc = velo_dev_cell_stats(d_cm.as<float>(), ld, G, C, stats.as<float>(), st))) return rc;
    if ((rc = velo_dev_coldeltacor_ex(transform, rule, e_cm.as<float>(), lo_nonzero ? e_lo.as<float>() : nullptr,
                                      d_cm.as<float>(), ld, stats.as<float>(), ixs ? ix32.as<int32_t>() : nullptr, m,
                                      out.as<float>(), m, G, C, 0, C, m, psc, st)))
        return rc;
    if (out_compact) {
        if (sigma > 0.0 &&
            (rc = velo_dev_transition_prob(out.as<float>(), m, ixs ? ix32.as<int32_t>() : nullptr, m, out.as<float>(),
                                           m, 0, C, m, sigma, st)))
            return rc;
        VELO_CUDA_TRY(cudaMemcpyAsync(out_compact, out.p, static_cast<size_t>(C * m) * 4, cudaMemcpyDeviceToHost, st));
    } else {
        // dense adapter: accumulate into the caller's rm (+= semantics, speedboosted.pyx:78,336)
        if ((rc = rm_dev.alloc(static_cast<size_t>(C * C) * 8, st))) return rc;
        VELO_CUDA_TRY(cudaMemcpyAsync(rm_dev.p, rm, static_cast<size_t>(C * C) * 8, cudaMemcpyHostToDevice, st));
        if ((rc = velo_dev_scatter_dense(out.as<float>(), m, ixs ? ix32.as<int32_t>() : nullptr, m,
                                         rm_dev.as<double>(), C, 0, C, m, st)))
            return rc;
        VELO_CUDA_TRY(cudaMemcpyAsync(rm, rm_dev.p, static_cast<size_t>(C * C) * 8, cudaMemcpyDeviceToHost, st));
    }
    VELO_CUDA_TRY(cudaStreamSynchronize(st));
    return VELO_OK;
}

}  // namespace velo

using namespace velo;

extern "C" int velo_abi_version(void) { return VELO_ABI_VERSION; }

extern "C" int velo_release_workspace(void)
{
    int dev = 0;
    VELO_CUDA_TRY(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    VELO_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    VELO_CUDA_TRY(cudaDeviceSynchronize());
    VELO_CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
    return VELO_OK;
}
extern "C" const char *velo_last_error(void) { return t_error.c_str(); }
extern "C" uint64_t velo_launch_count(void) { return g_launches.load(); }

extern "C" int velo_device_info(int *sm_count, int *smem_optin_bytes, size_t *hbm_bytes, int *cc_major, int *cc_minor)
{
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    if (sm_count) *sm_count = dp.sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = dp.smem_optin;
    if (hbm_bytes) *hbm_bytes = dp.hbm_bytes;
    if (cc_major) *cc_major = dp.cc_major;
    if (cc_minor) *cc_minor = dp.cc_minor;
    return VELO_OK;
}

extern "C" int velo_colDeltaCor(const double *e, const double *d, double *rm, int64_t rows, int64_t cols, int)
{
    return host_coldeltacor(VELO_LINEAR, VELO_RULE_FULL, e, d, 8, nullptr, rows, cols, 0, 0.0, rm, nullptr);
}
extern "C" int velo_colDeltaCorSqrt(const double *e, const double *d, double *rm, int64_t rows, int64_t cols, int,
                                    double psc)
{
    return host_coldeltacor(VELO_SQRT, VELO_RULE_FULL, e, d, 8, nullptr, rows, cols, 0, psc, rm, nullptr);
}
extern "C" int velo_colDeltaCorLog10(const double *e, const double *d, double *rm, int64_t rows, int64_t cols, int,
                                     double psc)
{
    return host_coldeltacor(VELO_LOG10, VELO_RULE_FULL, e, d, 8, nullptr, rows, cols, 0, psc, rm, nullptr);
}
extern "C" int velo_colDeltaCorpartial(const double *e, const double *d, double *rm, const int64_t *ixs, int64_t rows,
                                       int64_t cols, int64_t nrndm, int)
{
    VELO_REQUIRE(ixs, "colDeltaCorpartial: ixs is NULL");
    return host_coldeltacor(VELO_LINEAR, VELO_RULE_PARTIAL, e, d, 8, ixs, rows, cols, nrndm, 0.0, rm, nullptr);
}
extern "C" int velo_colDeltaCorSqrtpartial(const double *e, const double *d, double *rm, const int64_t *ixs,
                                           int64_t rows, int64_t cols, int64_t nrndm, int, double psc)
{
    VELO_REQUIRE(ixs, "colDeltaCorSqrtpartial: ixs is NULL");
    return host_coldeltacor(VELO_SQRT, VELO_RULE_PARTIAL, e, d, 8, ixs, rows, cols, nrndm, psc, rm, nullptr);
}
extern "C" int velo_colDeltaCorLog10partial(const double *e, const double *d, double *rm, const int64_t *ixs,
                                            int64_t rows, int64_t cols, int64_t nrndm, int, double psc)
{
    VELO_REQUIRE(ixs, "colDeltaCorLog10partial: ixs is NULL");
    return host_coldeltacor(VELO_LOG10, VELO_RULE_PARTIAL, e, d, 8, ixs, rows, cols, nrndm, psc, rm, nullptr);
}
extern "C" int velo_colDeltaCorpartial_compact(int transform, const void *e, const void *d, int elem_bytes,
                                               const int64_t *ixs, float *out, int64_t rows, int64_t cols,
                                               int64_t nrndm, double psc)
{
    VELO_REQUIRE(ixs && out, "colDeltaCorpartial_compact: null pointer");
    VELO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "colDeltaCorpartial_compact: elem_bytes must be 4 or 8");
    VELO_REQUIRE(transform >= VELO_LINEAR && transform <= VELO_LOG10, "colDeltaCorpartial_compact: unknown transform");
    return host_coldeltacor(transform, VELO_RULE_PARTIAL, e, d, elem_bytes, ixs, rows, cols, nrndm, psc, nullptr, out);
}
extern "C" int velo_transition_prob_partial(int transform, const void *e, const void *d, int elem_bytes,
                                            const int64_t *ixs, float *out, int64_t rows, int64_t cols,
                                            int64_t nrndm, double psc, double sigma)
{
    VELO_REQUIRE(ixs && out, "transition_prob_partial: null pointer");
    VELO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "transition_prob_partial: elem_bytes must be 4 or 8");
    VELO_REQUIRE(transform >= VELO_LINEAR && transform <= VELO_LOG10, "transition_prob_partial: unknown transform");
    VELO_REQUIRE(sigma > 0.0, "transition_prob_partial: sigma must be positive");
    return host_coldeltacor(transform, VELO_RULE_PARTIAL, e, d, elem_bytes, ixs, rows, cols, nrndm, psc, nullptr, out,
                            sigma);
}
