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

#include <string.h>

#include <mutex>
#include <thread>
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
    // of GB for the host tier) into fresh driver allocations; keep up to a QUARTER of the HBM cached instead (45 GB:
    // the whole working set of the host tier at 100k x 30k).  Anything above that goes back to the driver at the
    // next synchronisation, so a process that also allocates through another pool (torch's caching allocator behind
    // the CellMajor tensors) keeps 3/4 of the device; velo_release_workspace() trims the rest.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = static_cast<uint64_t>(prop.totalGlobalMem / 4);
        if (const char *env = getenv("VELO_WORKSPACE_KEEP_BYTES")) keep = strtoull(env, nullptr, 10);
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

// ------------------------------------------------------------------------------------------------------------
// Host <-> device staging.  A NumPy caller of the reference API hands over PAGEABLE memory: cudaMemcpyAsync from
// pageable memory is staged by the driver through one internal bounce buffer on the calling thread (a few GB/s,
// and synchronous) -- the chunk pipeline below would lose its overlap and PCIe would run at a fraction of its rate.
// The Stager does that staging itself: a process-wide ring of pinned bounce buffers, filled by several host
// threads (memcpy from pageable memory is bound by one core's copy rate, ~10 GB/s; PCIe 5 x16 wants ~55 GB/s)
// while the previous ring slot is on its way to the device.  Pinned / registered host memory skips it.
// ------------------------------------------------------------------------------------------------------------
static bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int host_threads()
{
    static int n = [] {
        if (const char *env = getenv("VELO_HOST_THREADS")) {
            const int v = atoi(env);
            if (v > 0) return v > 64 ? 64 : v;
        }
        const unsigned hw = std::thread::hardware_concurrency();
        int v = hw ? static_cast<int>(hw / 2) : 4;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    return n;
}

// rows x row_bytes, strided on one or both sides, copied by `host_threads()` threads (whole rows per thread; a single
// long row is cut into byte ranges instead)
static void parallel_copy_2d(char *dst, size_t dst_pitch, const char *src, size_t src_pitch, size_t row_bytes, size_t rows)
{
    const size_t total = row_bytes * rows;
    int nt = host_threads();
    if (total < (4u << 20)) nt = 1;
    auto work = [=](int t) {
        if (rows == 1) {
            const size_t a = total * t / nt, b = total * (t + 1) / nt;
            memcpy(dst + a, src + a, b - a);
        } else {
            const size_t a = rows * t / nt, b = rows * (t + 1) / nt;
            if (dst_pitch == row_bytes && src_pitch == row_bytes) {
                if (b > a) memcpy(dst + a * row_bytes, src + a * row_bytes, (b - a) * row_bytes);
            } else {
                for (size_t r = a; r < b; ++r) memcpy(dst + r * dst_pitch, src + r * src_pitch, row_bytes);
            }
        }
    };
    if (nt == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
}

class Stager {
public:
    static constexpr int kSlots = 4;
    static constexpr size_t kSlotBytes = 64u << 20;
    std::mutex call_mu;                 // one host-tier call at a time owns the ring

    ~Stager() { release(); }
    void release()
    {
        for (int i = 0; i < kSlots; ++i) {
            if (ev[i]) cudaEventDestroy(ev[i]);
            if (slot[i]) cudaFreeHost(slot[i]);
            ev[i] = nullptr;
            slot[i] = nullptr;
        }
    }
    // H2D of a (rows x row_bytes) host region with row pitch src_pitch into CONTIGUOUS device memory, on stream st.
    // Pinned sources: one asynchronous (2-D) copy.  Pageable sources: through the ring; returns once the last piece
    // has been handed to the stream (the host is busy copying meanwhile, the device side stays asynchronous).
    int h2d(void *dst_dev, const void *src_host, size_t src_pitch, size_t row_bytes, size_t rows, bool pinned,
            cudaStream_t st)
    {
        if (rows == 0 || row_bytes == 0) return VELO_OK;
        const char *src = static_cast<const char *>(src_host);
        char *dst = static_cast<char *>(dst_dev);
        if (pinned) {
            if (src_pitch == row_bytes || rows == 1)
                VELO_CUDA_TRY(cudaMemcpyAsync(dst, src, row_bytes * rows, cudaMemcpyHostToDevice, st));
            else
                VELO_CUDA_TRY(cudaMemcpy2DAsync(dst, row_bytes, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st));
            return VELO_OK;
        }
        int rc = ensure();
        if (rc) return rc;
        if (row_bytes > kSlotBytes) {                      // very long rows: cut every row into slot-sized pieces
            for (size_t r = 0; r < rows; ++r)
                for (size_t off = 0; off < row_bytes; off += kSlotBytes) {
                    const size_t n = row_bytes - off < kSlotBytes ? row_bytes - off : kSlotBytes;
                    if ((rc = piece_h2d(dst + r * row_bytes + off, src + r * src_pitch + off, n, n, 1, st))) return rc;
                }
            return VELO_OK;
        }
        const size_t rows_per = kSlotBytes / row_bytes;
        for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
            const size_t nr = rows - r0 < rows_per ? rows - r0 : rows_per;
            if ((rc = piece_h2d(dst + r0 * row_bytes, src + r0 * src_pitch, src_pitch, row_bytes, nr, st))) return rc;
        }
        return VELO_OK;
    }
    // D2H of `bytes` contiguous device bytes into host memory; blocks until the data is in place when the
    // destination is pageable (callers place it where the GPU has other work queued).
    int d2h(void *dst_host, const void *src_dev, size_t bytes, bool pinned, cudaStream_t st)
    {
        if (bytes == 0) return VELO_OK;
        if (pinned) {
            VELO_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
            return VELO_OK;
        }
        int rc = ensure();
        if (rc) return rc;
        // two slots in flight: the copy of piece k+1 runs while piece k is moved out of its slot
        size_t off_issue = 0, off_done = 0;
        int q[kSlots];
        size_t qn[kSlots];
        int head = 0, tail = 0, inflight = 0;
        while (off_done < bytes) {
            while (inflight < 2 && off_issue < bytes) {
                const int i = acquire();
                if (i < 0) return VELO_E_CUDA;
                const size_t n = bytes - off_issue < kSlotBytes ? bytes - off_issue : kSlotBytes;
                VELO_CUDA_TRY(cudaMemcpyAsync(slot[i], static_cast<const char *>(src_dev) + off_issue, n,
                                              cudaMemcpyDeviceToHost, st));
                VELO_CUDA_TRY(cudaEventRecord(ev[i], st));
                q[tail] = i;
                qn[tail] = n;
                tail = (tail + 1) % kSlots;
                ++inflight;
                off_issue += n;
            }
            const int i = q[head];
            const size_t n = qn[head];
            head = (head + 1) % kSlots;
            --inflight;
            VELO_CUDA_TRY(cudaEventSynchronize(ev[i]));
            parallel_copy_2d(static_cast<char *>(dst_host) + off_done, n, static_cast<const char *>(slot[i]), n, n, 1);
            off_done += n;
        }
        return VELO_OK;
    }

private:
    void *slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    int next = 0;

    int ensure()
    {
        for (int i = 0; i < kSlots; ++i) {
            if (!slot[i]) VELO_CUDA_TRY(cudaMallocHost(&slot[i], kSlotBytes));
            if (!ev[i]) {
                VELO_CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
                VELO_CUDA_TRY(cudaEventRecord(ev[i], nullptr));
            }
        }
        return VELO_OK;
    }
    int acquire()                          // next slot of the ring, once its previous transfer has completed
    {
        const int i = next;
        next = (next + 1) % kSlots;
        if (cudaEventSynchronize(ev[i]) != cudaSuccess) {
            set_error("staging ring: cudaEventSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
            return -1;
        }
        return i;
    }
    int piece_h2d(char *dst_dev, const char *src, size_t src_pitch, size_t row_bytes, size_t rows, cudaStream_t st)
    {
        const int i = acquire();
        if (i < 0) return VELO_E_CUDA;
        parallel_copy_2d(static_cast<char *>(slot[i]), row_bytes, src, src_pitch, row_bytes, rows);
        VELO_CUDA_TRY(cudaMemcpyAsync(dst_dev, slot[i], row_bytes * rows, cudaMemcpyHostToDevice, st));
        VELO_CUDA_TRY(cudaEventRecord(ev[i], st));
        return VELO_OK;
    }
};
static Stager g_stager;

// Upload a host gene-major block -- G rows of nc values at a row pitch of src_cols values (src_cols == nc: a whole
// matrix; src_cols > nc: a cell block of a larger matrix) -- into nc cell-major fp32 device rows, gene-row chunk by
// chunk: H2D into a device staging buffer, transpose + convert there (no host arithmetic).
static int upload_cellmajor(const void *host_gc, int elem_bytes, int64_t G, int64_t nc, int64_t src_cols, float *dst_cg,
                            int64_t ld, cudaStream_t st, float *dst_lo = nullptr, int *nz_flag = nullptr)
{
    const int64_t row_bytes = nc * elem_bytes;
    int64_t chunk_rows = (256LL << 20) / (row_bytes > 0 ? row_bytes : 1);
    if (chunk_rows < 32) chunk_rows = 32;
    if (chunk_rows > G) chunk_rows = G;
    chunk_rows = chunk_rows > 32 ? chunk_rows / 32 * 32 : chunk_rows;
    const bool pinned = host_ptr_is_pinned(host_gc);
    DevBuf stage;                   // the transpose of chunk k (HBM speed) is ~2 % of its PCIe time: one buffer, one stream
    int rc = stage.alloc(static_cast<size_t>(chunk_rows * row_bytes), st);
    if (rc) return rc;
    for (int64_t g0 = 0; g0 < G; g0 += chunk_rows) {
        const int64_t rows = (G - g0 < chunk_rows) ? G - g0 : chunk_rows;
        const char *src = static_cast<const char *>(host_gc) + g0 * src_cols * elem_bytes;
        if ((rc = g_stager.h2d(stage.p, src, static_cast<size_t>(src_cols * elem_bytes), static_cast<size_t>(row_bytes),
                               static_cast<size_t>(rows), pinned, st)))
            return rc;
        rc = velo_dev_pack_cellmajor_split(stage.p, elem_bytes, rows, nc, dst_cg, dst_lo, nz_flag, ld, g0, st);
        if (rc) return rc;
    }
    return VELO_OK;
}

// ------------------------------------------------------------------------------------------------------------
// The cell pipeline of the compact partial path.  The expression matrix of ALL cells has to be resident before
// the first cell can be processed (neighbours are arbitrary cells); a cell's velocity row d_c, its neighbour list
// and its output row are only needed while that cell is being processed.  So d / ixs / out move in cell chunks
// on their own streams underneath the correlation kernel of the neighbouring chunk.  Works on the local cells
// [c0, c0 + nc) of a C-cell problem: the single-GPU host call uses it with (0, C), a rank of the cell-sharded
// multi-GPU path with its own block after the all-gather of e.
// ------------------------------------------------------------------------------------------------------------
struct PipeStreams {                 // destroyed after every buffer has been handed back (cudaFreeAsync)
    cudaStream_t st = nullptr, cp = nullptr, dd = nullptr;
    std::vector<cudaEvent_t> ev;
    void sync()
    {
        if (st) cudaStreamSynchronize(st);
        if (cp) cudaStreamSynchronize(cp);
        if (dd) cudaStreamSynchronize(dd);
    }
    int create()
    {
        VELO_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        VELO_CUDA_TRY(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
        VELO_CUDA_TRY(cudaStreamCreateWithFlags(&dd, cudaStreamNonBlocking));
        return VELO_OK;
    }
    int new_event(cudaEvent_t *e)
    {
        VELO_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        ev.push_back(*e);
        return VELO_OK;
    }
    ~PipeStreams()
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
struct Quiesce {                     // declared LAST in a scope: nothing is still using the buffers when they go
    PipeStreams &s;
    ~Quiesce() { s.sync(); }
};

// e_ready: event after which e_cm / e_lo are complete (recorded by the caller), or nullptr.
// d: host gene-major, G rows of nc values for the local cells at a row pitch of d_cols values.
// ixs / out: host, nc x m, rows of the local cells.
static int pipeline_cells(PipeStreams &S, int transform, int rule, const float *e_cm, const float *e_lo, int64_t ld,
                          cudaEvent_t e_ready, const int *lo_nonzero_host, const void *d, int elem_bytes, int64_t d_cols,
                          const int64_t *ixs, float *out_host, int64_t G, int64_t C, int64_t c0_glob, int64_t nc,
                          int64_t m, double psc, double sigma, const DeviceProps &dp)
{
    const int64_t eb = elem_bytes;
    // chunking: a multiple of the SM count (K1 strides cells statically over one persistent CTA per SM)
    int64_t nchunks = (nc * G * eb >= (2LL << 30)) ? 8 : 1;
    int64_t chunk = round_up((nc + nchunks - 1) / nchunks, dp.sm_count > 0 ? dp.sm_count : 148);
    if (const char *env = getenv("VELO_HOST_CHUNK_CELLS")) {      // test hook: force the multi-chunk pipeline on small inputs
        const long long v = atoll(env);
        if (v > 0) chunk = v;
    }
    if (chunk > nc) chunk = nc;
    nchunks = (nc + chunk - 1) / chunk;
    const bool d_pinned = host_ptr_is_pinned(d), ix_pinned = host_ptr_is_pinned(ixs), out_pinned = host_ptr_is_pinned(out_host);

    DevBuf d_cm, stats, out, ix32, stage_d[2], stage_ix[2], flags;
    PinnedInts bad;
    Quiesce quiesce{S};
    cudaStream_t st = S.st, cp = S.cp, dd = S.dd;
    int rc;
    if ((rc = d_cm.alloc(static_cast<size_t>(nc * ld) * 4, st))) return rc;
    if ((rc = stats.alloc(static_cast<size_t>(nc) * 2 * 4, st))) return rc;
    if ((rc = out.alloc(static_cast<size_t>(nc * m) * 4, st))) return rc;
    if ((rc = ix32.alloc(static_cast<size_t>(nc * m) * 4, st))) return rc;
    const int nstage = nchunks > 1 ? 2 : 1;                       // double-buffered staging: copy j+1 under convert j
    for (int i = 0; i < nstage; ++i) {
        if ((rc = stage_d[i].alloc(static_cast<size_t>(chunk * G * eb), st))) return rc;
        if ((rc = stage_ix[i].alloc(static_cast<size_t>(chunk * m) * 8, st))) return rc;
    }
    if ((rc = flags.alloc(static_cast<size_t>(nchunks) * sizeof(int), st))) return rc;
    VELO_CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&bad.p), static_cast<size_t>(nchunks) * sizeof(int)));
    VELO_CUDA_TRY(cudaMemsetAsync(flags.p, 0, static_cast<size_t>(nchunks) * sizeof(int), st));
    if (ld != G) VELO_CUDA_TRY(cudaMemsetAsync(d_cm.p, 0, static_cast<size_t>(nc * ld) * 4, st));
    cudaEvent_t ev_alloc;
    if ((rc = S.new_event(&ev_alloc))) return rc;
    VELO_CUDA_TRY(cudaEventRecord(ev_alloc, st));
    VELO_CUDA_TRY(cudaStreamWaitEvent(cp, ev_alloc, 0));          // the copy stream may touch the buffers from here on

    std::vector<cudaEvent_t> ev_ready(nchunks), ev_out(nchunks);
    auto enqueue_copy = [&](int64_t j) -> int {                   // H2D + layout conversion of chunk j, copy stream
        const int64_t c0 = j * chunk, n = (nc - c0 < chunk) ? nc - c0 : chunk;
        const int b = static_cast<int>(j % nstage);
        int r = g_stager.h2d(stage_d[b].p, static_cast<const char *>(d) + c0 * eb, static_cast<size_t>(d_cols * eb),
                             static_cast<size_t>(n * eb), static_cast<size_t>(G), d_pinned, cp);
        if (r) return r;
        if ((r = velo_dev_pack_cellmajor(stage_d[b].p, elem_bytes, G, n, d_cm.as<float>() + c0 * ld, ld, 0, cp))) return r;
        if ((r = g_stager.h2d(stage_ix[b].p, ixs + c0 * m, static_cast<size_t>(n * m) * 8, static_cast<size_t>(n * m) * 8, 1,
                              ix_pinned, cp)))
            return r;
        if ((r = i64_to_i32_checked(stage_ix[b].as<int64_t>(), ix32.as<int32_t>() + c0 * m, n * m, C, flags.as<int>() + j, cp)))
            return r;
        VELO_CUDA_TRY(cudaMemcpyAsync(bad.p + j, flags.as<int>() + j, sizeof(int), cudaMemcpyDeviceToHost, cp));
        if ((r = S.new_event(&ev_ready[j]))) return r;
        VELO_CUDA_TRY(cudaEventRecord(ev_ready[j], cp));
        return VELO_OK;
    };
    auto fetch_out = [&](int64_t j) -> int {                      // D2H of chunk j's result
        const int64_t c0 = j * chunk, n = (nc - c0 < chunk) ? nc - c0 : chunk;
        VELO_CUDA_TRY(cudaStreamWaitEvent(dd, ev_out[j], 0));
        return g_stager.d2h(out_host + c0 * m, out.as<float>() + c0 * m, static_cast<size_t>(n * m) * 4, out_pinned, dd);
    };
    if ((rc = enqueue_copy(0))) return rc;
    if (e_ready) VELO_CUDA_TRY(cudaStreamWaitEvent(st, e_ready, 0));
    for (int64_t j = 0; j < nchunks; ++j) {
        const int64_t c0 = j * chunk, n = (nc - c0 < chunk) ? nc - c0 : chunk;
        VELO_CUDA_TRY(cudaEventSynchronize(ev_ready[j]));
        // the kernel must not gather through an out-of-range index: check before launching it
        VELO_REQUIRE(bad.p[j] == 0, "colDeltaCor: ixs holds an index outside [0, %lld)", static_cast<long long>(C));
        if (e_ready && j == 0) VELO_CUDA_TRY(cudaEventSynchronize(e_ready));     // *lo_nonzero_host is final
        const bool use_lo = e_lo && (!lo_nonzero_host || *lo_nonzero_host);
        VELO_CUDA_TRY(cudaStreamWaitEvent(st, ev_ready[j], 0));
        if ((rc = velo_dev_cell_stats(d_cm.as<float>() + c0 * ld, ld, G, n, stats.as<float>() + 2 * c0, st))) return rc;
        if ((rc = velo_dev_coldeltacor_ex(transform, rule, e_cm, use_lo ? e_lo : nullptr, d_cm.as<float>() + c0 * ld, ld,
                                          stats.as<float>() + 2 * c0, ix32.as<int32_t>() + c0 * m, m,
                                          out.as<float>() + c0 * m, m, G, C, c0_glob + c0, n, m, psc, st)))
            return rc;
        if (sigma > 0.0 && (rc = velo_dev_transition_prob(out.as<float>() + c0 * m, m, ix32.as<int32_t>() + c0 * m, m,
                                                          out.as<float>() + c0 * m, m, c0_glob + c0, n, m, sigma, st)))
            return rc;
        if ((rc = S.new_event(&ev_out[j]))) return rc;
        VELO_CUDA_TRY(cudaEventRecord(ev_out[j], st));
        // both of these fly under chunk j's kernel (a pageable source / destination keeps the HOST busy meanwhile,
        // which is why the kernel is launched first)
        if (j + 1 < nchunks && (rc = enqueue_copy(j + 1))) return rc;
        if (j > 0 && (rc = fetch_out(j - 1))) return rc;
    }
    if ((rc = fetch_out(nchunks - 1))) return rc;
    VELO_CUDA_TRY(cudaStreamSynchronize(st));
    VELO_CUDA_TRY(cudaStreamSynchronize(dd));
    VELO_CUDA_TRY(cudaStreamSynchronize(cp));
    return VELO_OK;
}

// single-GPU compact partial path: e goes up first (copy stream), the cell pipeline queues up right behind it
static int host_partial_compact_pipelined(int transform, int rule, const void *e, const void *d, int elem_bytes,
                                          const int64_t *ixs, int64_t G, int64_t C, int64_t m, double psc,
                                          float *out_compact, double sigma, const DeviceProps &dp)
{
    std::lock_guard<std::mutex> ring(g_stager.call_mu);
    const int64_t ld = round_up(G, 32);
    PipeStreams S;
    DevBuf e_cm, e_lo, lo_flag;
    PinnedInts lo_nz;
    Quiesce quiesce{S};
    int rc;
    if ((rc = S.create())) return rc;
    cudaStream_t st = S.st, cp = S.cp;
    if ((rc = e_cm.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
    if (ld != G) VELO_CUDA_TRY(cudaMemsetAsync(e_cm.p, 0, static_cast<size_t>(C * ld) * 4, st));
    // fp64 inputs + a transform that jumps at zero difference: keep the fp32 residuals of e
    const double jump = transform == VELO_SQRT ? 2.0 * sqrt(psc > 0 ? psc : 0.0)
                        : transform == VELO_LOG10 ? 2.0 * fabs(log10(psc > 0 ? psc : 1e-300)) : 0.0;
    const bool want_lo = elem_bytes == 8 && jump > 1e-4;
    VELO_CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&lo_nz.p), sizeof(int)));
    *lo_nz.p = 0;
    if (want_lo) {
        if ((rc = e_lo.alloc(static_cast<size_t>(C * ld) * 4, st))) return rc;
        if ((rc = lo_flag.alloc(sizeof(int), st))) return rc;
        VELO_CUDA_TRY(cudaMemsetAsync(e_lo.p, 0, static_cast<size_t>(C * ld) * 4, st));
        VELO_CUDA_TRY(cudaMemsetAsync(lo_flag.p, 0, sizeof(int), st));
    }
    cudaEvent_t ev_alloc, ev_e;
    if ((rc = S.new_event(&ev_alloc))) return rc;
    VELO_CUDA_TRY(cudaEventRecord(ev_alloc, st));
    VELO_CUDA_TRY(cudaStreamWaitEvent(cp, ev_alloc, 0));
    if ((rc = upload_cellmajor(e, elem_bytes, G, C, C, e_cm.as<float>(), ld, cp, want_lo ? e_lo.as<float>() : nullptr,
                               want_lo ? lo_flag.as<int>() : nullptr)))
        return rc;
    if (want_lo) VELO_CUDA_TRY(cudaMemcpyAsync(lo_nz.p, lo_flag.p, sizeof(int), cudaMemcpyDeviceToHost, cp));
    if ((rc = S.new_event(&ev_e))) return rc;
    VELO_CUDA_TRY(cudaEventRecord(ev_e, cp));
    return pipeline_cells(S, transform, rule, e_cm.as<float>(), want_lo ? e_lo.as<float>() : nullptr, ld, ev_e, lo_nz.p, d,
                          elem_bytes, C, ixs, out_compact, G, C, 0, C, m, psc, sigma, dp);
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
    std::lock_guard<std::mutex> ring(g_stager.call_mu);
    if ((rc = upload_cellmajor(e, elem_bytes, G, C, C, e_cm.as<float>(), ld, st, want_lo ? e_lo.as<float>() : nullptr,
                               want_lo ? lo_flag.as<int>() : nullptr)))
        return rc;
    if ((rc = upload_cellmajor(d, elem_bytes, G, C, C, d_cm.as<float>(), ld, st))) return rc;
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
extern "C" int velo_upload_cellmajor(const void *src_gc, int elem_bytes, int64_t G, int64_t nc, int64_t src_cols,
                                     float *dst_cm, float *dst_lo_cm, int *lo_nonzero, int64_t ld, velo_stream_t stream)
{
    VELO_REQUIRE(src_gc && dst_cm && G > 0 && nc > 0 && src_cols >= nc && ld >= G && (ld % 4) == 0,
                 "upload_cellmajor: bad arguments");
    VELO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "upload_cellmajor: elem_bytes must be 4 or 8");
    VELO_REQUIRE(dst_lo_cm == nullptr || (elem_bytes == 8 && lo_nonzero), "upload_cellmajor: residuals need fp64 input and a flag");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    std::lock_guard<std::mutex> ring(g_stager.call_mu);
    cudaStream_t st = as_stream(stream);
    DevBuf flag;
    if (dst_lo_cm) {
        if ((rc = flag.alloc(sizeof(int), st))) return rc;
        VELO_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    }
    if ((rc = upload_cellmajor(src_gc, elem_bytes, G, nc, src_cols, dst_cm, ld, st, dst_lo_cm, dst_lo_cm ? flag.as<int>() : nullptr)))
        return rc;
    if (dst_lo_cm) {
        VELO_CUDA_TRY(cudaMemcpyAsync(lo_nonzero, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        VELO_CUDA_TRY(cudaStreamSynchronize(st));          // *lo_nonzero is a host int: valid on return
    }
    return VELO_OK;
}

extern "C" int velo_transition_prob_partial_sharded(int transform, const float *e_all_cm, const float *e_lo_all_cm,
                                                    int64_t ld, velo_stream_t e_ready_stream, const void *d_block,
                                                    int elem_bytes, int64_t d_cols, const int64_t *ixs_block,
                                                    float *out_block, int64_t rows, int64_t cols, int64_t c0,
                                                    int64_t nc, int64_t nrndm, double psc, double sigma)
{
    VELO_REQUIRE(e_all_cm && d_block && ixs_block && out_block, "transition_prob_partial_sharded: null pointer");
    VELO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "transition_prob_partial_sharded: elem_bytes must be 4 or 8");
    VELO_REQUIRE(transform >= VELO_LINEAR && transform <= VELO_LOG10, "transition_prob_partial_sharded: unknown transform");
    VELO_REQUIRE(rows > 0 && cols > 0 && c0 >= 0 && nc >= 0 && c0 + nc <= cols && nrndm >= 0 && d_cols >= nc,
                 "transition_prob_partial_sharded: bad sizes");
    VELO_REQUIRE(ld >= rows && (ld % 4) == 0, "transition_prob_partial_sharded: ld must be >= rows and a multiple of 4");
    if (nc == 0 || nrndm == 0) return VELO_OK;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    std::lock_guard<std::mutex> ring(g_stager.call_mu);
    PipeStreams S;
    if ((rc = S.create())) return rc;
    cudaEvent_t e_ready;                                    // everything queued on the caller's stream so far (the
    if ((rc = S.new_event(&e_ready))) return rc;            // all-gather of e) precedes the first kernel
    VELO_CUDA_TRY(cudaEventRecord(e_ready, as_stream(e_ready_stream)));
    return pipeline_cells(S, transform, VELO_RULE_PARTIAL, e_all_cm, e_lo_all_cm, ld, e_ready, nullptr, d_block, elem_bytes,
                          d_cols, ixs_block, out_block, rows, cols, c0, nc, nrndm, psc, sigma, dp);
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
