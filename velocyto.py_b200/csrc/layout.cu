// Layout converters between the reference's host layout (gene-major rows x cols, fp64) and the
// device layout of this library (cell-major fp32 rows, 16-byte aligned, pad columns zero).
#include "velo_common.cuh"

namespace velo {

// SPLIT: also emit the fp32 residual lo = (float)(x - (double)(float)x) of fp64 sources and raise *nz_flag when
// any residual is non-zero (i.e. the data is not exactly representable in fp32).
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(256) k_pack_cellmajor(const T *__restrict__ src, int64_t G, int64_t C,
                                                        float *__restrict__ dst, float *__restrict__ dst_lo,
                                                        int *__restrict__ nz_flag, int64_t ld, int64_t g_off)
{
    __shared__ float tile[32][33];
    __shared__ float tile_lo[SPLIT ? 32 : 1][33];
    const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 32;
    const int64_t g0 = static_cast<int64_t>(blockIdx.y) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    bool nz = false;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t g = g0 + ty + k, c = c0 + tx;           // coalesced over cells (source rows)
        const T x = (g < G && c < C) ? src[g * C + c] : T(0);
        const float hi = static_cast<float>(x);
        tile[ty + k][tx] = hi;
        if (SPLIT) {
            const float lo = static_cast<float>(static_cast<double>(x) - static_cast<double>(hi));
            tile_lo[ty + k][tx] = lo;
            nz |= lo != 0.0f;
        }
    }
    // one flag write per CTA at most, and none once the flag is up: round 1 let EVERY thread with a non-zero residual
    // hit the same address (8 M same-address atomics per 256 MB chunk, ~0.4 s over a 24 GB upload)
    if (SPLIT) {
        if (__syncthreads_or(nz) && threadIdx.x == 0 && *reinterpret_cast<volatile int *>(nz_flag) == 0) atomicExch(nz_flag, 1);
    } else {
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t c = c0 + ty + k, g = g0 + tx;           // coalesced over genes (destination rows)
        if (c < C && g < G) {
            dst[c * ld + g_off + g] = tile[tx][ty + k];
            if (SPLIT) dst_lo[c * ld + g_off + g] = tile_lo[tx][ty + k];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_unpack_genemajor(const float *__restrict__ src, int64_t ld, int64_t G,
                                                          int64_t C, T *__restrict__ dst)
{
    __shared__ float tile[32][33];
    const int64_t g0 = static_cast<int64_t>(blockIdx.x) * 32;
    const int64_t c0 = static_cast<int64_t>(blockIdx.y) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t c = c0 + ty + k, g = g0 + tx;
        tile[ty + k][tx] = (c < C && g < G) ? src[c * ld + g] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t g = g0 + ty + k, c = c0 + tx;
        if (g < G && c < C) dst[g * C + c] = static_cast<T>(tile[tx][ty + k]);
    }
}

// bound > 0: values outside [0, bound) raise *flag (index validation happens on the device)
__global__ void k_i64_to_i32(const int64_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n, int64_t bound,
                             int *__restrict__ flag)
{
    bool bad = false;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t v = src[t];
        bad |= bound > 0 && (v < 0 || v >= bound);
        dst[t] = static_cast<int32_t>(v);
    }
    if (bad && flag) atomicExch(flag, 1);
}

int i64_to_i32_checked(const int64_t *src, int32_t *dst, int64_t n, int64_t bound, int *flag, cudaStream_t st)
{
    if (n == 0) return VELO_OK;
    const int64_t blocks = (n + 255) / 256;
    k_i64_to_i32<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, st>>>(src, dst, n, bound, flag);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

}  // namespace velo

using namespace velo;

extern "C" int velo_dev_pack_cellmajor(const void *src_gc, int elem_bytes, int64_t G, int64_t C, float *dst_cg,
                                       int64_t ld, int64_t g_off, velo_stream_t stream)
{
    return velo_dev_pack_cellmajor_split(src_gc, elem_bytes, G, C, dst_cg, nullptr, nullptr, ld, g_off, stream);
}

extern "C" int velo_dev_pack_cellmajor_split(const void *src_gc, int elem_bytes, int64_t G, int64_t C, float *dst_cg,
                                             float *dst_lo_cg, int *nonzero_flag, int64_t ld, int64_t g_off,
                                             velo_stream_t stream)
{
    VELO_REQUIRE(src_gc && dst_cg && G > 0 && C > 0 && g_off >= 0 && ld >= g_off + G, "pack_cellmajor: bad arguments");
    VELO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "pack_cellmajor: elem_bytes must be 4 or 8");
    VELO_REQUIRE(dst_lo_cg == nullptr || (elem_bytes == 8 && nonzero_flag), "pack_cellmajor: residuals need fp64 input and a flag");
    const int64_t gy = (G + 31) / 32;
    VELO_REQUIRE(gy <= 65535, "pack_cellmajor: too many genes per call (%lld); chunk the gene axis",
                 static_cast<long long>(G));
    dim3 grid(static_cast<unsigned>((C + 31) / 32), static_cast<unsigned>(gy));
    cudaStream_t st = as_stream(stream);
    if (dst_lo_cg)
        k_pack_cellmajor<double, true><<<grid, 256, 0, st>>>(static_cast<const double *>(src_gc), G, C, dst_cg,
                                                             dst_lo_cg, nonzero_flag, ld, g_off);
    else if (elem_bytes == 8)
        k_pack_cellmajor<double, false><<<grid, 256, 0, st>>>(static_cast<const double *>(src_gc), G, C, dst_cg,
                                                              nullptr, nullptr, ld, g_off);
    else
        k_pack_cellmajor<float, false><<<grid, 256, 0, st>>>(static_cast<const float *>(src_gc), G, C, dst_cg, nullptr,
                                                             nullptr, ld, g_off);
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_unpack_genemajor(const float *src_cg, int64_t ld, int64_t G, int64_t C, void *dst_gc,
                                         int elem_bytes, velo_stream_t stream)
{
    VELO_REQUIRE(src_cg && dst_gc && G > 0 && C > 0 && ld >= G, "unpack_genemajor: bad arguments");
    VELO_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "unpack_genemajor: elem_bytes must be 4 or 8");
    const int64_t gy = (C + 31) / 32;
    VELO_REQUIRE(gy <= 65535, "unpack_genemajor: too many cells per call (%lld)", static_cast<long long>(C));
    dim3 grid(static_cast<unsigned>((G + 31) / 32), static_cast<unsigned>(gy));
    if (elem_bytes == 8)
        k_unpack_genemajor<double><<<grid, 256, 0, as_stream(stream)>>>(src_cg, ld, G, C, static_cast<double *>(dst_gc));
    else
        k_unpack_genemajor<float><<<grid, 256, 0, as_stream(stream)>>>(src_cg, ld, G, C, static_cast<float *>(dst_gc));
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}

extern "C" int velo_dev_i64_to_i32(const int64_t *src, int32_t *dst, int64_t n, velo_stream_t stream)
{
    VELO_REQUIRE(src && dst && n >= 0, "i64_to_i32: bad arguments");
    return i64_to_i32_checked(src, dst, n, 0, nullptr, as_stream(stream));
}
