// K2g -- the all-pairs LINEAR correlation x_colDeltaCor (speedboosted.pyx:13-87) on the 5th-generation tensor
// cores.  It is the only GEMM-shaped member of the colDeltaCor family (SURVEY 8d, row "K2g"): with
//     x_c = e_c - gene_mean - mean_g(e_c - gene_mean)     (A = e_i - e_c is invariant under a per-gene shift)
//     b_c = d_c - mean_g(d_c)
// the reference's per-pair sums (pyx:30-78) collapse to two matrix products over the gene axis
//     sum_g (A - mean A) * b       =  P[c,i] - P[c,c]              P = B X^T
//     sum_g (A - mean A)^2         =  |x_i|^2 + |x_c|^2 - 2 Q[c,i]  Q = X X^T
//     corr[c,i] = (P[c,i] - P[c,c]) / sqrt((|x_i|^2 + |x_c|^2 - 2 Q[c,i]) * |b_c|^2)
//
// Precision (the contract is 1e-5 on the transition probabilities, i.e. ~5e-7 absolute on corr):
//   * every operand row is scaled by a power of two (max |x| * s in [2^13, 2^14)) and split x * s = hi + lo into two
//     fp16 matrices: 22 significant bits relative to the row maximum (a bf16 pair would give 16 and, measured on
//     the host, corr errors of 1e-6 at 500 genes; the fp16 pair gives 2e-8); hi*hi + hi*lo + lo*hi are three
//     `tcgen05.mma.kind::f16` per product and gene step, the dropped lo*lo term is 2^-24 of the product;
//   * the tensor core accumulates in fp32 WITHOUT round-to-nearest; over 30k genes that bias would reach 1e-4.
//     The TMEM accumulators therefore only ever hold the partial sum of ONE 64-gene block (12 accumulations);
//     eight epilogue warps drain them (tcgen05.ld) into fp32 running sums in registers (round-to-nearest adds)
//     while the MMAs of the next block run into the other TMEM buffer.
//
// Two kernels: k_coldeltacor_tc (one CTA per 128 x 128 tile, both accumulators in one CTA; kept for A/B runs) and the
// shipped k_coldeltacor_tc2 (CTA pairs, tcgen05.mma.cta_group::2, further down).  Common structure: warp 0 = TMA
// producer (128 x 64 fp16 operand tiles, SWIZZLE_128B boxes), warp 1 = MMA issuer (one thread), warps 2..9 = drain +
// final epilogue; tiles rasterised in groups of 16 cell blocks so the tiles in flight share operand rows through L2.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

#include "velo_common.cuh"

namespace velo {
namespace tc {

constexpr int kTM = 128, kTN = 128;                      // tile of pairs per CTA
constexpr int kGroup = 64;                               // genes per TMEM block sum (one drain)
#ifndef VELO_TC_BK
#define VELO_TC_BK 64
#endif
constexpr int kBK = VELO_TC_BK;                          // genes per pipeline stage: 64 (SWIZZLE_128B rows; 2 stages) or 32
                                                         // (SWIZZLE_64B, 4 stages: measured 10 % slower)
static_assert(kBK == 32 || kBK == 64, "kBK");
constexpr int kStages = kBK == 64 ? 2 : 4;               // 192 KB of operand stages either way
constexpr int kStagesPerGroup = kGroup / kBK;
constexpr uint32_t kRowBytes = kBK * 2u;                 // one swizzle row
constexpr uint32_t kTileBytes = 128u * kRowBytes;        // 8 / 16 KB per operand tile
constexpr uint32_t kStageBytes = 6u * kTileBytes;        // Bh_I, Bl_I, Xh_I, Xl_I, Xh_J, Xl_J
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);           // 320
constexpr uint32_t kTmemCols = 512;                      // 2 buffers x (P: 128 cols, Q: 128 cols)
constexpr int kRasterGroup = 16;
constexpr int kEpiPitch = 129;                           // fp32 row pitch of the parked sums (2 x 128 x 129 x 4 B = 129 KB)
constexpr size_t kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 128 /*barriers*/;

struct TcParams {
    const double *qd;       // C   : |x_c|^2
    const double *pcc;      // nc  : P[c,c]
    const float *isx;       // C   : 1 / (power-of-two scale of operand row x_c)
    const float *isb;       // nc  : 1 / (scale of operand row b_c)
    const float *stats;     // nc x 2 (mean_g d, sum_g (d - mean)^2)  -- velo_dev_cell_stats
    float *out;             // nc x out_ld
    float *dbgP, *dbgQ;     // optional raw products (nc x out_ld), NULL in production
    int64_t out_ld, C, c0, nc;
    int nkb;                // pipeline stages over the gene axis (Gp / kBK)
    int drain_groups;       // 64-gene groups per TMEM drain (1 in production; > 1 only to measure the accumulation bias)
    int nI, nJ;             // tile counts (cells, targets)
    // symmetric-Q scheme of the pair kernel (whole problem on one GPU: c0 == 0, nc == C); see k_coldeltacor_tc2
    const int2 *tiles;      // explicit (ti, tj) list for modes 1 / 2, nullptr = rasterise all nI x nJ tiles (mode 0)
    float *qscratch;        // C x C fp32: Q[c, i] written by mode 1, read transposed by mode 2
    int mode;               // 0: P and Q for every tile; 1: tiles on/right of the diagonal, also stores Q;
                            // 2: tiles left of the diagonal, TWO P tiles per pair (no Q MMAs), Q[c,i] = Q[i,c] from scratch
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("k_coldeltacor_tc: barrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *tm, int32_t x, int32_t y,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 x fp16 -> fp32, 128 x 128 x 16
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, K-major, rows of one swizzle span (128 B: SWIZZLE_128B = 2, 64 B: SWIZZLE_64B = 4),
// 8-row groups 8 * row bytes apart (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version [46,48) = 1, layout_type [61,64))
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
    constexpr uint64_t layout = kBK == 64 ? 2ull : 4ull;
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) |
           (static_cast<uint64_t>((8u * kRowBytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor (InstrDescriptor): c_format F32 [4,6) = 1, a/b_format F16 [7,10),[10,13) = 0, K-major A and B,
// n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(kTN >> 3) << 17) | (static_cast<uint32_t>(kTM >> 4) << 24);

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (thread = lane of its warp's quarter)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// ---- the kernel ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
k_coldeltacor_tc(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;              // SWIZZLE_128B tiles: 1024-byte aligned
    uint8_t *aligned = smem_raw + (base - raw_addr);
    uint64_t *bars = reinterpret_cast<uint64_t *>(aligned + kStages * kStageBytes);
    uint64_t *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages, *tempty = bars + 2 * kStages + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile of this CTA: groups of kRasterGroup cell blocks, targets fastest inside a group
    int ti, tj;
    {
        const int64_t t = blockIdx.x;
        const int64_t per_group = static_cast<int64_t>(kRasterGroup) * p.nJ;
        const int g = static_cast<int>(t / per_group);
        const int r = static_cast<int>(t - g * per_group);
        const int gsz = min(kRasterGroup, p.nI - g * kRasterGroup);
        ti = g * kRasterGroup + r % gsz;
        tj = r / gsz;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], kEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 1) {   // TMEM allocation: one full warp, which also owns the deallocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const int32_t rowB = ti * kTM;                                   // local cell rows of B
            const int32_t rowI = static_cast<int32_t>(p.c0) + ti * kTM;      // the same cells in X
            const int32_t rowJ = tj * kTN;                                   // targets in X
            for (int kb = 0; kb < p.nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait_bounded(&empty[s], ph ^ 1);                        // slot free (passes at once on first use)
                mbar_expect_tx(&full[s], kStageBytes);
                const uint32_t dst = base + s * kStageBytes;
                const int32_t g = kb * kBK;
                tma_load_2d(dst + 0 * kTileBytes, &tmBh, g, rowB, &full[s]);
                tma_load_2d(dst + 1 * kTileBytes, &tmBl, g, rowB, &full[s]);
                tma_load_2d(dst + 2 * kTileBytes, &tmXh, g, rowI, &full[s]);
                tma_load_2d(dst + 3 * kTileBytes, &tmXl, g, rowI, &full[s]);
                tma_load_2d(dst + 4 * kTileBytes, &tmXh, g, rowJ, &full[s]);
                tma_load_2d(dst + 5 * kTileBytes, &tmXl, g, rowJ, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const int per_drain = p.drain_groups * kStagesPerGroup;              // stages per TMEM block sum
            for (int kb = 0; kb < p.nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                const int dr = kb / per_drain, in_dr = kb - dr * per_drain;      // drain index, stage inside it
                const int b = dr & 1;
                const uint32_t bph = (dr >> 1) & 1;
                if (in_dr == 0) mbar_wait_bounded(&tempty[b], bph ^ 1);          // accumulator buffer drained
                mbar_wait_bounded(&full[s], ph);                                 // operands landed
                tcgen05_fence_after();
                const uint64_t d0 = make_smem_desc(base + s * kStageBytes);      // tile t, gene step k: + (t * tile + k * 32 B) >> 4
                const uint32_t dP = tmem_base + b * 256, dQ = dP + 128;
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t ko = static_cast<uint64_t>(k * 2);            // 16 fp16 = 32 bytes along the swizzled row
                    const uint64_t aBh = d0 + (0 * (kTileBytes >> 4) + ko), aBl = d0 + (1 * (kTileBytes >> 4) + ko);
                    const uint64_t aXh = d0 + (2 * (kTileBytes >> 4) + ko), aXl = d0 + (3 * (kTileBytes >> 4) + ko);
                    const uint64_t bXh = d0 + (4 * (kTileBytes >> 4) + ko), bXl = d0 + (5 * (kTileBytes >> 4) + ko);
                    const uint32_t acc = (in_dr > 0 || k > 0) ? 1u : 0u;         // first MMA of a block sum overwrites
                    umma_f16(dP, aBh, bXh, kIdesc, acc);
                    umma_f16(dP, aBh, bXl, kIdesc, 1);
                    umma_f16(dP, aBl, bXh, kIdesc, 1);
                    umma_f16(dQ, aXh, bXh, kIdesc, acc);
                    umma_f16(dQ, aXh, bXl, kIdesc, 1);
                    umma_f16(dQ, aXl, bXh, kIdesc, 1);
                }
                tcgen05_commit(&empty[s]);                                       // smem slot reusable when these MMAs retire
                if (in_dr == per_drain - 1 || kb == p.nkb - 1) tcgen05_commit(&tfull[b]);   // block sums are in TMEM
            }
        }
        __syncwarp();
    } else {
        // ===== drain + epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).., columns 64*h.. =====
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        float accP[64], accQ[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) accP[j] = accQ[j] = 0.f;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const int per_drain = p.drain_groups * kStagesPerGroup;
        const int ndrains = (p.nkb + per_drain - 1) / per_drain;
        for (int dr = 0; dr < ndrains; ++dr) {
            const int b = dr & 1;
            const uint32_t bph = (dr >> 1) & 1;
            mbar_wait_bounded(&tfull[b], bph);
            tcgen05_fence_after();
            const uint32_t t0 = tmem_base + lane_base + b * 256 + h * 64;
            float v[32];
            tmem_ld32(t0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) accP[j] += v[j];
            tmem_ld32(t0 + 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) accP[32 + j] += v[j];
            tmem_ld32(t0 + 128, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) accQ[j] += v[j];
            tmem_ld32(t0 + 160, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) accQ[32 + j] += v[j];
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
        }
        // final epilogue.  All MMAs have retired (last tfull), so the operand stages are free: park the sums there as
        // two 128 x 129 fp32 tiles (odd row pitch: thread = row writes are conflict-free), then let each warp finish
        // whole rows with lanes = consecutive targets (coalesced 128-byte stores).
        float *sP = reinterpret_cast<float *>(aligned);
        float *sQ = sP + kTM * kEpiPitch;
        {
            const int row = q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                sP[row * kEpiPitch + h * 64 + j] = accP[j];
                sQ[row * kEpiPitch + h * 64 + j] = accQ[j];
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");       // the 8 epilogue warps only
        const int ew = warp - 2;
        for (int row = ew; row < kTM; row += kEpiWarps) {
            const int64_t r = static_cast<int64_t>(ti) * kTM + row;
            if (r >= p.nc) break;
            const int64_t c = p.c0 + r;
            const double qc = p.qd[c], pc = p.pcc[r], sb = p.stats[2 * r + 1];
            const double isx_c = p.isx[c], isb_r = p.isb[r];
#pragma unroll
            for (int cc = 0; cc < kTN / 32; ++cc) {
                const int col = cc * 32 + lane;
                const int64_t i = static_cast<int64_t>(tj) * kTN + col;
                if (i < p.C) {
                    const double qi = p.qd[i], isx_i = p.isx[i];
                    const double Pt = static_cast<double>(sP[row * kEpiPitch + col]) * (isb_r * isx_i);   // undo the
                    const double Qt = static_cast<double>(sQ[row * kEpiPitch + col]) * (isx_c * isx_i);   // row scales
                    const double dist2 = qc + qi - 2.0 * Qt;
                    const double num = Pt - pc;
                    // self pair and coincident cells: the reference has 0 * inf = NaN there (pyx:57-78); the
                    // products resolve |x_i - x_c|^2 only down to ~1e-6 of the norms
                    const bool degenerate = (i == c) || !(dist2 > 4e-6 * (qc + qi)) || !(sb > 0.0);
                    p.out[r * p.out_ld + i] =
                        degenerate ? __int_as_float(0x7fc00000) : static_cast<float>(num * rsqrt(dist2 * sb));
                    if (p.dbgP) {
                        p.dbgP[r * p.out_ld + i] = static_cast<float>(Pt);
                        p.dbgQ[r * p.out_ld + i] = static_cast<float>(Qt);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

}  // namespace tc

// ---- the two-CTA kernel (cta_group::2) ---------------------------------------------------------------------
// The one-CTA kernel above is bound by the shared-memory port, not by the tensor pipe: per 64 genes its MMAs read
// 24 x 8 KB of operands and TMA writes another 96 KB through the same 128 B/clk port (ncu: tensor pipe 70 % active,
// l1tex tc wavefronts 70 %).  Here a CTA PAIR works on 128 cells x 256 targets with ONE M = 256 instruction stream:
//   rows   0..127 of the A operand live in CTA 0 and are B rows  -> its TMEM accumulates P = B X^T
//   rows 128..255 of the A operand live in CTA 1 and are X rows  -> its TMEM accumulates Q = X X^T
//   the N = 256 target rows of X are staged half in each CTA (tcgen05.mma.cta_group::2 reads both halves)
// Per CTA and 64 genes: operands 64 KB (was 96), MMA operand reads 12 x 8 KB (was 24 x 8 KB) for the same 1536 tensor
// cycles; one 256-column accumulator per CTA, double buffered (512 TMEM columns), 128 running sums per epilogue
// thread as before.  CTA 0 issues all MMAs; completion is multicast to both CTAs' barriers; the peer's TMA loads and
// drain arrivals signal CTA 0's barriers through the cluster shared-memory window.
namespace tc2 {
constexpr int kTM = 128, kTN = 256, kBK = 64;
constexpr int kStages = 3;
constexpr uint32_t kTileBytes = 128u * kBK * 2u;         // 16 KB
constexpr uint32_t kStageBytes = 4u * kTileBytes;        // A hi, A lo, target-half hi, target-half lo
constexpr int kEpiPitch = 257;                           // parked sums: 128 x 257 fp32 = 129 KB per CTA
constexpr size_t kSmemBytes = kStages * kStageBytes + 1024 + 128;
constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t cluster_addr)
{
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}
// TMA into this CTA's shared memory, completion bytes on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap *tm, int32_t x, int32_t y,
                                                 uint32_t bar_cluster_addr)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have retired
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t *bar)
{
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr)
{
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc::kThreads, 1)
k_coldeltacor_tc2(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const tc::TcParams p)
{
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t *aligned = smem_raw + (base - raw_addr);
    uint64_t *bars = reinterpret_cast<uint64_t *>(aligned + kStages * kStageBytes);
    uint64_t *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages, *tempty = bars + 2 * kStages + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    int ti, tj;   // pair tile: 128 cells x 256 targets; groups of kRasterGroup cell blocks, targets fastest inside a group
    if (p.tiles) {
        const int2 t = p.tiles[blockIdx.x >> 1];     // modes 1 / 2: host-built list (same raster idea, triangular sets)
        ti = t.x;                                    // mode 2: index u of the PAIR of cell blocks (2u, 2u + 1)
        tj = t.y;
    } else {
        const int64_t t = blockIdx.x >> 1;
        const int64_t per_group = static_cast<int64_t>(kRasterGroup) * p.nJ;
        const int g = static_cast<int>(t / per_group);
        const int r = static_cast<int>(t - g * per_group);
        const int gsz = min(kRasterGroup, p.nI - g * kRasterGroup);
        ti = g * kRasterGroup + r % gsz;
        tj = r / gsz;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);              // CTA 0's copy is the one in use: its producer's arrive.expect_tx
            mbar_init(&empty[s], 1);             // multicast commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);             // multicast commit
            mbar_init(&tempty[b], 2 * kEpiWarps);   // CTA 0's copy: the epilogue warps of both CTAs
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync();                              // barriers of BOTH CTAs initialised before any remote signal
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int per_drain = p.drain_groups;        // kBK == kGroup here: one stage per 64-gene group

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            // modes 0 / 1: A rows 0..127 (CTA 0) = the cells' b rows, 128..255 (CTA 1) = the same cells' x rows.
            // mode 2: BOTH halves are b rows -- of cell block 2u in CTA 0 and 2u + 1 in CTA 1 -- so the one M = 256
            // instruction stream yields two P tiles and no Q (it comes from the scratch, by symmetry).
            const bool two_p = p.mode == 2;
            const CUtensorMap *tmAh = (rank && !two_p) ? &tmXh : &tmBh, *tmAl = (rank && !two_p) ? &tmXl : &tmBl;
            const int32_t rowA = two_p ? (2 * ti + static_cast<int32_t>(rank)) * kTM
                                       : (rank ? static_cast<int32_t>(p.c0) + ti * kTM : ti * kTM);
            const int32_t rowJ = tj * kTN + static_cast<int32_t>(rank) * 128;
            for (int kb = 0; kb < p.nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait_bounded(&empty[s], ph ^ 1);
                if (rank == 0) mbar_expect_tx(&full[s], 2 * kStageBytes);    // this CTA's bytes and the peer's
                const uint32_t leader_full = mapa(smem_u32(&full[s]), 0);
                const uint32_t dst = base + s * kStageBytes;
                const int32_t g = kb * kBK;
                tma_load_2d_pair(dst + 0 * kTileBytes, tmAh, g, rowA, leader_full);
                tma_load_2d_pair(dst + 1 * kTileBytes, tmAl, g, rowA, leader_full);
                tma_load_2d_pair(dst + 2 * kTileBytes, &tmXh, g, rowJ, leader_full);
                tma_load_2d_pair(dst + 3 * kTileBytes, &tmXl, g, rowJ, leader_full);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of CTA 0 drives both tensor cores =====
        if (rank == 0 && lane == 0) {
            for (int kb = 0; kb < p.nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                const int dr = kb / per_drain, in_dr = kb - dr * per_drain;
                const int b = dr & 1;
                const uint32_t bph = (dr >> 1) & 1;
                if (in_dr == 0) mbar_wait_bounded(&tempty[b], bph ^ 1);
                mbar_wait_bounded(&full[s], ph);
                tcgen05_fence_after();
                const uint64_t d0 = make_smem_desc_sw128(base + s * kStageBytes);
                const uint32_t dD = tmem_base + b * 256;
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t ko = static_cast<uint64_t>(k * 2);
                    const uint64_t aH = d0 + ko, aL = d0 + ((kTileBytes >> 4) + ko);
                    const uint64_t bH = d0 + (2 * (kTileBytes >> 4) + ko), bL = d0 + (3 * (kTileBytes >> 4) + ko);
                    umma_f16_pair(dD, aH, bH, kIdesc, (in_dr > 0 || k > 0) ? 1u : 0u);
                    umma_f16_pair(dD, aH, bL, kIdesc, 1);
                    umma_f16_pair(dD, aL, bH, kIdesc, 1);
                }
                tcgen05_commit_pair(&empty[s]);
                if (in_dr == per_drain - 1 || kb == p.nkb - 1) tcgen05_commit_pair(&tfull[b]);
            }
        }
        __syncwarp();
    } else {
        // ===== drain: this CTA's accumulator (P in CTA 0, Q in CTA 1), lanes 32*(w%4).., columns 128*h.. =====
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        const int ndrains = (p.nkb + per_drain - 1) / per_drain;
        for (int dr = 0; dr < ndrains; ++dr) {
            const int b = dr & 1;
            const uint32_t bph = (dr >> 1) & 1;
            mbar_wait_bounded(&tfull[b], bph);
            tcgen05_fence_after();
            const uint32_t t0 = tmem_base + lane_base + b * 256 + h * 128;
            float v[32];
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                tmem_ld32(t0 + part * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[part * 32 + j] += v[j];
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&tempty[b]), 0));
        }
        // park the sums in this CTA's (now idle) operand stages: 128 x 257 fp32
        float *sS = reinterpret_cast<float *>(aligned);
        const int row = q * 32 + lane;
#pragma unroll
        for (int j = 0; j < 128; ++j) sS[row * kEpiPitch + h * 128 + j] = acc[j];
    }
    cluster_sync();                              // both tiles parked and visible cluster-wide
    if (warp >= 2 && p.mode == 2) {
        // two P tiles: CTA r owns all 128 rows of cell block 2u + r; P from its OWN parked sums, Q[c, i] = Q[i, c] from
        // the scratch mode 1 filled (row i of the scratch, column c: a strided read, 4 bytes per 32-byte sector --
        // C^2 / 2 such reads are ~15 ms at 50k cells against the ~0.2 s of MMAs they replace)
        const float *sS = reinterpret_cast<const float *>(aligned);
        const int ew = warp - 2;
        for (int row = ew; row < 128; row += kEpiWarps) {
            const int64_t r = (2 * static_cast<int64_t>(ti) + rank) * kTM + row;
            if (r >= p.nc) break;
            const int64_t c = r;                                   // c0 == 0 in this mode
            const double qc = p.qd[c], pc = p.pcc[r], sb = p.stats[2 * r + 1];
            const double isb_r = p.isb[r];
#pragma unroll
            for (int cc = 0; cc < kTN / 32; ++cc) {
                const int col = cc * 32 + lane;
                const int64_t i = static_cast<int64_t>(tj) * kTN + col;
                if (i < p.C) {
                    const double qi = p.qd[i], isx_i = p.isx[i];
                    const double Pt = static_cast<double>(sS[row * kEpiPitch + col]) * (isb_r * isx_i);
                    const double Qt = static_cast<double>(__ldg(p.qscratch + i * p.C + c));
                    const double dist2 = qc + qi - 2.0 * Qt;
                    const double num = Pt - pc;
                    const bool degenerate = (i == c) || !(dist2 > 4e-6 * (qc + qi)) || !(sb > 0.0);
                    p.out[r * p.out_ld + i] =
                        degenerate ? __int_as_float(0x7fc00000) : static_cast<float>(num * rsqrt(dist2 * sb));
                    if (p.dbgP) {
                        p.dbgP[r * p.out_ld + i] = static_cast<float>(Pt);
                        p.dbgQ[r * p.out_ld + i] = static_cast<float>(Qt);
                    }
                }
            }
        }
    } else if (warp >= 2) {
        // CTA r finishes rows 64*r .. 64*r+63 of the pair tile: P from CTA 0's shared memory, Q from CTA 1's
        const uint32_t sbase = smem_u32(aligned);
        const uint32_t baseP = mapa(sbase, 0), baseQ = mapa(sbase, 1);
        const int ew = warp - 2;
        for (int rr = ew; rr < 64; rr += kEpiWarps) {
            const int row = static_cast<int>(rank) * 64 + rr;
            const int64_t r = static_cast<int64_t>(ti) * kTM + row;
            if (r >= p.nc) break;
            const int64_t c = p.c0 + r;
            const double qc = p.qd[c], pc = p.pcc[r], sb = p.stats[2 * r + 1];
            const double isx_c = p.isx[c], isb_r = p.isb[r];
#pragma unroll
            for (int cc = 0; cc < kTN / 32; ++cc) {
                const int col = cc * 32 + lane;
                const int64_t i = static_cast<int64_t>(tj) * kTN + col;
                if (i < p.C) {
                    const uint32_t off = static_cast<uint32_t>(row * kEpiPitch + col) * 4u;
                    const double qi = p.qd[i], isx_i = p.isx[i];
                    const double Pt = static_cast<double>(ld_cluster_f32(baseP + off)) * (isb_r * isx_i);
                    const double Qt = static_cast<double>(ld_cluster_f32(baseQ + off)) * (isx_c * isx_i);
                    const double dist2 = qc + qi - 2.0 * Qt;
                    const double num = Pt - pc;
                    const bool degenerate = (i == c) || !(dist2 > 4e-6 * (qc + qi)) || !(sb > 0.0);
                    p.out[r * p.out_ld + i] =
                        degenerate ? __int_as_float(0x7fc00000) : static_cast<float>(num * rsqrt(dist2 * sb));
                    // mode 1: keep Q for the mirrored pair (the scales are powers of two: the fp32 store is exact)
                    if (p.mode == 1) p.qscratch[c * p.C + i] = static_cast<float>(Qt);
                    if (p.dbgP) {
                        p.dbgP[r * p.out_ld + i] = static_cast<float>(Pt);
                        p.dbgQ[r * p.out_ld + i] = static_cast<float>(Qt);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync();                              // nobody leaves while the peer still reads its shared memory
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}
}  // namespace tc2

namespace tc {
// ---- operand preparation ----------------------------------------------------------------------------------
// gene means over all cells (any per-gene shift leaves e_i - e_c unchanged; the mean keeps the operands small so
// that Q[c,i] is not dominated by the common expression profile): partial[y][g], fixed reduction order
__global__ void k_tc_gene_partial(const float *e_cm, int64_t ld, int64_t G, int64_t C, int parts, double *partial)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t per = (C + parts - 1) / parts;
    const int64_t a = static_cast<int64_t>(blockIdx.y) * per, b = min(C, a + per);
    double s = 0.0;
    for (int64_t c = a; c < b; ++c) s += static_cast<double>(__ldg(e_cm + c * ld + g));
    partial[static_cast<int64_t>(blockIdx.y) * G + g] = s;
}
__global__ void k_tc_gene_mean(const double *partial, int64_t G, int64_t C, int parts, float *mu)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    for (int y = 0; y < parts; ++y) s += partial[static_cast<int64_t>(y) * G + g];
    mu[g] = static_cast<float>(s / static_cast<double>(C));
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();                      // sh reuse
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += sh[i];   // same order in every thread: deterministic
    return t;
}

__device__ __forceinline__ float block_max(float v, float *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t = fmaxf(t, sh[i]);
    return t;
}
// power-of-two s with bound * s in [2^13, 2^14): fp16 holds the scaled row without overflow, hi*hi sums over one
// 64-gene block stay below 2^34
__device__ __forceinline__ float pow2_scale(float bound)
{
    if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
    int e;
    (void)frexpf(bound, &e);                  // bound = f * 2^e, f in [0.5, 1)
    return ldexpf(1.f, 14 - e);
}

// one CTA per cell: centred row x (fp32) -> scaled fp16 hi/lo; |x|^2; for local cells also b = d - mean, hi/lo, P[c,c]
__global__ void __launch_bounds__(256)
k_tc_prep(const float *e_cm, const float *d_cm, int64_t ld, const float *mu, const float *stats, int64_t G, int64_t Gp,
          int64_t C, int64_t c0, int64_t nc, __half *Xh, __half *Xl, __half *Bh, __half *Bl, double *qd, double *pcc,
          float *isx, float *isb)
{
    __shared__ double sh[8];
    __shared__ float shf[8];
    const int64_t c = blockIdx.x;
    const float *erow = e_cm + c * ld;
    const int64_t r = c - c0;
    const bool local = r >= 0 && r < nc;
    const float *drow = local ? d_cm + r * ld : nullptr;
    const float dm = local ? stats[2 * r] : 0.f;
    // quads of genes: rows are 16-byte aligned, ld % 4 == 0, mu holds Gp zero-padded entries
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    double s = 0.0;
    float mx = 0.f, mb = 0.f;
    for (int64_t g = 4 * static_cast<int64_t>(threadIdx.x); g < G; g += 4 * blockDim.x) {
        const float4 ev = *reinterpret_cast<const float4 *>(erow + g);
        const float4 mv = *reinterpret_cast<const float4 *>(mu + g);
        const float4 dv = local ? *reinterpret_cast<const float4 *>(drow + g) : z4;
        const float v[4] = {ev.x - mv.x, ev.y - mv.y, ev.z - mv.z, ev.w - mv.w};
        const float w[4] = {dv.x - dm, dv.y - dm, dv.z - dm, dv.w - dm};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (g + i < G) {
                s += static_cast<double>(v[i]);
                mx = fmaxf(mx, fabsf(v[i]));
                if (local) mb = fmaxf(mb, fabsf(w[i]));
            }
    }
    const float rm = static_cast<float>(block_sum(s, sh) / static_cast<double>(G));
    const float sx = pow2_scale(block_max(mx, shf) + fabsf(rm));      // |x| <= max |e - mu| + |row mean|
    const float sb = pow2_scale(block_max(mb, shf));
    double q = 0.0, pc = 0.0;
    for (int64_t g = 4 * static_cast<int64_t>(threadIdx.x); g < Gp; g += 4 * blockDim.x) {
        float x[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (g < G) {
            const float4 ev = *reinterpret_cast<const float4 *>(erow + g);
            const float4 mv = *reinterpret_cast<const float4 *>(mu + g);
            const float4 dv = local ? *reinterpret_cast<const float4 *>(drow + g) : z4;
            const float xe[4] = {ev.x - mv.x, ev.y - mv.y, ev.z - mv.z, ev.w - mv.w};
            const float be[4] = {dv.x - dm, dv.y - dm, dv.z - dm, dv.w - dm};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (g + i < G) {
                    x[i] = xe[i] - rm;
                    if (local) b[i] = be[i];
                }
        }
        __align__(8) __half xh[4], xl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float xs = x[i] * sx;
            xh[i] = __float2half_rn(xs);
            xl[i] = __float2half_rn(xs - __half2float(xh[i]));
            q += static_cast<double>(x[i]) * static_cast<double>(x[i]);
        }
        *reinterpret_cast<uint2 *>(Xh + c * Gp + g) = *reinterpret_cast<const uint2 *>(xh);
        *reinterpret_cast<uint2 *>(Xl + c * Gp + g) = *reinterpret_cast<const uint2 *>(xl);
        if (local) {
            __align__(8) __half bh[4], bl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float bs = b[i] * sb;
                bh[i] = __float2half_rn(bs);
                bl[i] = __float2half_rn(bs - __half2float(bh[i]));
                pc += static_cast<double>(x[i]) * static_cast<double>(b[i]);
            }
            *reinterpret_cast<uint2 *>(Bh + r * Gp + g) = *reinterpret_cast<const uint2 *>(bh);
            *reinterpret_cast<uint2 *>(Bl + r * Gp + g) = *reinterpret_cast<const uint2 *>(bl);
        }
    }
    q = block_sum(q, sh);
    pc = block_sum(pc, sh);
    if (threadIdx.x == 0) {
        qd[c] = q;
        isx[c] = 1.f / sx;
        if (local) {
            pcc[r] = pc;
            isb[r] = 1.f / sb;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn *fn)
{
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        VELO_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        VELO_REQUIRE(sym != nullptr && qres == cudaDriverEntryPointSuccess,
                     "coldeltacor_tc: the driver does not export cuTensorMapEncodeTiled");
        cached = reinterpret_cast<EncodeTiledFn>(sym);
    }
    *fn = cached;
    return VELO_OK;
}

// rows x Gp fp16 row-major matrix, boxes of 128 rows x kBK genes (one swizzle row), zero fill outside
static int make_map(EncodeTiledFn enc, CUtensorMap *tm, const void *ptr, int64_t rows, int64_t Gp)
{
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(Gp), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(Gp) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, kBK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VELO_REQUIRE(rc == CUDA_SUCCESS, "coldeltacor_tc: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(rc));
    return VELO_OK;
}

struct Scratch {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~Scratch()
    {
        if (p) cudaFreeAsync(p, st);
    }
    int alloc(size_t bytes, cudaStream_t s)
    {
        st = s;
        VELO_CUDA_TRY(cudaMallocAsync(&p, bytes ? bytes : 16, s));
        return VELO_OK;
    }
};

}  // namespace tc
}  // namespace velo

using namespace velo;

static std::atomic<int> g_tensor_cores{1};
extern "C" void velo_set_tensor_cores(int enable) { g_tensor_cores.store(enable ? 1 : 0); }
extern "C" int velo_get_tensor_cores(void) { return g_tensor_cores.load(); }

extern "C" size_t velo_coldeltacor_tc_workspace_bytes(int64_t G, int64_t C, int64_t nc)
{
    const int64_t Gp = round_up(G, tc::kGroup);
    // + the C x C fp32 copy of Q that the symmetric scheme keeps when the whole problem is on one GPU (nc == C)
    return static_cast<size_t>(2 * (C + nc) * Gp * 2 + (C + nc) * 12 + G * 4 + 64 * G * 8 + 4096) +
           (nc == C ? static_cast<size_t>(C) * C * 4 : 0);
}

extern "C" int velo_dev_coldeltacor_tc(const float *e_cm, const float *d_cm, int64_t ld, const float *stats, float *out,
                                       int64_t out_ld, int64_t G, int64_t C, int64_t c0, int64_t nc, float *dbgP,
                                       float *dbgQ, velo_stream_t stream)
{
    using namespace velo::tc;
    VELO_REQUIRE(e_cm && d_cm && stats && out, "coldeltacor_tc: null pointer");
    VELO_REQUIRE(G > 0 && C > 0 && nc >= 0 && c0 >= 0 && c0 + nc <= C, "coldeltacor_tc: bad sizes");
    VELO_REQUIRE(ld >= G && out_ld >= C, "coldeltacor_tc: leading dimensions too small");
    VELO_REQUIRE((ld % 4) == 0 && (reinterpret_cast<uintptr_t>(e_cm) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_cm) & 15) == 0,
                 "coldeltacor_tc: e_cm/d_cm must be 16-byte aligned with ld a multiple of 4");
    VELO_REQUIRE(C < (1LL << 31) - 256 && G < (1LL << 31) - 64, "coldeltacor_tc: sizes must fit int32 TMA coordinates");
    VELO_REQUIRE((dbgP == nullptr) == (dbgQ == nullptr), "coldeltacor_tc: dbgP and dbgQ go together");
    if (nc == 0) return VELO_OK;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const int64_t Gp = round_up(G, kGroup);
    const int parts = 64;

    Scratch xh, xl, bh, bl, qd, pcc, mu, partial, isx, isb;
    if ((rc = xh.alloc(static_cast<size_t>(C * Gp) * 2, st))) return rc;
    if ((rc = xl.alloc(static_cast<size_t>(C * Gp) * 2, st))) return rc;
    if ((rc = bh.alloc(static_cast<size_t>(nc * Gp) * 2, st))) return rc;
    if ((rc = bl.alloc(static_cast<size_t>(nc * Gp) * 2, st))) return rc;
    if ((rc = qd.alloc(static_cast<size_t>(C) * 8, st))) return rc;
    if ((rc = pcc.alloc(static_cast<size_t>(nc) * 8, st))) return rc;
    if ((rc = mu.alloc(static_cast<size_t>(Gp) * 4, st))) return rc;
    VELO_CUDA_TRY(cudaMemsetAsync(mu.p, 0, static_cast<size_t>(Gp) * 4, st));
    if ((rc = isx.alloc(static_cast<size_t>(C) * 4, st))) return rc;
    if ((rc = isb.alloc(static_cast<size_t>(nc) * 4, st))) return rc;
    if ((rc = partial.alloc(static_cast<size_t>(parts) * G * 8, st))) return rc;

    {
        dim3 grid(static_cast<unsigned>((G + 255) / 256), parts);
        k_tc_gene_partial<<<grid, 256, 0, st>>>(e_cm, ld, G, C, parts, static_cast<double *>(partial.p));
        VELO_LAUNCH_CHECK();
        k_tc_gene_mean<<<static_cast<unsigned>((G + 255) / 256), 256, 0, st>>>(static_cast<double *>(partial.p), G, C,
                                                                                parts, static_cast<float *>(mu.p));
        VELO_LAUNCH_CHECK();
        k_tc_prep<<<static_cast<unsigned>(C), 256, 0, st>>>(
            e_cm, d_cm, ld, static_cast<float *>(mu.p), stats, G, Gp, C, c0, nc, static_cast<__half *>(xh.p),
            static_cast<__half *>(xl.p), static_cast<__half *>(bh.p), static_cast<__half *>(bl.p),
            static_cast<double *>(qd.p), static_cast<double *>(pcc.p), static_cast<float *>(isx.p),
            static_cast<float *>(isb.p));
        VELO_LAUNCH_CHECK();
    }

    EncodeTiledFn enc;
    if ((rc = get_encode_fn(&enc))) return rc;
    CUtensorMap tmXh, tmXl, tmBh, tmBl;
    if ((rc = make_map(enc, &tmXh, xh.p, C, Gp))) return rc;
    if ((rc = make_map(enc, &tmXl, xl.p, C, Gp))) return rc;
    if ((rc = make_map(enc, &tmBh, bh.p, nc, Gp))) return rc;
    if ((rc = make_map(enc, &tmBl, bl.p, nc, Gp))) return rc;

    TcParams p;
    p.qd = static_cast<double *>(qd.p);
    p.pcc = static_cast<double *>(pcc.p);
    p.isx = static_cast<float *>(isx.p);
    p.isb = static_cast<float *>(isb.p);
    p.stats = stats;
    p.out = out;
    p.dbgP = dbgP;
    p.dbgQ = dbgQ;
    p.out_ld = out_ld;
    p.C = C;
    p.c0 = c0;
    p.nc = nc;
    p.nkb = static_cast<int>(Gp / kBK);
    p.drain_groups = 1;
    if (const char *env = getenv("VELO_TC_DRAIN_GROUPS")) {   // experiment knob (profiles/r1_k2g_*): bias of long TMEM sums
        const int v = atoi(env);
        if (v >= 1 && v <= 4096) p.drain_groups = v;
    }
    p.nI = static_cast<int>((nc + kTM - 1) / kTM);
    p.tiles = nullptr; p.qscratch = nullptr; p.mode = 0;
    // variant: 2 = CTA pairs (cta_group::2, 128 x 256 pair tiles; default), 1 = single CTAs (128 x 128 tiles)
    int variant = 2;
    if (const char *env = getenv("VELO_TC_VARIANT")) variant = atoi(env) == 1 ? 1 : 2;
    if (kBK != 64) variant = 1;                   // the SWIZZLE_64B experiment build only has the one-CTA kernel
    if (variant == 1) {
        p.nJ = static_cast<int>((C + kTN - 1) / kTN);
        const int64_t tiles = static_cast<int64_t>(p.nI) * p.nJ;
        VELO_REQUIRE(tiles < (1LL << 31), "coldeltacor_tc: too many tiles for one launch");
        VELO_REQUIRE(static_cast<size_t>(dp.smem_optin) >= kSmemBytes, "coldeltacor_tc: needs %zu bytes of shared memory",
                     kSmemBytes);
        VELO_CUDA_TRY(cudaFuncSetAttribute(k_coldeltacor_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kSmemBytes)));
        k_coldeltacor_tc<<<static_cast<unsigned>(tiles), kThreads, kSmemBytes, st>>>(tmXh, tmXl, tmBh, tmBl, p);
    } else {
        p.nJ = static_cast<int>((C + tc2::kTN - 1) / tc2::kTN);
        const int64_t pairs = static_cast<int64_t>(p.nI) * p.nJ;
        VELO_REQUIRE(2 * pairs < (1LL << 31), "coldeltacor_tc: too many tiles for one launch");
        VELO_REQUIRE(static_cast<size_t>(dp.smem_optin) >= tc2::kSmemBytes,
                     "coldeltacor_tc: needs %zu bytes of shared memory", tc2::kSmemBytes);
        VELO_CUDA_TRY(cudaFuncSetAttribute(tc2::k_coldeltacor_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(tc2::kSmemBytes)));
        // Symmetric-Q scheme: Q = X X^T is symmetric, so when the WHOLE problem is on this GPU only the tiles on or
        // right of the diagonal compute it (launch 1, which also stores it: C x C fp32 scratch); the tiles left of the
        // diagonal (launch 2) spend both halves of the M = 256 instruction on P tiles of two adjacent cell blocks and
        // read Q[c, i] = Q[i, c] back: 0.5 + 0.25 = 0.75 of the MMA cycles.  VELO_TC_SYMMETRIC=0 switches it off.
        bool symmetric = c0 == 0 && nc == C && p.nI >= 4;
        if (const char *env = getenv("VELO_TC_SYMMETRIC")) symmetric = symmetric && atoi(env) != 0;
        size_t free_b = 0, total_b = 0;
        if (symmetric && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            symmetric = static_cast<size_t>(C) * C * 4 + (2ull << 30) < free_b;
        Scratch qs, tl1, tl2;
        if (symmetric && qs.alloc(static_cast<size_t>(C) * C * 4, st) != VELO_OK) {
            (void)cudaGetLastError();
            symmetric = false;
        }
        if (!symmetric) {
            p.tiles = nullptr; p.qscratch = nullptr; p.mode = 0;
            tc2::k_coldeltacor_tc2<<<static_cast<unsigned>(2 * pairs), kThreads, tc2::kSmemBytes, st>>>(tmXh, tmXl, tmBh,
                                                                                                       tmBl, p);
        } else {
            std::vector<int2> upper, lower;
            for (int g0 = 0; g0 < p.nI; g0 += kRasterGroup) {          // launch 1: tj >= ti / 2, groups of cell blocks
                const int g1 = g0 + kRasterGroup < p.nI ? g0 + kRasterGroup : p.nI;
                for (int tj = g0 >> 1; tj < p.nJ; ++tj)
                    for (int ti = g0; ti < g1; ++ti)
                        if (tj >= (ti >> 1)) upper.push_back(make_int2(ti, tj));
            }
            const int nU = (p.nI + 1) / 2;                             // launch 2: pairs u of cell blocks, tj < u
            for (int g0 = 0; g0 < nU; g0 += kRasterGroup / 2) {
                const int g1 = g0 + kRasterGroup / 2 < nU ? g0 + kRasterGroup / 2 : nU;
                for (int tj = 0; tj < g1 - 1 && tj < p.nJ; ++tj)
                    for (int u = g0; u < g1; ++u)
                        if (tj < u) lower.push_back(make_int2(u, tj));
            }
            if ((rc = tl1.alloc(upper.size() * sizeof(int2), st))) return rc;
            if ((rc = tl2.alloc(lower.size() * sizeof(int2) + 16, st))) return rc;
            VELO_CUDA_TRY(cudaMemcpyAsync(tl1.p, upper.data(), upper.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
            if (!lower.empty())
                VELO_CUDA_TRY(cudaMemcpyAsync(tl2.p, lower.data(), lower.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
            VELO_CUDA_TRY(cudaStreamSynchronize(st));                  // the host vectors go out of scope below
            p.qscratch = static_cast<float *>(qs.p);
            p.tiles = static_cast<const int2 *>(tl1.p); p.mode = 1;
            tc2::k_coldeltacor_tc2<<<static_cast<unsigned>(2 * upper.size()), kThreads, tc2::kSmemBytes, st>>>(
                tmXh, tmXl, tmBh, tmBl, p);
            VELO_LAUNCH_CHECK();
            if (!lower.empty()) {
                p.tiles = static_cast<const int2 *>(tl2.p); p.mode = 2;
                tc2::k_coldeltacor_tc2<<<static_cast<unsigned>(2 * lower.size()), kThreads, tc2::kSmemBytes, st>>>(
                    tmXh, tmXl, tmBh, tmBl, p);
            }
        }
    }
    VELO_LAUNCH_CHECK();
    return VELO_OK;
}
