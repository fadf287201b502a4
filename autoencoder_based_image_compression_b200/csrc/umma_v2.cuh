// Version 2: A operand in TMEM, GDN / IGDN fused as a second contraction, clusters with TMA multicast (EAE_UMMA_VERSION=2).
// Private to conv_umma.cu (one translation unit): everything here sits in its anonymous namespace.
#pragma once

#include "umma_v1.cuh"

namespace eae {
namespace {

// =================================================================================================
// Version 2: the A operand lives in TMEM, GDN / IGDN fused as a second contraction.
//
//  * The 128 epilogue threads (thread = output row = TMEM lane) read their row of the TMA-staged
//    activation tile from shared memory, round it to TF32 (cvt.rna: hi) and store hi and, for the split
//    modes, lo = a - hi into TMEM with tcgen05.st. The MMAs then take A from TMEM and only B from shared
//    memory: at M = N = 128 a TF32 MMA with both operands in shared memory needs the full 128 B/clk of
//    shared-memory bandwidth, so moving A out of it is what lets the tensor pipe run.
//  * With `fuse` set the accumulator never leaves the SM before the normalisation: the same threads read
//    x = acc + bias back from TMEM, store (x^2)_hi / (x^2)_lo into the A slots, the MMA thread contracts
//    them with gamma (TMA-staged through the same ring) into a second TMEM accumulator, and the final
//    epilogue writes x / sqrt(norm + beta) (GDN) or x * sqrt(norm + beta) (IGDN).
//
//  TMEM columns: [0,128) accumulator, [128,256) norm accumulator, [256,512) 4 A slots x (32 hi + 32 lo).
constexpr int kStages2 = 4;
constexpr int kUmmaThreads2 = 320;               // warp 0 TMA, warp 1 MMA, warps 2-5 and 6-9: two conversion / epilogue sets
constexpr int kStageBytes2 = 3 * kTileBytes;     // A (raw fp32 from TMA) | B_hi | B_lo
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256;
constexpr uint32_t kTmemCols2 = 512;
constexpr uint32_t kColAcc = 0, kColNrm = 128, kColA = 256;

struct UmmaParams2 {
    int n_taps, kchunks;
    int tile_w, tile_h, tiles_x, tiles_y;
    int Hg, Wg;
    float* out;
    const float* bias;
    const float* beta;       // fused GDN / IGDN
    const float* xin;        // standalone GDN / IGDN: the un-squared input
    int Hout, Wout, out_mul, out_r, out_s, out_split;
    int mode;                // EpilogueMode of a standalone launch
    int fuse;                // 0 none, 1 GDN, 2 IGDN after the contraction
    int exact_main;          // 3xTF32 for the main contraction
    int exact_gdn;           // 3xTF32 for the fused norm (versions 3 and 4)
    int cluster;             // CTAs per cluster (1, 2 or 4): each loads 1/cluster of every B tile and multicasts it
    int n_tiles;             // real tiles; the grid is rounded up to a multiple of `cluster`
    int tile_w_log2;
    int half_da, half_db;    // version 3: offset of the second 128-row half of a tile in the position grid
    int conv1;               // version 3: A rows are the 9x9 patches (k9 s4) of a uint8 image tile staged by TMA
    int debug;               // timing experiments (env EAE_UMMA_DEBUG): 1 no conversion, 2 no MMA, 4 no B loads, 8 no A loads
    long long* times;        // version 3, env EAE_UMMA_TIMING=1: [grid][8] clock64 stamps of the phases of each CTA
    uint32_t* error_flag;
    UmmaTap taps[kMaxTaps];
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// B slice load that lands at the same shared-memory offset (and signals the same mbarrier offset) in every
// CTA of `mask`.
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5, %6}], [%2], %3;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1),
          "r"(c2)
        : "memory");
}
// MMA completion -> the same mbarrier in every CTA of `mask` (stage released cluster-wide).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
                 :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kUmmaThreads2, 1)
gemm_umma2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_g_hi,
                  const __grid_constant__ CUtensorMap map_g_lo, const __grid_constant__ UmmaParams2 p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2);
    uint64_t* full = bars;                      // TMA bytes of the stage landed
    uint64_t* split = bars + kStages2;          // A slot of the stage written to TMEM
    uint64_t* empty = bars + 2 * kStages2;      // MMAs that read the stage (smem B and TMEM A) completed
    uint64_t* acc_full = bars + 3 * kStages2;
    uint64_t* nrm_full = bars + 3 * kStages2 + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages2 + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    // Phantom CTAs that pad the grid to a whole cluster redo the last tile without storing it: they must
    // still load and multicast their share of every B tile.
    const bool store_ok = (int)blockIdx.x < p.n_tiles;
    const int tile = store_ok ? (int)blockIdx.x : p.n_tiles - 1;
    const int img = tile / tiles_per_img;
    const int trem = tile - img * tiles_per_img;
    const uint32_t crank = p.cluster > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
    const int b_rows = kCout / p.cluster;      // rows of every B tile this CTA loads
    const int a0 = (trem / p.tiles_x) * p.tile_h, b0 = (trem % p.tiles_x) * p.tile_w;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages2; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 128);
            mbar_init(&empty[s], (uint32_t)p.cluster);   // one MMA commit per CTA of the cluster
        }
        mbar_init(acc_full, 1);
        mbar_init(nrm_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();     // every CTA's barriers exist before any remote arrive / multicast
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int n_main = p.n_taps * p.kchunks;
    const int n_total = n_main + (p.fuse ? 4 : 0);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < n_total; it++) {
                const int s = it % kStages2;
                if (!mbar_wait(&empty[s], ((it / kStages2) & 1) ^ 1, p.error_flag, 0)) break;
                uint8_t* st = smem + s * kStageBytes2;
                if (it < n_main) {
                    const int t = it / p.kchunks, kc = it - t * p.kchunks;
                    const UmmaTap tap = p.taps[t];
                    const bool ld_a = !(p.debug & 8), ld_b = !(p.debug & 4);      // timing experiments only
                    const uint32_t bytes = (ld_a ? kTileBytes : 0) + (ld_b ? (p.exact_main ? 2 : 1) * kTileBytes : 0);
                    if (bytes) mbar_expect_tx(&full[s], bytes); else mbar_arrive(&full[s]);
                    if (ld_a) tma_load_5d(st, &map_a, &full[s], kc * kChunkK, b0 + tap.fx, a0 + tap.fy, tap.plane, img);
                    const int boff = (int)crank * b_rows * 128;
                    if (!ld_b) {
                    } else if (p.cluster > 1) {
                        tma_load_3d_mc(st + kTileBytes + boff, &map_b_hi, &full[s], kc * kChunkK, (int)crank * b_rows,
                                       tap.w_tap, cmask);
                        if (p.exact_main)
                            tma_load_3d_mc(st + 2 * kTileBytes + boff, &map_b_lo, &full[s], kc * kChunkK,
                                           (int)crank * b_rows, tap.w_tap, cmask);
                    } else {
                        tma_load_3d(st + kTileBytes, &map_b_hi, &full[s], kc * kChunkK, 0, tap.w_tap);
                        if (p.exact_main) tma_load_3d(st + 2 * kTileBytes, &map_b_lo, &full[s], kc * kChunkK, 0, tap.w_tap);
                    }
                } else {
                    const int kc = it - n_main;     // gamma chunk
                    mbar_expect_tx(&full[s], 2 * kTileBytes);
                    const int boff = (int)crank * b_rows * 128;
                    if (p.cluster > 1) {
                        tma_load_3d_mc(st + kTileBytes + boff, &map_g_hi, &full[s], kc * kChunkK, (int)crank * b_rows, 0, cmask);
                        tma_load_3d_mc(st + 2 * kTileBytes + boff, &map_g_lo, &full[s], kc * kChunkK, (int)crank * b_rows, 0,
                                       cmask);
                    } else {
                        tma_load_3d(st + kTileBytes, &map_g_hi, &full[s], kc * kChunkK, 0, 0);
                        tma_load_3d(st + 2 * kTileBytes, &map_g_lo, &full[s], kc * kChunkK, 0, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int it = 0; it < n_total; it++) {
                const int s = it % kStages2;
                if (!mbar_wait(&split[s], (it / kStages2) & 1, p.error_flag, 1)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool gdn = it >= n_main;
                const bool exact = gdn || p.exact_main;
                const uint32_t d = tmem_base + (gdn ? kColNrm : kColAcc);
                const bool first = gdn ? (it == n_main) : (it == 0);
                const uint32_t st = smem_u32(smem + s * kStageBytes2);
                const uint32_t a_slot = tmem_base + kColA + 64u * (uint32_t)s;
                #pragma unroll
                for (int k = 0; k < kChunkK / 8; k++) {
                    if (p.debug & 2) break;
                    const uint64_t b_hi = make_desc(st + kTileBytes + k * 32);
                    umma_tf32_ts(d, a_slot + 8 * k, b_hi, (first && k == 0) ? 0u : 1u);
                    if (exact) {
                        umma_tf32_ts(d, a_slot + 32 + 8 * k, b_hi, 1u);
                        umma_tf32_ts(d, a_slot + 8 * k, make_desc(st + 2 * kTileBytes + k * 32), 1u);
                    }
                }
                if (p.cluster > 1) umma_commit_mc(&empty[s], cmask); else umma_commit(&empty[s]);
                if (it == n_main - 1) umma_commit(acc_full);
                if (gdn && it == n_total - 1) umma_commit(nrm_full);
            }
        }
    } else {
        // ===== warps 2..9: operand conversion into TMEM, then the epilogue =====
        // Two sets of four warps (a warp may only touch TMEM lanes 32 * (warp % 4) ..): set 0 converts the
        // even iterations, set 1 the odd ones, so that two stages are in conversion at any time; in the
        // final epilogue each set writes half of the 128 output channels.
        const int quarter = warp & 3;
        const int set = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        bool ok = true;
        uint32_t r[32];
        for (int it = set; it < n_total && ok; it += 2) {
            const int s = it % kStages2;
            ok = mbar_wait(&full[s], (it / kStages2) & 1, p.error_flag, 2);
            if (!ok) break;
            const bool gdn = it >= n_main;
            if (p.debug & 1) { mbar_arrive(&split[s]); continue; }
            if (!gdn) {
                // this thread's row of the SWIZZLE_128B tile: 16-byte chunk c sits at chunk (c ^ (row & 7))
                const uint8_t* rowp = smem + s * kStageBytes2 + row * 128;
                #pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (row & 7)) << 4));
                    r[4 * c + 0] = __float_as_uint(v.x); r[4 * c + 1] = __float_as_uint(v.y);
                    r[4 * c + 2] = __float_as_uint(v.z); r[4 * c + 3] = __float_as_uint(v.w);
                }
                if (p.mode != kEpiBias) {
                    #pragma unroll
                    for (int i = 0; i < 32; i++) { const float x = __uint_as_float(r[i]); r[i] = __float_as_uint(x * x); }
                }
            } else {
                if (it == n_main || it == n_main + 1) {   // first GDN chunk of this set
                    ok = mbar_wait(acc_full, 0, p.error_flag, 3);
                    if (!ok) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const int c0 = (it - n_main) * kChunkK;
                tmem_ld32(lane_base + kColAcc + c0, r);
                #pragma unroll
                for (int i = 0; i < 32; i++) {
                    float x = __uint_as_float(r[i]);
                    if (p.bias) x += __ldg(p.bias + c0 + i);
                    r[i] = __float_as_uint(x * x);
                }
            }
            const bool exact = gdn || p.exact_main;
            // hi = the fp32 value itself (the tensor core reads only the TF32 bits, i.e. truncates);
            // lo = x - trunc_tf32(x), exact in fp32. One LOP + one FADD per element: cvt.rna here made the
            // conversion, not the MMA, the slowest stage of the ring (scripts/ubench.cu, profiles/).
            tmem_st32(lane_base + kColA + 64u * (uint32_t)s, r);
            if (exact) {
                #pragma unroll
                for (int i = 0; i < 32; i++)
                    r[i] = __float_as_uint(__uint_as_float(r[i]) - __uint_as_float(r[i] & 0xFFFFE000u));
                tmem_st32(lane_base + kColA + 64u * (uint32_t)s + 32u, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&split[s]);
        }
        if (ok) ok = mbar_wait(p.fuse ? nrm_full : acc_full, 0, p.error_flag, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int a = a0 + row / p.tile_w, b = b0 + row % p.tile_w;
        const bool valid = ok && store_ok && a < p.Hg && b < p.Wg;
        size_t opix = 0;
        if (valid) {
            const int oy = a * p.out_mul + p.out_r, ox = b * p.out_mul + p.out_s;
            if (p.out_split)
                opix = (((size_t)img * 4 + (size_t)((oy & 1) * 2 + (ox & 1))) * (p.Hout / 2) + (oy >> 1)) * (p.Wout / 2) + (ox >> 1);
            else
                opix = ((size_t)img * p.Hout + oy) * p.Wout + ox;
        }
        float* o = p.out + opix * kCout;
        const float* xi = p.xin + opix * kCout;
        #pragma unroll 1
        for (int c0 = set * 64; c0 < set * 64 + 64; c0 += 32) {
            uint32_t nr[32];
            tmem_ld32(lane_base + kColAcc + c0, r);
            if (p.fuse) tmem_ld32(lane_base + kColNrm + c0, nr);
            if (valid) {
                #pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                           __uint_as_float(r[j + 3]));
                    if (p.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                    }
                    if (p.fuse) {
                        const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + j));
                        const float n0 = __fsqrt_rn(__uint_as_float(nr[j]) + be.x), n1 = __fsqrt_rn(__uint_as_float(nr[j + 1]) + be.y);
                        const float n2 = __fsqrt_rn(__uint_as_float(nr[j + 2]) + be.z), n3 = __fsqrt_rn(__uint_as_float(nr[j + 3]) + be.w);
                        if (p.fuse == 1) {
                            v.x = __fdiv_rn(v.x, n0); v.y = __fdiv_rn(v.y, n1); v.z = __fdiv_rn(v.z, n2); v.w = __fdiv_rn(v.w, n3);
                        } else {
                            v.x = __fmul_rn(v.x, n0); v.y = __fmul_rn(v.y, n1); v.z = __fmul_rn(v.z, n2); v.w = __fmul_rn(v.w, n3);
                        }
                    } else if (p.mode != kEpiBias) {
                        const float4 x = *reinterpret_cast<const float4*>(xi + c0 + j);
                        if (p.mode == kEpiGdn) {
                            v.x = __fdiv_rn(x.x, __fsqrt_rn(v.x)); v.y = __fdiv_rn(x.y, __fsqrt_rn(v.y));
                            v.z = __fdiv_rn(x.z, __fsqrt_rn(v.z)); v.w = __fdiv_rn(x.w, __fsqrt_rn(v.w));
                        } else {
                            v.x = __fmul_rn(x.x, __fsqrt_rn(v.x)); v.y = __fmul_rn(x.y, __fsqrt_rn(v.y));
                            v.z = __fmul_rn(x.z, __fsqrt_rn(v.z)); v.w = __fmul_rn(x.w, __fsqrt_rn(v.w));
                        }
                    }
                    *reinterpret_cast<float4*>(o + c0 + j) = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
    }
    // No CTA may exit while a peer can still multicast into its shared memory or arrive on its barriers.
    if (p.cluster > 1) cluster_sync_all();
}

}  // namespace
}  // namespace eae
