// Tensor-core path of the tap-list implicit GEMM (conv_plan.cuh) for sm_100a: tcgen05.mma kind::tf32
// with the fp32 accumulator in TMEM, A (activation tile of one tap, 32 channels) and B (weight slice
// of that tap) staged in shared memory by TMA in the canonical K-major SWIZZLE_128B layout, an
// mbarrier ring between the TMA producer warp, the single MMA-issuing thread and the epilogue warps.
//
//   out tile  : 128 positions (tile_w x tile_h of the layer's position grid) x 128 output channels
//   K loop    : taps x (Cin / 32); one stage = one (tap, 32-channel chunk): A 16 KB + B 16 KB
//   A via TMA : 5-D tensor [C, W, H, plane, image]; the tap only shifts the box origin; rows/columns
//               outside the image are zero-filled by TMA, which IS TensorFlow's SAME padding
//   stride 2  : the producing layer writes its output parity-split ([image][y&1][x&1][y/2][x/2][C]),
//               so a stride-2 tap is a dense box of one parity plane (no strided gather)
//   exact3x   : 3xTF32. B is pre-split on the host (hi = rna_tf32(w), lo = rna_tf32(w - hi)); A is
//               split on the fly by the epilogue warps (lo = a - trunc_tf32(a) written next to the
//               TMA tile); D += A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation.
//
// Every wait is bounded (clock64 timeout -> error flag) so that a protocol bug cannot hang the GPU.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "conv_plan.cuh"
#include "transforms.cuh"

namespace eae {

namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                 // fp32 elements per 128-byte swizzle row
constexpr int kTileBytes = kTileM * 128;    // 16 KB: 128 rows x 128 bytes
constexpr int kUmmaThreads = 192;           // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue / A split
constexpr uint32_t kTmemCols = 128;
constexpr long long kTimeoutCycles = 400ll * 1000 * 1000;   // ~0.2 s

template <bool kExact> struct Cfg {
    static constexpr int kStageBytes = kExact ? 4 * kTileBytes : 2 * kTileBytes;   // A [A_lo] B [B_lo]
    static constexpr int kStages = kExact ? 3 : 6;
    static constexpr int kOffAlo = kTileBytes;
    static constexpr int kOffBhi = kExact ? 2 * kTileBytes : kTileBytes;
    static constexpr int kOffBlo = 3 * kTileBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
};

struct UmmaTap { int plane, fy, fx, w_tap; };

struct UmmaParams {
    int n_taps, kchunks;
    int tile_w, tile_h, tiles_x, tiles_y;
    int Hg, Wg;
    float* out;
    const float* bias;
    const float* xin;        // GDN / IGDN: the un-squared input, same flat [M,128] indexing as `out`
    int Hout, Wout, out_mul, out_r, out_s, out_split;
    int mode;                // EpilogueMode
    uint32_t* error_flag;
    UmmaTap taps[kMaxTaps];
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: false (and the error flag set) if the phase does not complete in time.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* error_flag, uint32_t who)
{
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > kTimeoutCycles) {
            atomicOr(error_flag, 1u << who);
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor bit layout): start >> 4 in
// [0,14), LBO in [16,30) (unused here: one swizzle atom along K), SBO = 1024 B (8 rows x 128 B) in
// [32,46), version 1 in [46,48), layout SWIZZLE_128B (2) in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::tf32, D fp32, A/B K-major, M = 128, N = 128 (cute::UMMA::InstrDescriptor bit layout).
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
// One lane of a converged warp (elect.sync): the MMA warp runs its loop with all lanes and issues from the elected
// one, so that every tcgen05 operand is a warp-uniform value (see the note on kTmemBase0).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---- the kernel --------------------------------------------------------------------------------
template <bool kExact>
__global__ void __launch_bounds__(kUmmaThreads, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                 const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ UmmaParams p)
{
    using C = Cfg<kExact>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* full = bars;                       // TMA bytes landed
    uint64_t* split = bars + C::kStages;         // A_lo written (exact mode)
    uint64_t* empty = bars + 2 * C::kStages;     // MMAs that read the stage have completed
    uint64_t* acc_full = bars + 3 * C::kStages;  // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int a0 = (trem / p.tiles_x) * p.tile_h, b0 = (trem % p.tiles_x) * p.tile_w;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 128);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int n_iters = p.n_taps * p.kchunks;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < n_iters; it++) {
                const int s = it % C::kStages;
                if (!mbar_wait(&empty[s], ((it / C::kStages) & 1) ^ 1, p.error_flag, 0)) break;
                const int t = it / p.kchunks, kc = it - t * p.kchunks;
                const UmmaTap tap = p.taps[t];
                uint8_t* st = smem + s * C::kStageBytes;
                mbar_expect_tx(&full[s], kExact ? 3 * kTileBytes : 2 * kTileBytes);
                tma_load_5d(st, &map_a, &full[s], kc * kChunkK, b0 + tap.fx, a0 + tap.fy, tap.plane, img);
                tma_load_3d(st + C::kOffBhi, &map_b_hi, &full[s], kc * kChunkK, 0, tap.w_tap);
                if (kExact) tma_load_3d(st + C::kOffBlo, &map_b_lo, &full[s], kc * kChunkK, 0, tap.w_tap);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            bool ok = true;
            for (int it = 0; it < n_iters && ok; it++) {
                const int s = it % C::kStages;
                ok = mbar_wait((kExact || p.mode != kEpiBias) ? &split[s] : &full[s], (it / C::kStages) & 1,
                               p.error_flag, 1);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + s * C::kStageBytes);
                #pragma unroll
                for (int k = 0; k < kChunkK / 8; k++) {
                    const uint64_t a_hi = make_desc(st + k * 32);
                    const uint64_t b_hi = make_desc(st + C::kOffBhi + k * 32);
                    umma_tf32(tmem_base, a_hi, b_hi, (it | k) ? 1u : 0u);
                    if (kExact) {
                        umma_tf32(tmem_base, make_desc(st + C::kOffAlo + k * 32), b_hi, 1u);
                        umma_tf32(tmem_base, a_hi, make_desc(st + C::kOffBlo + k * 32), 1u);
                    }
                }
                umma_commit(&empty[s]);   // implies tcgen05.fence::before_thread_sync
            }
            umma_commit(acc_full);
        }
    } else {
        // ===== warps 2..5: A split (exact mode), then the epilogue =====
        const int et = threadIdx.x - 64;   // 0..127
        bool ok = true;
        if (kExact || p.mode != kEpiBias) {
            for (int it = 0; it < n_iters && ok; it++) {
                const int s = it % C::kStages;
                ok = mbar_wait(&full[s], (it / C::kStages) & 1, p.error_flag, 2);
                if (!ok) break;
                float4* a = reinterpret_cast<float4*>(smem + s * C::kStageBytes);
                float4* alo = reinterpret_cast<float4*>(smem + s * C::kStageBytes + C::kOffAlo);
                #pragma unroll
                for (int j = 0; j < kTileBytes / 16 / 128; j++) {
                    float4 v = a[et + 128 * j];
                    if (p.mode != kEpiBias) {      // GDN / IGDN contract the squared input
                        v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w;
                        a[et + 128 * j] = v;
                    }
                    if (kExact) {
                        float4 l;
                        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                        alo[et + 128 * j] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor core reads
                mbar_arrive(&split[s]);
            }
        }
        if (ok) ok = mbar_wait(acc_full, 0, p.error_flag, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // TMEM lane quarter of this warp is (warp % 4); accumulator row = TMEM lane.
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int a = a0 + row / p.tile_w, b = b0 + row % p.tile_w;
        const bool valid = ok && a < p.Hg && b < p.Wg;
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
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        #pragma unroll 1
        for (int c0 = 0; c0 < kCout; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            if (valid) {
                #pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 r = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (p.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
                        r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
                    }
                    if (p.mode != kEpiBias) {
                        const float4 x = *reinterpret_cast<const float4*>(xi + c0 + j);
                        if (p.mode == kEpiGdn) {
                            r.x = __fdiv_rn(x.x, __fsqrt_rn(r.x)); r.y = __fdiv_rn(x.y, __fsqrt_rn(r.y));
                            r.z = __fdiv_rn(x.z, __fsqrt_rn(r.z)); r.w = __fdiv_rn(x.w, __fsqrt_rn(r.w));
                        } else {
                            r.x = __fmul_rn(x.x, __fsqrt_rn(r.x)); r.y = __fmul_rn(x.y, __fsqrt_rn(r.y));
                            r.z = __fmul_rn(x.z, __fsqrt_rn(r.z)); r.w = __fmul_rn(x.w, __fsqrt_rn(r.w));
                        }
                    }
                    *reinterpret_cast<float4*>(o + c0 + j) = r;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

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

// =================================================================================================
// Version 3: 256 output positions per CTA (two 128-row halves, two TMEM accumulators).
//
// Measured on version 2 (profiles/): with the loads, the conversion and the MMAs all knocked out, the k5
// convolutions still took half of their time, i.e. the kernel was bound by per-tile fixed cost (TMEM
// allocation, barrier setup, an uncoalesced epilogue) and by the latency of the four-hop mbarrier ring
// with only four 48 KB stages in flight, not by the tensor pipe or by L2. Version 3 therefore
//  * doubles the work per ring iteration and per CTA: one B (weight) stage feeds two accumulators, so the
//    same shared memory holds twice the MMA work per stage and every weight tile is fetched half as often;
//  * stages the epilogue through shared memory and writes whole 512-byte pixel rows per warp instruction;
//  * keeps the norm accumulators of both halves in TMEM during the fused GDN / IGDN: the (x^2)_hi / (x^2)_lo
//    operands of that second contraction are written to shared memory in the canonical swizzled layout.
//
//  TMEM: [0,128) ACC0, [128,256) ACC1, [256,512) two A slots of 128 columns (hi0 lo0 hi1 lo1) during the main
//        loop, then NRM0 [256,384) and NRM1 [384,512).
//  smem: 3 stages x { A0 16K | A1 16K | B_hi 16K | B_lo 16K }.
constexpr int kStages3 = 3;
constexpr int kStageBytes3 = 4 * kTileBytes;
// uint8 image region of 16 x 16 positions of the k9 s4 convolution: rows 4 a0 - 2 .. + 68, columns from 4 b0 - 16
// (TMA needs a 16-byte aligned start in the innermost dimension; the patches start kImgPadX = 14 bytes in).
constexpr int kImgBoxW = 96, kImgBoxH = 69, kImgPadX = 14;
constexpr int kImgBytes = ((kImgBoxW * kImgBoxH + 127) / 128) * 128;
constexpr int kSmemBytes3 = kStages3 * kStageBytes3 + kImgBytes + 1024 + 256;
constexpr int kUmmaThreads3 = 320;
// Registers per thread of versions 3 and 4 (experiment knob): ten warps sit 3 / 3 / 2 / 2 on the four sub-partitions.
#ifndef EAE_MAXREGS34
#define EAE_MAXREGS34 128
#endif
constexpr int kMaxRegs34 = EAE_MAXREGS34;
constexpr uint32_t kCol3Acc0 = 0, kCol3Acc1 = 128, kCol3Slots = 256, kCol3Nrm0 = 256, kCol3Nrm1 = 384;
constexpr uint32_t kTmemBase0 = 0;      // TMEM address of a 512-column allocation on an otherwise empty SM

// 32 consecutive im2col entries (k = ky * 9 + kx, chunk kChunk of three) of one 9x9 uint8 patch as fp32 bit
// patterns: a byte b becomes 0x4B0000bb = 2^23 + b, minus 2^23 (exact).
template <int kChunk>
__device__ __forceinline__ void patch_chunk(const uint8_t* __restrict__ patch, uint32_t* r)
{
    #pragma unroll
    for (int i = 0; i < 32; i++) {
        const int k = kChunk * 32 + i;
        if (k < 81) {
            const int ky = k / 9, kx = k % 9 + (kImgPadX & 3);     // `patch` is the 4-byte aligned address before the patch
            const uint32_t w = *reinterpret_cast<const uint32_t*>(patch + ky * kImgBoxW + (kx & ~3));
            r[i] = __float_as_uint(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 + (kx & 3))) - 8388608.f);
        } else {
            r[i] = 0u;
        }
    }
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// Epilogue staging shared by versions 3 and 4: this thread's row, this set's 64 channels of both halves:
// TMEM (accumulator and, when GDN / IGDN is fused, the norm accumulator) -> bias, normalisation -> shared memory
// (half h at smem + h * stage_bytes, four swizzled [128 x 32] sub-tiles). The TMEM reads of the next 32-column
// chunk are in flight while the current one is processed (TMEM reads run at 64 B/clk per SM and were the longest
// part of the epilogue when every chunk waited for its own load).
__device__ __forceinline__ float rsqrt_fast(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // x = norm + beta >= 2e-5: never denormal
    return r;
}
__device__ __forceinline__ void stage_chunk(uint8_t* sub, int row, int c0, const uint32_t* r, const uint32_t* nr, bool gdn,
                                            int fuse, const float* __restrict__ bias, const float* __restrict__ beta)
{
    #pragma unroll
    for (int c = 0; c < 8; c++) {
        float4 v = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                               __uint_as_float(r[4 * c + 3]));
        if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * c));
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (gdn) {
            // x * rsqrt(norm + beta) (GDN) or x * (n * rsqrt(n)) (IGDN): the 2-ulp MUFU forms; their error (2^-22) is
            // below that of the 3xTF32 contraction that produced x.
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4 * c));
            const float n0 = __uint_as_float(nr[4 * c]) + be.x, n1 = __uint_as_float(nr[4 * c + 1]) + be.y;
            const float n2 = __uint_as_float(nr[4 * c + 2]) + be.z, n3 = __uint_as_float(nr[4 * c + 3]) + be.w;
            if (fuse == 1) {
                v.x *= rsqrt_fast(n0); v.y *= rsqrt_fast(n1); v.z *= rsqrt_fast(n2); v.w *= rsqrt_fast(n3);
            } else {
                v.x *= n0 * rsqrt_fast(n0); v.y *= n1 * rsqrt_fast(n1); v.z *= n2 * rsqrt_fast(n2); v.w *= n3 * rsqrt_fast(n3);
            }
        }
        *reinterpret_cast<float4*>(sub + ((c ^ (row & 7)) << 4)) = v;
    }
}
__device__ __forceinline__ void stage_tile(uint8_t* smem, int stage_bytes, uint32_t lane_base, int set, int row, bool gdn,
                                           int fuse, const float* __restrict__ bias, const float* __restrict__ beta)
{
    uint32_t ra[32], na[32], rb[32], nb[32];
    // chunk q = 2 h + cc covers columns set * 64 + cc * 32 .. + 32 of half h
    tmem_ld32_nowait(lane_base + kCol3Acc0 + set * 64, ra);
    if (gdn) tmem_ld32_nowait(lane_base + kCol3Nrm0 + set * 64, na);
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        const int h = q >> 1, c0 = set * 64 + (q & 1) * 32;
        tmem_ld_wait();
        uint32_t* cur_r = (q & 1) ? rb : ra;
        uint32_t* cur_n = (q & 1) ? nb : na;
        if (q < 3) {
            const int h2 = (q + 1) >> 1, c2 = set * 64 + ((q + 1) & 1) * 32;
            tmem_ld32_nowait(lane_base + (h2 ? kCol3Acc1 : kCol3Acc0) + c2, (q & 1) ? ra : rb);
            if (gdn) tmem_ld32_nowait(lane_base + (h2 ? kCol3Nrm1 : kCol3Nrm0) + c2, (q & 1) ? na : nb);
        }
        stage_chunk(smem + h * stage_bytes + (c0 / 32) * kTileBytes + row * 128, row, c0, cur_r, cur_n, gdn, fuse, bias, beta);
    }
}

// Coalesced store of a staged half (16 x 16 positions per tile, half h = rows 8 h .. 8 h + 7).
struct OutGeom4 {
    float* out;
    int img, a0, b0, Hg, Wg, Hout, Wout, out_mul, out_r, out_s, out_split;
};
__device__ __forceinline__ void store_half4(const OutGeom4& g, const uint8_t* stage, int h, int wq, int lane, bool ok)
{
    #pragma unroll 1
    for (int j0 = 0; j0 < kTileM / 8; j0 += 4) {
        float4 v[4];
        float* dst[4];
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int rr = wq + 8 * (j0 + j);
            const int a = g.a0 + h * 8 + (rr >> 4), b = g.b0 + (rr & 15);
            const int oy = a * g.out_mul + g.out_r, ox = b * g.out_mul + g.out_s;
            size_t opix;
            if (g.out_split)
                opix = (((size_t)g.img * 4 + (size_t)((oy & 1) * 2 + (ox & 1))) * (g.Hout / 2) + (oy >> 1)) * (g.Wout / 2) + (ox >> 1);
            else
                opix = ((size_t)g.img * g.Hout + oy) * g.Wout + ox;
            dst[j] = (ok && a < g.Hg && b < g.Wg) ? g.out + opix * kCout + lane * 4 : nullptr;
            v[j] = *reinterpret_cast<const float4*>(stage + (lane >> 3) * kTileBytes + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
        }
        #pragma unroll
        for (int j = 0; j < 4; j++)
            if (dst[j]) *reinterpret_cast<float4*>(dst[j]) = v[j];
    }
}

// ---- fused GDN / IGDN tail, tensor-memory operand form (versions 3 and 4) ------------------------------------
// The shared-memory-operand GDN MMAs of the first tail ran at ~113 cycles each instead of 64 (A and B both stream
// from shared memory: 128 B/clk, the whole port). Here the A operand ((x^2)_hi | (x^2)_lo of a 32-channel chunk) goes
// to a TMEM slot, as in the main loop, and only gamma streams from shared memory, where it is resident:
//   TMEM   [0,128) ACC0   [128,256) ACC1   [256,384) NRM0   [384,512) two A slots {hi 32 | lo 32}
//          half 1's norm is accumulated in ACC0's columns: each conversion set copies its 64 channels of x_0 = ACC0 + bias
//          to the output staging area right after its two half-0 conversions (acc0_read), before step 4 can overwrite them
//   smem   area + kc * 32K : gamma chunk kc {hi 16K | lo 16K} (loaded once);  area + 128K : staging of half 0;
//          area + 0 : staging of half 1 (after the last MMA)
// Order per conversion set k (steps j = k + 2 i): i = 0, 1 (half 0), copy-out of x_0, i = 2, 3 (half 1), then half 0 is
// normalised in place from NRM0, stored, and half 1 follows after the last MMA.
struct GdnTailTs {
    uint8_t* area;
    uint64_t* g_full;      // [4] gamma chunk kc landed (single use)
    uint64_t* x_ready;     // [4] A slot (set k, sub-slot u) = [2 k + u] written (one arrival per conversion warp of set k)
    uint64_t* x_free;      // [4] the MMAs that read that slot completed
                           //     (3xTF32: one {hi | lo} slot per set, u = 0. Single pass: the 32 lo columns are a second hi
                           //      slot, so a set converts step j + 2 while the tensor pipe still reads step j)
    uint64_t* acc0_read;   // x_0 copied out of TMEM (8 arrivals: every conversion warp)
    uint64_t* acc_full;
    uint64_t* nrm0_full;
    uint64_t* nrm_full;
    int exact;             // 1: 3xTF32 norm (hi / lo squares, hi / lo gamma); 0: single pass, squares rounded to nearest TF32
};
__device__ __forceinline__ void gdn_tail_ts_init(const GdnTailTs& t)
{
    for (int s = 0; s < 4; s++) mbar_init(&t.g_full[s], 1);
    for (int s = 0; s < 4; s++) { mbar_init(&t.x_ready[s], 4); mbar_init(&t.x_free[s], 1); }
    mbar_init(t.acc0_read, 8);
    mbar_init(t.nrm0_full, 1);
}
__device__ __forceinline__ void gdn_tail_ts_producer(const GdnTailTs& t, const CUtensorMap* map_g_hi, const CUtensorMap* map_g_lo,
                                                     uint32_t* error_flag)
{
    if (!mbar_wait(t.acc_full, 0, error_flag, 0)) return;      // gamma lands on the buffers of the main loop
    for (int kc = 0; kc < 4; kc++) {
        uint8_t* g = t.area + kc * 2 * kTileBytes;
        mbar_expect_tx(&t.g_full[kc], (t.exact ? 2 : 1) * kTileBytes);
        tma_load_3d(g, map_g_hi, &t.g_full[kc], kc * kChunkK, 0, 0);
        if (t.exact) tma_load_3d(g + kTileBytes, map_g_lo, &t.g_full[kc], kc * kChunkK, 0, 0);
    }
}
__device__ __forceinline__ void gdn_tail_ts_mma(const GdnTailTs& t, uint32_t* error_flag)
{
    for (int j = 0; j < 8; j++) {
        const int h = j >> 2, kc = j & 3, sl = j & 1, i = j >> 1;
        const int u = t.exact ? 0 : (i & 1);                                 // sub-slot of set sl
        const uint32_t par = t.exact ? (uint32_t)i & 1u : (uint32_t)(i >> 1) & 1u;
        bool ok = mbar_wait(&t.x_ready[2 * sl + u], par, error_flag, 1);
        if (ok) ok = mbar_wait(&t.g_full[kc], 0, error_flag, 1);
        if (ok && j == 4) ok = mbar_wait(t.acc0_read, 0, error_flag, 1);
        if (!__all_sync(0xFFFFFFFFu, ok)) return;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint32_t g = smem_u32(t.area + kc * 2 * kTileBytes);
            const uint32_t d = kTmemBase0 + (h ? kCol3Acc0 : kCol3Nrm0);
            const uint32_t a_hi = kTmemBase0 + kCol3Nrm1 + 64u * (uint32_t)sl + 32u * (uint32_t)u, a_lo = a_hi + 32u;
            #pragma unroll
            for (int k = 0; k < kChunkK / 8; k++) {
                const uint64_t g_hi = make_desc(g + k * 32);
                umma_tf32_ts(d, a_hi + 8 * k, g_hi, (kc == 0 && k == 0) ? 0u : 1u);
                if (t.exact) {
                    umma_tf32_ts(d, a_lo + 8 * k, g_hi, 1u);
                    umma_tf32_ts(d, a_hi + 8 * k, make_desc(g + kTileBytes + k * 32), 1u);
                }
            }
            umma_commit(&t.x_free[2 * sl + u]);
            if (j == 3) umma_commit(t.nrm0_full);
            if (j == 7) umma_commit(t.nrm_full);
        }
        __syncwarp();
    }
}
// Conversion warps of set `set` (thread = accumulator row): conversions, copy-out, normalisation and stores of both halves.
__device__ __forceinline__ bool gdn_tail_ts_run(const GdnTailTs& t, int set, int row, int lane, int wq, uint32_t lane_base,
                                                int fuse, const float* __restrict__ bias, const float* __restrict__ beta,
                                                const OutGeom4& geom, uint32_t* error_flag, long long* stamp)
{
    uint32_t r[32], nr[32];
    uint8_t* stage0 = t.area + 8 * kTileBytes;
    uint8_t* stage1 = t.area;
    bool ok = mbar_wait(t.acc_full, 0, error_flag, 3);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (stamp && threadIdx.x == 64) stamp[4] = clock64();
    #pragma unroll
    for (int i = 0; i < 4 && ok; i++) {
        const int j = set + 2 * i, c0 = (j & 3) * kChunkK;
        const int u = t.exact ? 0 : (i & 1);
        const uint32_t slot = lane_base + kCol3Nrm1 + 64u * (uint32_t)set + 32u * (uint32_t)u;
        tmem_ld32_nowait(lane_base + ((j >> 2) ? kCol3Acc1 : kCol3Acc0) + c0, r);
        // the MMAs that read this slot last: step j - 2 (3xTF32) or step j - 4 (single pass, second use of the sub-slot)
        if (t.exact ? i >= 1 : i >= 2) ok = mbar_wait(&t.x_free[2 * set + u], t.exact ? (uint32_t)(i - 1) & 1u : 0u, error_flag, 7);
        tmem_ld_wait();
        if (!ok) break;
        if (i >= 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        #pragma unroll
        for (int c = 0; c < 8; c++) {
            float4 x = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                                   __uint_as_float(r[4 * c + 3]));
            if (bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * c));
                x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
            }
            r[4 * c] = __float_as_uint(x.x * x.x); r[4 * c + 1] = __float_as_uint(x.y * x.y);
            r[4 * c + 2] = __float_as_uint(x.z * x.z); r[4 * c + 3] = __float_as_uint(x.w * x.w);
        }
        if (!t.exact) {               // single pass: round the squares to the nearest TF32 (see the main loop of version 4)
            #pragma unroll
            for (int q = 0; q < 32; q++) r[q] += 0x1000u;
        }
        tmem_st32(slot, r);           // hi = the value itself (the tensor core truncates), lo = x^2 - trunc_tf32(x^2)
        if (t.exact) {
            #pragma unroll
            for (int q = 0; q < 32; q++) r[q] = __float_as_uint(__uint_as_float(r[q]) - __uint_as_float(r[q] & 0xFFFFE000u));
            tmem_st32(slot + 32u, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&t.x_ready[2 * set + u]);
        if (i == 1) {
            // x_0 = ACC0 + bias of this set's 64 channels -> staging of half 0; ACC0's columns then belong to NRM1
            #pragma unroll
            for (int cc = 0; cc < 2; cc++) {
                const int c1 = set * 64 + cc * 32;
                tmem_ld32(lane_base + kCol3Acc0 + c1, r);
                uint8_t* sub = stage0 + (c1 / 32) * kTileBytes + row * 128;
                #pragma unroll
                for (int c = 0; c < 8; c++) {
                    float4 x = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                                           __uint_as_float(r[4 * c + 3]));
                    if (bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c1 + 4 * c));
                        x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                    }
                    *reinterpret_cast<float4*>(sub + ((c ^ (row & 7)) << 4)) = x;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(t.acc0_read);
        }
    }
    // ---- half 0: normalise the staged x_0 in place with NRM0, store
    if (ok) ok = mbar_wait(t.nrm0_full, 0, error_flag, 4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    #pragma unroll
    for (int cc = 0; cc < 2; cc++) {
        const int c1 = set * 64 + cc * 32;
        tmem_ld32(lane_base + kCol3Nrm0 + c1, nr);
        uint8_t* sub = stage0 + (c1 / 32) * kTileBytes + row * 128;
        #pragma unroll
        for (int c = 0; c < 8; c++) {
            float4* px = reinterpret_cast<float4*>(sub + ((c ^ (row & 7)) << 4));
            float4 x = *px;
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c1 + 4 * c));
            const float n0 = __uint_as_float(nr[4 * c]) + be.x, n1 = __uint_as_float(nr[4 * c + 1]) + be.y;
            const float n2 = __uint_as_float(nr[4 * c + 2]) + be.z, n3 = __uint_as_float(nr[4 * c + 3]) + be.w;
            if (fuse == 1) {
                x.x *= rsqrt_fast(n0); x.y *= rsqrt_fast(n1); x.z *= rsqrt_fast(n2); x.w *= rsqrt_fast(n3);
            } else {
                x.x *= n0 * rsqrt_fast(n0); x.y *= n1 * rsqrt_fast(n1); x.z *= n2 * rsqrt_fast(n2); x.w *= n3 * rsqrt_fast(n3);
            }
            *px = x;
        }
    }
    named_bar_sync(1, 256);     // both sets finished half 0
    store_half4(geom, stage0, 0, wq, lane, ok);
    // ---- half 1: ACC1 and NRM1 (in ACC0's columns) -> staging -> store
    if (ok) ok = mbar_wait(t.nrm_full, 0, error_flag, 4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (stamp && threadIdx.x == 64) stamp[5] = clock64();
    #pragma unroll
    for (int cc = 0; cc < 2; cc++) {
        const int c1 = set * 64 + cc * 32;
        tmem_ld32_nowait(lane_base + kCol3Acc1 + c1, r);
        tmem_ld32_nowait(lane_base + kCol3Acc0 + c1, nr);
        tmem_ld_wait();
        stage_chunk(stage1 + (c1 / 32) * kTileBytes + row * 128, row, c1, r, nr, true, fuse, bias, beta);
    }
    named_bar_sync(1, 256);
    if (stamp && threadIdx.x == 64) stamp[6] = clock64();
    store_half4(geom, stage1, 1, wq, lane, ok);
    return ok;
}

__global__ void __maxnreg__(kMaxRegs34)
gemm_umma3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_g_hi,
                  const __grid_constant__ CUtensorMap map_g_lo, const __grid_constant__ CUtensorMap map_img,
                  const __grid_constant__ UmmaParams2 p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* img_tile = smem + kStages3 * kStageBytes3;      // conv1: [69 rows][80] uint8, rows/cols outside the image are 0
    uint64_t* bars = reinterpret_cast<uint64_t*>(img_tile + kImgBytes);
    uint64_t* full = bars;
    uint64_t* split = bars + kStages3;
    uint64_t* empty = bars + 2 * kStages3;
    uint64_t* acc_full = bars + 3 * kStages3;
    uint64_t* nrm_full = bars + 3 * kStages3 + 1;
    uint64_t* img_full = bars + 3 * kStages3 + 2;
    const GdnTailTs tail{smem, bars + 12 /* g_full[4] */, bars + 16 /* x_ready[4] */, bars + 20 /* x_free[4] */,
                         bars + 24 /* acc0_read */, acc_full, bars + 25 /* nrm0_full */, nrm_full, p.exact_gdn};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* stamp = p.times ? p.times + (size_t)blockIdx.x * 8 : nullptr;
    if (stamp && threadIdx.x == 64) stamp[0] = clock64();
    // tile = tile_w x (2 * tile_h) positions: half h covers rows [a0 + h * tile_h, +tile_h)
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int a0 = (trem / p.tiles_x) * (p.tile_h + p.half_da), b0 = (trem % p.tiles_x) * (p.tile_w + p.half_db);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages3; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 128);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(nrm_full, 1);
        mbar_init(img_full, 1);
        gdn_tail_ts_init(tail);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // This CTA owns the whole TMEM of its SM (512 columns, 1 CTA per SM), so the allocation starts at column 0. The
    // MMA-issuing thread uses that CONSTANT: with the base read from shared memory every tcgen05.mma operand went
    // through an ELECT / R2UR / BRA.U.ANY waterfall (~80 cycles of issue per MMA, more than the 64 it executes).
    if (tmem_base != kTmemBase0 && threadIdx.x == 0) atomicOr(p.error_flag, 1u << 8);
    if (stamp && threadIdx.x == 64) stamp[1] = clock64();

    const int n_main = p.n_taps * p.kchunks;
    const int n_gdn = p.fuse ? 8 : 0;             // (half, gamma chunk) pairs
    // conv1: the A operand is an exact small integer (a pixel), so it has no low part
    const bool a_has_lo = p.exact_main && !p.conv1;
    const int n_total = n_main;                   // the fused GDN steps run in the shared tail (gdn_tail_*), not in this ring

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < n_total; it++) {
                const int s = it % kStages3;
                if (!mbar_wait(&empty[s], ((it / kStages3) & 1) ^ 1, p.error_flag, 0)) break;
                uint8_t* st = smem + s * kStageBytes3;
                if (it < n_main && p.conv1) {
                    if (it == 0) {      // the uint8 image region of this tile (SAME padding = out-of-bounds zero fill)
                        mbar_expect_tx(img_full, kImgBoxW * kImgBoxH);
                        tma_load_3d(img_tile, &map_img, img_full, 4 * b0 - 2 - kImgPadX, 4 * a0 - 2, img);
                    }
                    mbar_expect_tx(&full[s], (p.exact_main ? 2 : 1) * kTileBytes);
                    tma_load_3d(st + 2 * kTileBytes, &map_b_hi, &full[s], it * kChunkK, 0, 0);
                    if (p.exact_main) tma_load_3d(st + 3 * kTileBytes, &map_b_lo, &full[s], it * kChunkK, 0, 0);
                } else if (it < n_main) {
                    const int t = it / p.kchunks, kc = it - t * p.kchunks;
                    const UmmaTap tap = p.taps[t];
                    mbar_expect_tx(&full[s], (p.exact_main ? 4 : 3) * kTileBytes);
                    tma_load_5d(st, &map_a, &full[s], kc * kChunkK, b0 + tap.fx, a0 + tap.fy, tap.plane, img);
                    tma_load_5d(st + kTileBytes, &map_a, &full[s], kc * kChunkK, b0 + p.half_db + tap.fx,
                                a0 + p.half_da + tap.fy, tap.plane, img);
                    tma_load_3d(st + 2 * kTileBytes, &map_b_hi, &full[s], kc * kChunkK, 0, tap.w_tap);
                    if (p.exact_main) tma_load_3d(st + 3 * kTileBytes, &map_b_lo, &full[s], kc * kChunkK, 0, tap.w_tap);
                }
            }
            if (n_gdn) gdn_tail_ts_producer(tail, &map_g_hi, &map_g_lo, p.error_flag);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (warp-uniform operands) =====
        for (int it = 0; it < n_total; it++) {
            const int s = it % kStages3;
            const bool ok = __all_sync(0xFFFFFFFFu, mbar_wait(&split[s], (it / kStages3) & 1, p.error_flag, 1));
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t st = smem_u32(smem + s * kStageBytes3);
                if (it < n_main) {
                    const uint32_t slot = kTmemBase0 + kCol3Slots + 128u * (uint32_t)(it & 1);
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t d = kTmemBase0 + (h ? kCol3Acc1 : kCol3Acc0);
                        const uint32_t a_hi = slot + 64u * (uint32_t)h, a_lo = a_hi + 32u;
                        #pragma unroll
                        for (int k = 0; k < kChunkK / 8; k++) {
                            const uint64_t b_hi = make_desc(st + 2 * kTileBytes + k * 32);
                            umma_tf32_ts(d, a_hi + 8 * k, b_hi, (it == 0 && k == 0) ? 0u : 1u);
                            if (a_has_lo) umma_tf32_ts(d, a_lo + 8 * k, b_hi, 1u);
                            if (p.exact_main) umma_tf32_ts(d, a_hi + 8 * k, make_desc(st + 3 * kTileBytes + k * 32), 1u);
                        }
                    }
                }
                umma_commit(&empty[s]);
                if (it == n_main - 1) { umma_commit(acc_full); if (stamp) stamp[3] = clock64(); }
            }
            __syncwarp();
        }
        if (n_gdn) gdn_tail_ts_mma(tail, p.error_flag);
    } else {
        // ===== warps 2..9: two conversion / epilogue sets (set = iteration parity) =====
        const int quarter = warp & 3;
        const int set = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        bool ok = true;
        uint32_t r[32];
        for (int it = set; it < n_total && ok; it += 2) {
            const int s = it % kStages3;
            ok = mbar_wait(&full[s], (it / kStages3) & 1, p.error_flag, 2);
            if (!ok) break;
            if (stamp && it == 0 && threadIdx.x == 64) stamp[2] = clock64();
            uint8_t* st = smem + s * kStageBytes3;
            if (it < n_main) {
                // TMEM slot (it & 1) was last read by the MMAs of iteration it - 2
                if (it >= 2) {
                    ok = mbar_wait(&empty[(it - 2) % kStages3], ((it - 2) / kStages3) & 1, p.error_flag, 5);
                    if (!ok) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t slot = lane_base + kCol3Slots + 128u * (uint32_t)(it & 1);
                if (p.conv1) {
                    // A rows straight from the pixels: row (a, b) of half h is the 9x9 patch whose top-left pixel is
                    // tile byte (4 (8 h + a), 4 b); chunk `it` covers k = ky * 9 + kx in [32 it, 32 it + 32), k >= 81 is 0.
                    if (it == 0 || it == 1) {
                        ok = mbar_wait(img_full, 0, p.error_flag, 6);
                        if (!ok) break;
                    }
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint8_t* patch = img_tile + (4 * (8 * h + (row >> 4))) * kImgBoxW + (kImgPadX & ~3) + 4 * (row & 15);
                        if (it == 0) patch_chunk<0>(patch, r);
                        else if (it == 1) patch_chunk<1>(patch, r);
                        else patch_chunk<2>(patch, r);
                        tmem_st32(slot + 64u * (uint32_t)h, r);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&split[s]);
                    continue;
                }
                #pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint8_t* rowp = st + h * kTileBytes + row * 128;
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
                    // hi = raw fp32 (the tensor core truncates to TF32), lo = x - trunc_tf32(x)
                    tmem_st32(slot + 64u * (uint32_t)h, r);
                    if (p.exact_main) {
                        #pragma unroll
                        for (int i = 0; i < 32; i++)
                            r[i] = __float_as_uint(__uint_as_float(r[i]) - __uint_as_float(r[i] & 0xFFFFE000u));
                        tmem_st32(slot + 64u * (uint32_t)h + 32u, r);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            mbar_arrive(&split[s]);
        }
        if (n_gdn) {
            // ---- fused GDN / IGDN (tile geometry 16 x 16, as in version 4)
            const int wq = warp - 2;
            const OutGeom4 geom{p.out, img, a0, b0, p.Hg, p.Wg, p.Hout, p.Wout, p.out_mul, p.out_r, p.out_s, p.out_split};
            if (ok) ok = gdn_tail_ts_run(tail, set, row, lane, wq, lane_base, p.fuse, p.bias, p.beta, geom, p.error_flag, stamp);
        } else {
        if (ok) ok = mbar_wait(acc_full, 0, p.error_flag, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (stamp && threadIdx.x == 64) stamp[5] = clock64();

        // ---- epilogue: this set's 64 channels of both halves -> shared-memory staging (stage h of the ring,
        // four swizzled [128 x 32] sub-tiles) -> coalesced 512-byte rows.
        stage_tile(smem, kStageBytes3, lane_base, set, row, false, 0, p.bias, p.beta);
        named_bar_sync(1, 256);     // both sets finished staging
        if (stamp && threadIdx.x == 64) stamp[6] = clock64();
        // Coalesced stores: warp wq writes rows wq, wq + 8, ... of each half, one 512-byte pixel per instruction;
        // four rows are in flight at a time.
        const int wq = warp - 2;
        const bool fixup = !n_gdn && p.mode != kEpiBias;    // standalone GDN / IGDN
        #pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const uint8_t* stage = smem + h * kStageBytes3;
            #pragma unroll 1
            for (int j0 = 0; j0 < kTileM / 8; j0 += 4) {
                float4 v[4];
                float* dst[4];
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int rr = wq + 8 * (j0 + j);
                    const int a = a0 + h * p.half_da + (rr >> p.tile_w_log2), b = b0 + h * p.half_db + (rr & (p.tile_w - 1));
                    const int oy = a * p.out_mul + p.out_r, ox = b * p.out_mul + p.out_s;
                    size_t opix;
                    if (p.out_split)
                        opix = (((size_t)img * 4 + (size_t)((oy & 1) * 2 + (ox & 1))) * (p.Hout / 2) + (oy >> 1)) * (p.Wout / 2) + (ox >> 1);
                    else
                        opix = ((size_t)img * p.Hout + oy) * p.Wout + ox;
                    dst[j] = (ok && a < p.Hg && b < p.Wg) ? p.out + opix * kCout + lane * 4 : nullptr;
                    v[j] = *reinterpret_cast<const float4*>(stage + (lane >> 3) * kTileBytes + rr * 128 +
                                                            (((lane & 7) ^ (rr & 7)) << 4));
                }
                if (fixup) {
                    // v holds norm (+ beta via bias); combine with the un-squared input
                    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (!dst[j]) continue;
                        const float4 x = *reinterpret_cast<const float4*>(p.xin + (dst[j] - p.out));
                        if (p.mode == kEpiGdn) {
                            v[j].x = __fdiv_rn(x.x, __fsqrt_rn(v[j].x)); v[j].y = __fdiv_rn(x.y, __fsqrt_rn(v[j].y));
                            v[j].z = __fdiv_rn(x.z, __fsqrt_rn(v[j].z)); v[j].w = __fdiv_rn(x.w, __fsqrt_rn(v[j].w));
                        } else {
                            v[j].x = __fmul_rn(x.x, __fsqrt_rn(v[j].x)); v[j].y = __fmul_rn(x.y, __fsqrt_rn(v[j].y));
                            v[j].z = __fmul_rn(x.z, __fsqrt_rn(v[j].z)); v[j].w = __fmul_rn(x.w, __fsqrt_rn(v[j].w));
                        }
                    }
                }
                #pragma unroll
                for (int j = 0; j < 4; j++)
                    if (dst[j]) *reinterpret_cast<float4*>(dst[j]) = v[j];
            }
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // (a clock read right after a barrier gives the time this warp ISSUED the barrier, not its release)
    if (stamp && threadIdx.x == 64) stamp[7] = clock64();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
    }
}

// =================================================================================================
// Version 4 (default for the multi-tap layers): version 3 with every activation box fetched ONCE.
//
// Measured on version 3 (profiles/r01_ncu_full_gemm_umma3_layers.md): a ring iteration of a k5 layer moves
// 64 KB from L2 into shared memory (two 16 KB activation boxes + 32 KB of split weights) for 24 MMAs, i.e.
// the kernel asks for ~42 B/clk/SM against an L2 throughput cap of ~43 B/clk/SM (6.3 KB/clk over 148 SMs),
// and the conversion warps spend 38 % of their samples waiting for TMA data: iterations take 1.95 k cycles
// instead of the 1.54 k the MMAs need. But the boxes of the taps of one input plane are the same pixels shifted
// by one position: the 9 taps of the (odd, odd) parity plane of a k5 s2 convolution overlap in 15/16 of their
// rows. Here the loop runs channel chunk -> tap group -> tap, the UNION box of a group (18 x 18 positions x 32
// channels, 41 KB) is loaded once into one of two buffers, and the conversion warps read each tap's rows from
// it at a shifted offset; only the weights stream per tap (32 KB stages, 4 deep). Activation traffic drops from
// 3.2 MB to 0.65 MB per 256-position tile of a 25-tap layer (total L2 -> SM traffic -40 %).
//
//  smem: union buffers 2 x 41 KB | weight stages 4 x { B_hi 16K | B_lo 16K } | barriers. The fused GDN phase and
//        the epilogue alias the first 192 KB as in version 3 (3 stages x 64 KB), after the main loop has drained.
constexpr int kUnionW = 18, kUnionH = 18;
constexpr int kUnionTx = kUnionW * kUnionH * 128;          // bytes one union load delivers
constexpr int kUnionBytes = 41 * 1024;
constexpr int kBStages4 = 4, kBStageBytes4 = 2 * kTileBytes;
constexpr int kOffB4 = 2 * kUnionBytes;
constexpr int kOffBars4 = kOffB4 + kBStages4 * kBStageBytes4;
constexpr int kSmemBytes4 = kOffBars4 + 512 + 1024;
constexpr int kGdnStageBytes4 = 4 * kTileBytes;
static_assert(3 * kGdnStageBytes4 <= kOffBars4, "GDN / epilogue stages must fit below the barriers");
constexpr int kMaxGroups4 = 4;

struct UmmaTap4 { int w_tap, off, grp, last; };            // off: row offset of this tap's box inside its group's union
struct UmmaGroup4 { int plane, fy, fx, pad; };             // union origin relative to the tile origin
struct UmmaParams4 {
    int n_taps, kchunks, n_groups;
    int tiles_x, tiles_y, Hg, Wg;
    float* out;
    const float* bias;
    const float* beta;
    int Hout, Wout, out_mul, out_r, out_s, out_split;
    int fuse, exact_main, exact_gdn;
    long long* times;
    uint32_t* error_flag;
    UmmaTap4 taps[kMaxTaps];
    UmmaGroup4 groups[kMaxGroups4];
};

// Up to four launches that read the same input through the same weight array (the four output phases of a transposed
// convolution) run as ONE grid: CTA b works on tile b / n_phases of phase b % n_phases. The phases of a tile are neighbours
// in the grid, so the input box they share is fetched from HBM once; and six dependent launches per step disappear
// (beside other streams' kernels a dependent launch waits ~10 us).
struct UmmaParams4x {
    int n_phases, pad;
    UmmaParams4 ph[4];
};

__global__ void __maxnreg__(kMaxRegs34)
gemm_umma4_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_g_hi,
                  const __grid_constant__ CUtensorMap map_g_lo, const __grid_constant__ UmmaParams4x pp)
{
    const UmmaParams4& p = pp.ph[blockIdx.x % (unsigned)pp.n_phases];
    const int tile_linear = (int)(blockIdx.x / (unsigned)pp.n_phases);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars4);
    uint64_t* b_full = bars;               // [4] weight stage landed
    uint64_t* done = bars + 4;             // [4] the MMAs of iteration it (it & 3) completed: ONE commit per iteration
                                           //     releases the weight stage (it + 4), the TMEM A slot (it + 2) and, after the
                                           //     last tap of a group, its union buffer (a tcgen05.commit costs ~100 cycles of
                                           //     tensor-pipe time, three per iteration made the loop 15 % slower)
    uint64_t* u_full = bars + 8;           // [2] union box landed
    uint64_t* split = bars + 12;           // [4] TMEM A slot of iteration it (it & 3) written (one arrival per conversion warp)
    uint64_t* acc_full = bars + 16;
    uint64_t* nrm_full = bars + 17;
    const GdnTailTs tail{smem, bars + 18 /* g_full[4] */, bars + 22 /* x_ready[4] */, bars + 26 /* x_free[4] */,
                         bars + 30 /* acc0_read */, acc_full, bars + 31 /* nrm0_full */, nrm_full, p.exact_gdn};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* stamp = p.times ? p.times + (size_t)blockIdx.x * 12 : nullptr;      // [8] start ns, [9] end ns, [10] SM id
    if (stamp && threadIdx.x == 64) {
        stamp[0] = clock64();
        uint32_t smid;
        long long t;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        stamp[8] = t; stamp[10] = smid;
    }
    // tile = 16 x 16 positions: half h covers rows [a0 + 8 h, + 8)
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int img = tile_linear / tiles_per_img;
    const int trem = tile_linear - img * tiles_per_img;
    const int a0 = (trem / p.tiles_x) * 16, b0 = (trem % p.tiles_x) * 16;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 4; s++) { mbar_init(&b_full[s], 1); mbar_init(&done[s], 1); }
        for (int s = 0; s < 2; s++) mbar_init(&u_full[s], 1);
        for (int s = 0; s < 4; s++) mbar_init(&split[s], 4);      // one arrival per conversion warp
        gdn_tail_ts_init(tail);
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
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // This CTA owns the whole TMEM of its SM (512 columns, 1 CTA per SM), so the allocation starts at column 0. The
    // MMA-issuing thread uses that CONSTANT: with the base read from shared memory every tcgen05.mma operand went
    // through an ELECT / R2UR / BRA.U.ANY waterfall (~80 cycles of issue per MMA, more than the 64 it executes).
    if (tmem_base != kTmemBase0 && threadIdx.x == 0) atomicOr(p.error_flag, 1u << 8);
    if (stamp && threadIdx.x == 64) stamp[1] = clock64();

    const int n_main = p.n_taps * p.kchunks;      // iteration it = kc * n_taps + t
    const int n_gdn = p.fuse ? 8 : 0;             // (half, gamma chunk) pairs
    const int n_unions = p.kchunks * p.n_groups;  // union g = kc * n_groups + group

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            bool ok = true;
            int issued = 0;                       // unions requested so far
            for (int it = 0; it < n_main && ok; it++) {
                const int kc = it / p.n_taps, t = it - kc * p.n_taps;
                const UmmaTap4 tap = p.taps[t];
                const int g = kc * p.n_groups + tap.grp;
                // the union this tap reads, and (without waiting) the one after it as soon as its buffer is free
                while (ok && issued < n_unions && issued <= g + 1) {
                    const int buf = issued & 1;
                    if (issued >= 2) {
                        // the buffer held union issued - 2: free once the MMAs of that group's last tap are done
                        const int pk = (issued - 2) / p.n_groups, pg = (issued - 2) - pk * p.n_groups;
                        const int last_it = pk * p.n_taps + p.groups[pg].pad;      // pad = index of the group's last tap
                        const uint32_t par = (uint32_t)(last_it >> 2) & 1u;
                        if (issued <= g) ok = mbar_wait(&done[last_it & 3], par, p.error_flag, 0);
                        else if (!mbar_try(&done[last_it & 3], par)) break;
                        if (!ok) break;
                    }
                    const UmmaGroup4 grp = p.groups[issued % p.n_groups];
                    mbar_expect_tx(&u_full[buf], kUnionTx);
                    tma_load_5d(smem + buf * kUnionBytes, &map_u, &u_full[buf], (issued / p.n_groups) * kChunkK, b0 + grp.fx,
                                a0 + grp.fy, grp.plane, img);
                    issued++;
                }
                if (!ok) break;
                const int s = it & 3;
                if (!mbar_wait(&done[s], ((uint32_t)(it >> 2) & 1u) ^ 1u, p.error_flag, 0)) { ok = false; break; }
                uint8_t* st = smem + kOffB4 + s * kBStageBytes4;
                mbar_expect_tx(&b_full[s], (p.exact_main ? 2 : 1) * kTileBytes);
                tma_load_3d(st, &map_b_hi, &b_full[s], kc * kChunkK, 0, tap.w_tap);
                if (p.exact_main) tma_load_3d(st + kTileBytes, &map_b_lo, &b_full[s], kc * kChunkK, 0, tap.w_tap);
            }
            if (ok && n_gdn) gdn_tail_ts_producer(tail, &map_g_hi, &map_g_lo, p.error_flag);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        {
            bool ok = true;
            for (int it = 0; it < n_main && ok; it++) {
                const int slot_i = it & 1, s = it & 3;
                ok = mbar_wait(&split[s], (uint32_t)(it >> 2) & 1u, p.error_flag, 1);
                if (ok) ok = mbar_wait(&b_full[s], (uint32_t)(it >> 2) & 1u, p.error_flag, 1);
                ok = __all_sync(0xFFFFFFFFu, ok);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t st = smem_u32(smem + kOffB4 + s * kBStageBytes4);
                    // single pass: the lo columns of a set's slot are a second hi slot (iterations it, it + 2 of the set)
                    const uint32_t slot = kTmemBase0 + kCol3Slots + 128u * (uint32_t)slot_i +
                                          (p.exact_main ? 0u : 32u * (uint32_t)((it >> 1) & 1));
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t d = kTmemBase0 + (h ? kCol3Acc1 : kCol3Acc0);
                        const uint32_t a_hi = slot + 64u * (uint32_t)h, a_lo = a_hi + 32u;
                        #pragma unroll
                        for (int k = 0; k < kChunkK / 8; k++) {
                            const uint64_t b_hi = make_desc(st + k * 32);
                            umma_tf32_ts(d, a_hi + 8 * k, b_hi, (it == 0 && k == 0) ? 0u : 1u);
                            if (p.exact_main) {
                                umma_tf32_ts(d, a_lo + 8 * k, b_hi, 1u);
                                umma_tf32_ts(d, a_hi + 8 * k, make_desc(st + kTileBytes + k * 32), 1u);
                            }
                        }
                    }
                    umma_commit(&done[s]);
                    if (it == n_main - 1) { umma_commit(acc_full); if (stamp) stamp[3] = clock64(); }
                }
                __syncwarp();
            }
            if (ok && n_gdn) gdn_tail_ts_mma(tail, p.error_flag);
        }
    } else {
        // ===== warps 2..9: two conversion / epilogue sets; set k owns TMEM A slot k and the iterations of parity k
        const int quarter = warp & 3;
        const int set = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t slot_set = lane_base + kCol3Slots + 128u * (uint32_t)set;
        const int reuse = p.exact_main ? 2 : 4;      // the slot written now was read by the MMAs of iteration it - reuse
        const int row_in_union = (row >> 4) * kUnionW + (row & 15);     // half 1 adds 8 union rows
        bool ok = true;
        uint32_t r[32], hi[32];
        for (int it = set; it < n_main && ok; it += 2) {
            const int kc = it / p.n_taps, t = it - kc * p.n_taps;
            const UmmaTap4 tap = p.taps[t];
            const int g = kc * p.n_groups + tap.grp;
            ok = mbar_wait(&u_full[g & 1], (uint32_t)(g >> 1) & 1u, p.error_flag, 2);
            if (!ok) break;
            if (stamp && it == 0 && threadIdx.x == 64) stamp[2] = clock64();
            // Read both halves' rows first: the shared-memory reads do not depend on the TMEM slot, so they overlap the
            // wait for the MMAs of iteration it - 2 (the completion -> conversion -> issue chain paces the loop).
            const uint8_t* ubuf = smem + (g & 1) * kUnionBytes;
            #pragma unroll
            for (int h = 0; h < 2; h++) {
                const int ur = row_in_union + h * 8 * kUnionW + tap.off;
                const uint8_t* rowp = ubuf + ur * 128;
                uint32_t* dst = h ? hi : r;
                #pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (ur & 7)) << 4));
                    dst[4 * c + 0] = __float_as_uint(v.x); dst[4 * c + 1] = __float_as_uint(v.y);
                    dst[4 * c + 2] = __float_as_uint(v.z); dst[4 * c + 3] = __float_as_uint(v.w);
                }
            }
            const uint32_t slot = slot_set + (p.exact_main ? 0u : 32u * (uint32_t)((it >> 1) & 1));
            if (it >= reuse) {
                ok = mbar_wait(&done[(it - reuse) & 3], (uint32_t)((it - reuse) >> 2) & 1u, p.error_flag, 5);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            // hi = raw fp32 (the tensor core truncates to TF32), lo = x - trunc_tf32(x). Single pass: round to nearest
            // instead (add half a TF32 ulp to the magnitude before the truncation) - truncation shrinks every product by
            // 2^-12 on average, a bias that does not average out over the ~1 200 terms of a sum.
            if (!p.exact_main) {
                #pragma unroll
                for (int i = 0; i < 32; i++) { r[i] += 0x1000u; hi[i] += 0x1000u; }
            }
            tmem_st32(slot, r);
            tmem_st32(slot + 64u, hi);
            if (p.exact_main) {
                #pragma unroll
                for (int i = 0; i < 32; i++) {
                    r[i] = __float_as_uint(__uint_as_float(r[i]) - __uint_as_float(r[i] & 0xFFFFE000u));
                    hi[i] = __float_as_uint(__uint_as_float(hi[i]) - __uint_as_float(hi[i] & 0xFFFFE000u));
                }
                tmem_st32(slot + 32u, r);
                tmem_st32(slot + 96u, hi);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[it & 3]);   // 4 arrivals instead of 128: the arrive chain is on the critical path
        }
        const int wq = warp - 2;
        const OutGeom4 geom{p.out, img, a0, b0, p.Hg, p.Wg, p.Hout, p.Wout, p.out_mul, p.out_r, p.out_s, p.out_split};
        uint8_t* stage0 = smem;                          // un-fused epilogue: both halves staged side by side
        uint8_t* stage1 = smem + kGdnStageBytes4;
        if (ok && n_gdn) {
            ok = gdn_tail_ts_run(tail, set, row, lane, wq, lane_base, p.fuse, p.bias, p.beta, geom, p.error_flag, stamp);
        } else {
            if (ok) ok = mbar_wait(acc_full, 0, p.error_flag, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (stamp && threadIdx.x == 64) stamp[5] = clock64();
            stage_tile(smem, kGdnStageBytes4, lane_base, set, row, false, 0, p.bias, p.beta);
            named_bar_sync(1, 256);     // both sets finished staging
            if (stamp && threadIdx.x == 64) stamp[6] = clock64();
            store_half4(geom, stage0, 0, wq, lane, ok);
            store_half4(geom, stage1, 1, wq, lane, ok);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (stamp && threadIdx.x == 64) {
        stamp[7] = clock64();
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        stamp[9] = t;
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
    }
}

// =================================================================================================
// Version 5: the 1-tap ("thin") layers, 128 positions per CTA, TWO CTAs per SM.
//
// Layer 1 (k9 s4 from one channel, K = 96) and the last layer (k9 s4 to one channel, K = 128) have 589 824 output rows
// per 24 images and almost no contraction: measured on versions 3 / 4, a 256-row tile of layer 1 spends 5 k cycles in
// its main loop and 18 k in its tail (fused GDN with shared-memory operands, TMEM read-out at 64 B/clk, stores), during
// which the tensor pipe, the TMEM port or the store path sit idle in turn. With half-size CTAs (128 rows, 256 TMEM
// columns, 104 KB of shared memory) two CTAs are resident per SM and one's tail overlaps the other's loads, MMAs and
// stores; their weights are tiny (96 or 128 KB per layer, L2-resident), so the smaller tile costs no L2 bandwidth.
//
//  192 threads: warp 0 TMA, warp 1 MMA (converged, elected lane issues), warps 2-5 conversion / epilogue (thread = row).
//  TMEM (256 columns from the allocator's base): ACC [0,128) | two A slots {hi 32 | lo 32} at [128,256), the norm
//        accumulator of a fused GDN takes their place afterwards.
//  smem: 96 KB of stages | uint8 image tile (layer 1) | barriers.
//        layer 1  : its three weight chunks {hi 16K | lo 16K} at kc * 32K, all requested up front
//        otherwise: 2 stages x {A 16K | B_hi 16K | B_lo 16K}
//        tail     : ONE GDN stage {(x^2)_hi | (x^2)_lo | gamma_hi | gamma_lo} (the other CTA of the SM fills the bubbles),
//                   then the 64 KB output staging, both at offset 0.
constexpr int kThreads5 = 192;
constexpr int kMainBytes5 = 96 * 1024;
constexpr int kSmemBytes5 = kMainBytes5 + kImgBytes + 256 + 1024;
constexpr uint32_t kTmemCols5 = 256;
constexpr uint32_t kCol5Slots = 128, kCol5Nrm = 128;

struct UmmaParams5 {
    int kchunks;             // 3 (layer 1) or Cin / 32
    int conv1;               // A rows gathered from the uint8 image
    int tiles_x, tiles_y, Hg, Wg;
    float* out;
    const float* bias;
    const float* beta;
    int Hout, Wout, out_mul, out_r, out_s, out_split;
    int fuse, exact_main;
    long long* times;        // debug (EAE_UMMA_TIMING): [grid][3] = SM id, globaltimer at CTA start / end
    uint32_t* error_flag;
};

// (max-threads 256 in the launch bounds caps the kernel at 128 registers: the 12 warps of two CTAs can land four to a
// sub-partition, whose register file holds 16 K registers)
__global__ void __launch_bounds__(256, 2)
gemm_umma5_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                  const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_g_hi,
                  const __grid_constant__ CUtensorMap map_g_lo, const __grid_constant__ CUtensorMap map_img,
                  const __grid_constant__ UmmaParams5 p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint8_t* img_tile = smem + kMainBytes5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(img_tile + kImgBytes);
    uint64_t* full = bars;                 // [3] stage landed
    uint64_t* done = bars + 3;             // [3] MMAs of the iteration that used the stage completed
    uint64_t* split = bars + 6;            // [2] TMEM A slot written (one arrival per conversion warp)
    uint64_t* acc_full = bars + 8;
    uint64_t* img_full = bars + 9;
    uint64_t* g_full = bars + 10;          // gamma chunk of a GDN step landed
    uint64_t* x_ready = bars + 11;         // x^2 of the step written
    uint64_t* x_free = bars + 12;          // the MMAs of the step completed
    uint64_t* nrm_full = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.times && threadIdx.x == 0) {
        uint32_t smid;
        long long t;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.times[(size_t)blockIdx.x * 3] = smid;
        p.times[(size_t)blockIdx.x * 3 + 1] = t;
    }
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int a0 = (trem / p.tiles_x) * 8, b0 = (trem % p.tiles_x) * 16;      // tile = 16 x 8 positions
    const int n_stage = p.conv1 ? 3 : 2;
    const int stage_bytes = p.conv1 ? 2 * kTileBytes : 3 * kTileBytes;
    const int b_off = p.conv1 ? 0 : kTileBytes;
    const int n_main = p.kchunks;
    const int n_gdn = p.fuse ? 4 : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 3; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); }
        for (int s = 0; s < 2; s++) mbar_init(&split[s], 4);
        mbar_init(acc_full, 1); mbar_init(img_full, 1); mbar_init(g_full, 1); mbar_init(x_ready, 4); mbar_init(x_free, 1);
        mbar_init(nrm_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols5) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;      // 0 or 256: two CTAs share the SM

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            bool ok = true;
            if (p.conv1) {
                mbar_expect_tx(img_full, kImgBoxW * kImgBoxH);
                tma_load_3d(img_tile, &map_img, img_full, 4 * b0 - 2 - kImgPadX, 4 * a0 - 2, img);
            }
            for (int it = 0; it < n_main && ok; it++) {
                const int s = it % n_stage;
                if (it >= n_stage) ok = mbar_wait(&done[s], (uint32_t)(it / n_stage - 1) & 1u, p.error_flag, 0);
                if (!ok) break;
                uint8_t* st = smem + s * stage_bytes;
                mbar_expect_tx(&full[s], ((p.conv1 ? 0 : 1) + (p.exact_main ? 2 : 1)) * kTileBytes);
                if (!p.conv1) tma_load_5d(st, &map_a, &full[s], it * kChunkK, b0, a0, 0, img);
                tma_load_3d(st + b_off, &map_b_hi, &full[s], it * kChunkK, 0, 0);
                if (p.exact_main) tma_load_3d(st + b_off + kTileBytes, &map_b_lo, &full[s], it * kChunkK, 0, 0);
            }
            if (ok && n_gdn) {
                ok = mbar_wait(acc_full, 0, p.error_flag, 0);      // the GDN stage aliases the main stages
                for (int j = 0; j < n_gdn && ok; j++) {
                    if (j >= 1) ok = mbar_wait(x_free, (uint32_t)(j - 1) & 1u, p.error_flag, 0);
                    if (!ok) break;
                    mbar_expect_tx(g_full, 2 * kTileBytes);
                    tma_load_3d(smem + 2 * kTileBytes, &map_g_hi, g_full, j * kChunkK, 0, 0);
                    tma_load_3d(smem + 3 * kTileBytes, &map_g_lo, g_full, j * kChunkK, 0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        const uint32_t tb = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        const bool a_has_lo = p.exact_main && !p.conv1;      // a pixel is exact in TF32
        bool ok = true;
        for (int it = 0; it < n_main && ok; it++) {
            const int s = it % n_stage, slot_i = it & 1;
            ok = __all_sync(0xFFFFFFFFu, mbar_wait(&split[slot_i], (uint32_t)(it >> 1) & 1u, p.error_flag, 1));
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t st = smem_u32(smem + s * stage_bytes + b_off);
                const uint32_t a_hi = tb + kCol5Slots + 64u * (uint32_t)slot_i, a_lo = a_hi + 32u;
                #pragma unroll
                for (int k = 0; k < kChunkK / 8; k++) {
                    const uint64_t b_hi = make_desc(st + k * 32);
                    umma_tf32_ts(tb, a_hi + 8 * k, b_hi, (it == 0 && k == 0) ? 0u : 1u);
                    if (a_has_lo) umma_tf32_ts(tb, a_lo + 8 * k, b_hi, 1u);
                    if (p.exact_main) umma_tf32_ts(tb, a_hi + 8 * k, make_desc(st + kTileBytes + k * 32), 1u);
                }
                umma_commit(&done[s]);
                if (it == n_main - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
        for (int j = 0; j < n_gdn && ok; j++) {
            ok = mbar_wait(x_ready, (uint32_t)j & 1u, p.error_flag, 1);
            if (ok) ok = mbar_wait(g_full, (uint32_t)j & 1u, p.error_flag, 1);
            ok = __all_sync(0xFFFFFFFFu, ok);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t st = smem_u32(smem);
                #pragma unroll
                for (int k = 0; k < kChunkK / 8; k++) {
                    const uint64_t x_hi = make_desc(st + k * 32), x_lo = make_desc(st + kTileBytes + k * 32);
                    const uint64_t g_hi = make_desc(st + 2 * kTileBytes + k * 32);
                    umma_tf32(tb + kCol5Nrm, x_hi, g_hi, (j == 0 && k == 0) ? 0u : 1u);
                    umma_tf32(tb + kCol5Nrm, x_lo, g_hi, 1u);
                    umma_tf32(tb + kCol5Nrm, x_hi, make_desc(st + 3 * kTileBytes + k * 32), 1u);
                }
                umma_commit(x_free);
                if (j == n_gdn - 1) umma_commit(nrm_full);
            }
            __syncwarp();
        }
    } else {
        // ===== warps 2..5: operand conversion, GDN squares, epilogue (thread = accumulator row = TMEM lane) =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        bool ok = true;
        uint32_t r[32], nr[32];
        for (int it = 0; it < n_main && ok; it++) {
            const int s = it % n_stage;
            ok = mbar_wait(&full[s], (uint32_t)(it / n_stage) & 1u, p.error_flag, 2);
            if (!ok) break;
            if (p.conv1) {
                if (it == 0) { ok = mbar_wait(img_full, 0, p.error_flag, 6); if (!ok) break; }
                const uint8_t* patch = img_tile + (4 * (row >> 4)) * kImgBoxW + (kImgPadX & ~3) + 4 * (row & 15);
                if (it == 0) patch_chunk<0>(patch, r);
                else if (it == 1) patch_chunk<1>(patch, r);
                else patch_chunk<2>(patch, r);
            } else {
                const uint8_t* rowp = smem + s * stage_bytes + row * 128;
                #pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (row & 7)) << 4));
                    r[4 * c + 0] = __float_as_uint(v.x); r[4 * c + 1] = __float_as_uint(v.y);
                    r[4 * c + 2] = __float_as_uint(v.z); r[4 * c + 3] = __float_as_uint(v.w);
                }
            }
            if (it >= 2) {      // the MMAs of iteration it - 2 read this TMEM slot
                const int s2 = (it - 2) % n_stage;
                ok = mbar_wait(&done[s2], (uint32_t)((it - 2) / n_stage) & 1u, p.error_flag, 5);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t slot = lane_base + kCol5Slots + 64u * (uint32_t)(it & 1);
            if (!p.exact_main && !p.conv1) {      // single pass: round to nearest TF32 (see version 4)
                #pragma unroll
                for (int i = 0; i < 32; i++) r[i] += 0x1000u;
            }
            tmem_st32(slot, r);
            if (p.exact_main && !p.conv1) {
                #pragma unroll
                for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) - __uint_as_float(r[i] & 0xFFFFE000u));
                tmem_st32(slot + 32u, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[it & 1]);
        }
        if (ok && n_gdn) {
            // ---- fused GDN: four steps through ONE stage; the accumulator chunk of the next step is read meanwhile
            ok = mbar_wait(acc_full, 0, p.error_flag, 3);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            #pragma unroll 1
            for (int j = 0; j < 4 && ok; j++) {
                uint32_t* cur = r;
                tmem_ld32(lane_base + j * kChunkK, r);
                if (j >= 1) ok = mbar_wait(x_free, (uint32_t)(j - 1) & 1u, p.error_flag, 7);
                if (!ok) break;
                uint8_t* rowp = smem + row * 128;
                #pragma unroll
                for (int c = 0; c < 8; c++) {
                    float4 x = make_float4(__uint_as_float(cur[4 * c]), __uint_as_float(cur[4 * c + 1]),
                                           __uint_as_float(cur[4 * c + 2]), __uint_as_float(cur[4 * c + 3]));
                    if (p.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + j * kChunkK + 4 * c));
                        x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                    }
                    x.x *= x.x; x.y *= x.y; x.z *= x.z; x.w *= x.w;
                    float4 xl;
                    xl.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                    xl.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                    xl.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                    xl.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                    *reinterpret_cast<float4*>(rowp + ((c ^ (row & 7)) << 4)) = x;
                    *reinterpret_cast<float4*>(rowp + kTileBytes + ((c ^ (row & 7)) << 4)) = xl;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(x_ready);
            }
        }
        if (ok) ok = mbar_wait(n_gdn ? nrm_full : acc_full, 0, p.error_flag, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: all 128 channels of this row -> staging (four swizzled [128 x 32] sub-tiles) -> 512-byte pixels
        #pragma unroll 1
        for (int q = 0; q < 4; q++) {
            tmem_ld32_nowait(lane_base + q * kChunkK, r);
            if (n_gdn) tmem_ld32_nowait(lane_base + kCol5Nrm + q * kChunkK, nr);
            tmem_ld_wait();
            stage_chunk(smem + q * kTileBytes + row * 128, row, q * kChunkK, r, nr, n_gdn != 0, p.fuse, p.bias, p.beta);
        }
        named_bar_sync(1, 128);
        const int wq = warp - 2;      // rows wq, wq + 4, ...
        #pragma unroll 1
        for (int j0 = 0; j0 < kTileM / 4; j0 += 4) {
            float4 v[4];
            float* dst[4];
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int rr = wq + 4 * (j0 + j);
                const int a = a0 + (rr >> 4), b = b0 + (rr & 15);
                const int oy = a * p.out_mul + p.out_r, ox = b * p.out_mul + p.out_s;
                size_t opix;
                if (p.out_split)
                    opix = (((size_t)img * 4 + (size_t)((oy & 1) * 2 + (ox & 1))) * (p.Hout / 2) + (oy >> 1)) * (p.Wout / 2) + (ox >> 1);
                else
                    opix = ((size_t)img * p.Hout + oy) * p.Wout + ox;
                dst[j] = (ok && a < p.Hg && b < p.Wg) ? p.out + opix * kCout + lane * 4 : nullptr;
                v[j] = *reinterpret_cast<const float4*>(smem + (lane >> 3) * kTileBytes + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            }
            #pragma unroll
            for (int j = 0; j < 4; j++)
                if (dst[j]) *reinterpret_cast<float4*>(dst[j]) = v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (p.times && threadIdx.x == 64) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.times[(size_t)blockIdx.x * 3 + 2] = t;
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols5) : "memory");
    }
}

// =================================================================================================
// Version 6: the LAST layer (conv2d_transpose k9 s4, 128 -> 1, components.py:79-84) with its col2im gather and the
// BT.601 cast (tools.py:61-93) inside the kernel.
//
// Measured on version 5 (profiles/r01_ncu_full_gemm_layers_final.md): the per-position tap matrix [positions, 128] is
// written to HBM (257 MB per 24 images) only to be read back by col2im_k9s4_kernel (another 302 MB + 94 us), for 9 MB
// of pixels. Here a CTA contracts a tile of 8 x 16 positions (as version 5: thread = position = TMEM lane, two CTAs
// per SM), dumps the 81 tap columns of its accumulator to shared memory and gathers the pixels of the 6 x 14 pixel
// blocks whose contributing positions all lie inside the tile: pixel row oy = 4 q + r - 2 (block q, r in [0, 4))
// receives position q through ky = r, q - 1 through ky = r + 4 and, for r = 0, q - 2 through ky = 8 - so blocks
// [q0, q0 + 6) need positions [q0 - 2, q0 + 6). Tiles overlap by two positions (65 % of the contracted rows are
// new); positions outside the layer's input are zero-filled by TMA and contribute exact zeros, which is what the
// skip in col2im_k9s4_kernel amounts to. The sum runs in that kernel's order, so the two paths agree bit for bit.
// MMA N = 96 (81 taps used). HBM traffic: the activations once (the overlap is served by L2) + the pixels.
constexpr int kColStride6 = 87;                    // odd: the thread-per-row dump is conflict-free
constexpr int kBlkY6 = 6, kBlkX6 = 14;             // pixel blocks (4 x 4 pixels) a tile completes
constexpr int kSmemBytes6 = kMainBytes5 + 256 + 1024;
constexpr uint32_t kInstrDescN96 = (1u << 4) | (2u << 7) | (2u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);
static_assert(kTileM * kColStride6 * 4 <= kMainBytes5, "tap columns alias the stages");

struct UmmaParams6 {
    int kchunks;
    int tiles_x, tiles_y;
    int H, W;                // output image
    uint8_t* out_u8;         // [n, H, W] or NULL
    float* out_f32;          // [n, H, W] un-clipped, or NULL
    int exact_main;
    uint32_t* error_flag;
};

__device__ __forceinline__ void umma_tf32_ts_n96(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kInstrDescN96), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(256, 2)
tconv9s4_umma6_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                      const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ UmmaParams6 p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMainBytes5);
    uint64_t* full = bars;                 // [2] stage landed
    uint64_t* done = bars + 2;             // [2] MMAs of the iteration that used the stage completed
    uint64_t* split = bars + 4;            // [2] TMEM A slot written (one arrival per conversion warp)
    uint64_t* acc_full = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int q0 = (trem / p.tiles_x) * kBlkY6, p0 = (trem % p.tiles_x) * kBlkX6;      // first pixel block of the tile
    const int a0 = q0 - 2, b0 = p0 - 2;                                                // first position of the tile
    constexpr int kStageBytes = 3 * kTileBytes;                                         // A | B_hi | B_lo
    const int n_main = p.kchunks;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); mbar_init(&split[s], 4); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols5) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;      // 0 or 256: two CTAs share the SM

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            bool ok = true;
            for (int it = 0; it < n_main && ok; it++) {
                const int s = it & 1;
                if (it >= 2) ok = mbar_wait(&done[s], (uint32_t)((it >> 1) - 1) & 1u, p.error_flag, 0);
                if (!ok) break;
                uint8_t* st = smem + s * kStageBytes;
                mbar_expect_tx(&full[s], kTileBytes + (p.exact_main ? 2 : 1) * 96 * 128);
                tma_load_5d(st, &map_a, &full[s], it * kChunkK, b0, a0, 0, img);        // rows outside the input: zeros
                tma_load_3d(st + kTileBytes, &map_b_hi, &full[s], it * kChunkK, 0, 0);
                if (p.exact_main) tma_load_3d(st + 2 * kTileBytes, &map_b_lo, &full[s], it * kChunkK, 0, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        const uint32_t tb = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        bool ok = true;
        for (int it = 0; it < n_main && ok; it++) {
            const int s = it & 1;
            ok = __all_sync(0xFFFFFFFFu, mbar_wait(&split[s], (uint32_t)(it >> 1) & 1u, p.error_flag, 1));
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t st = smem_u32(smem + s * kStageBytes + kTileBytes);
                const uint32_t a_hi = tb + kCol5Slots + 64u * (uint32_t)s, a_lo = a_hi + 32u;
                #pragma unroll
                for (int k = 0; k < kChunkK / 8; k++) {
                    const uint64_t b_hi = make_desc(st + k * 32);
                    umma_tf32_ts_n96(tb, a_hi + 8 * k, b_hi, (it == 0 && k == 0) ? 0u : 1u);
                    if (p.exact_main) {
                        umma_tf32_ts_n96(tb, a_lo + 8 * k, b_hi, 1u);
                        umma_tf32_ts_n96(tb, a_hi + 8 * k, make_desc(st + kTileBytes + k * 32), 1u);
                    }
                }
                umma_commit(&done[s]);
                if (it == n_main - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ===== warps 2..5: operand conversion, then the gather (thread = position = accumulator row = TMEM lane) =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        bool ok = true;
        uint32_t r[32];
        for (int it = 0; it < n_main && ok; it++) {
            const int s = it & 1;
            ok = mbar_wait(&full[s], (uint32_t)(it >> 1) & 1u, p.error_flag, 2);
            if (!ok) break;
            const uint8_t* rowp = smem + s * kStageBytes + row * 128;
            #pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (row & 7)) << 4));
                r[4 * c + 0] = __float_as_uint(v.x); r[4 * c + 1] = __float_as_uint(v.y);
                r[4 * c + 2] = __float_as_uint(v.z); r[4 * c + 3] = __float_as_uint(v.w);
            }
            if (it >= 2) {      // the MMAs of iteration it - 2 read this TMEM slot
                ok = mbar_wait(&done[s], (uint32_t)((it >> 1) - 1) & 1u, p.error_flag, 5);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t slot = lane_base + kCol5Slots + 64u * (uint32_t)s;
            if (!p.exact_main) {      // single pass: round to nearest TF32 (see version 4)
                #pragma unroll
                for (int i = 0; i < 32; i++) r[i] += 0x1000u;
            }
            tmem_st32(slot, r);
            if (p.exact_main) {
                #pragma unroll
                for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) - __uint_as_float(r[i] & 0xFFFFE000u));
                tmem_st32(slot + 32u, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
        if (ok) ok = mbar_wait(acc_full, 0, p.error_flag, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- tap columns of this position -> shared memory (every MMA, hence every read of the stages, has completed)
        float* col = reinterpret_cast<float*>(smem);
        {
            float* mine = col + row * kColStride6;
            tmem_ld32(lane_base, r);
            #pragma unroll
            for (int i = 0; i < 32; i++) mine[i] = __uint_as_float(r[i]);
            tmem_ld32(lane_base + 32u, r);
            #pragma unroll
            for (int i = 0; i < 32; i++) mine[32 + i] = __uint_as_float(r[i]);
            tmem_ld32(lane_base + 64u, r);
            #pragma unroll
            for (int i = 0; i < 17; i++) mine[64 + i] = __uint_as_float(r[i]);
        }
        named_bar_sync(1, 128);
        // ---- gather: one item = two horizontally adjacent pixels (s0, s0 + 1) of pixel block (qa, pb), row r
        const int t = threadIdx.x - 64;
        #pragma unroll 1
        for (int item = t; item < 4 * kBlkY6 * 2 * kBlkX6; item += 128) {
            const int ly = item / (2 * kBlkX6), pr = item - ly * (2 * kBlkX6);
            const int qa = ly >> 2, rr = ly & 3;
            const int pb = pr >> 1, s0 = (pr & 1) * 2;
            const int oy = 4 * (q0 + qa) + rr - 2, ox = 4 * (p0 + pb) + s0 - 2;
            if (!ok || oy < 0 || oy >= p.H || ox < 0 || ox >= p.W) continue;
            float acc0 = 0.f, acc1 = 0.f;
            #pragma unroll
            for (int da = 0; da < 3; da++) {
                const int ky = rr + 4 * da;
                if (ky > 8) continue;
                const float* rowc = col + ((qa + 2 - da) * 16 + pb + 2) * kColStride6 + ky * 9 + s0;
                #pragma unroll
                for (int db = 0; db < 3; db++) {
                    const float* c = rowc - db * kColStride6 + 4 * db;      // position pb + 2 - db, tap kx = s0 + 4 db
                    if (s0 + 4 * db <= 8) acc0 += c[0];
                    if (s0 + 4 * db + 1 <= 8) acc1 += c[1];
                }
            }
            const size_t at = ((size_t)img * p.H + oy) * p.W + ox;
            if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + at) = make_float2(acc0, acc1);
            if (p.out_u8) {
                uchar2 v;
                v.x = (uint8_t)(int)rintf(fminf(fmaxf(acc0, 16.f), 235.f));
                v.y = (uint8_t)(int)rintf(fminf(fmaxf(acc1, 16.f), 235.f));
                *reinterpret_cast<uchar2*>(p.out_u8 + at) = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols5) : "memory");
    }
}

// =================================================================================================
// Version 7: version 6 as a PERSISTENT, fully pipelined kernel (one CTA per SM, tiles round-robin).
//
// Measured on version 6 (two 128-position CTAs per SM): 170 us per 24 images = 13 k cycles per CTA, of which the
// tensor pipe needs 2.3 k (3xTF32) and shared-memory / L2 bandwidth less - the time is the serial chain TMEM
// allocation -> first TMA round trip -> 4 x (convert -> MMA) -> dump -> gather -> stores of every CTA, and each CTA
// re-fetches the 96 KB of split weights from L2. Here the weights are loaded ONCE per SM and stay in shared memory,
// the accumulator is double-buffered in TMEM, and four warp roles run tiles j + 1 / j / j - 1 concurrently:
//
//   warp 0      TMA: weights once, then the activation chunks (4-stage ring = one tile in flight)
//   warps 2-5   operand conversion: shared memory -> {hi | lo} TF32 split -> TMEM slot (4 slots)
//   warp 1      MMA issue into accumulator (tile & 1)
//   warps 6-13  epilogue: accumulator -> tap columns in shared memory -> gather -> pixels (two warps per TMEM lane
//               quarter, 48 tap columns each), while the next tile is being contracted
//
//  smem: B_hi 4 x 12 KB | B_lo 4 x 12 KB | A ring 4 x 16 KB | tap columns 43.5 KB | barriers
//  TMEM (512 columns): ACC0 [0,96) | ACC1 [128,224) | slots [256 + 64 kc, +64) = {hi 32 | lo 32}
constexpr int kThreads7 = 448;
constexpr int kBChunkBytes7 = 96 * 128;
constexpr int kOffBlo7 = 4 * kBChunkBytes7;
constexpr int kOffA7 = 8 * kBChunkBytes7;
constexpr int kOffCol7 = kOffA7 + 4 * kTileBytes;
constexpr int kOffBars7 = kOffCol7 + ((kTileM * kColStride6 * 4 + 1023) / 1024) * 1024;
constexpr int kSmemBytes7 = kOffBars7 + 256 + 1024;
static_assert(kSmemBytes7 <= 227 * 1024, "version 7 shared memory");

struct UmmaParams7 {
    int n_tiles;             // n images x tiles_y x tiles_x
    int tiles_x, tiles_y;
    int H, W;
    uint8_t* out_u8;
    float* out_f32;
    int exact_main;
    uint32_t* error_flag;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__global__ void __launch_bounds__(kThreads7, 1)
tconv9s4_umma7_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                      const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ UmmaParams7 p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space: LDS / STS, not generic LD / ST
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars7);
    uint64_t* b_full = bars;               // weights resident
    uint64_t* a_full = bars + 1;           // [4] activation chunk landed
    uint64_t* a_free = bars + 5;           // [4] the conversion warps have read it (4 arrivals)
    uint64_t* split = bars + 9;            // [4] TMEM slot written (4 arrivals)
    uint64_t* slot_free = bars + 13;       // [4] the MMAs that read the slot completed
    uint64_t* acc_full = bars + 17;        // [2] the last MMA of a tile completed
    uint64_t* acc_empty = bars + 19;       // [2] the epilogue warps have read the accumulator (8 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int n_mine = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        mbar_init(b_full, 1);
        for (int s = 0; s < 4; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_free[s], 4); mbar_init(&split[s], 4); mbar_init(&slot_free[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && n_mine > 0) {
            mbar_expect_tx(b_full, (p.exact_main ? 8 : 4) * kBChunkBytes7);
            for (int kc = 0; kc < 4; kc++) {
                tma_load_3d(smem + kc * kBChunkBytes7, &map_b_hi, b_full, kc * kChunkK, 0, 0);
                if (p.exact_main) tma_load_3d(smem + kOffBlo7 + kc * kBChunkBytes7, &map_b_lo, b_full, kc * kChunkK, 0, 0);
            }
            bool ok = true;
            for (int j = 0; j < n_mine && ok; j++) {
                const int tile = (int)blockIdx.x + j * (int)gridDim.x;
                const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
                const int a0 = (trem / p.tiles_x) * kBlkY6 - 2, b0 = (trem % p.tiles_x) * kBlkX6 - 2;
                for (int kc = 0; kc < 4 && ok; kc++) {
                    if (j >= 1) ok = mbar_wait(&a_free[kc], (uint32_t)(j - 1) & 1u, p.error_flag, 0);
                    if (!ok) break;
                    mbar_expect_tx(&a_full[kc], kTileBytes);
                    tma_load_5d(smem + kOffA7 + kc * kTileBytes, &map_a, &a_full[kc], kc * kChunkK, b0, a0, 0, img);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        const uint32_t tb = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
        bool ok = n_mine > 0;
        if (ok) ok = __all_sync(0xFFFFFFFFu, mbar_wait(b_full, 0, p.error_flag, 1));
        for (int j = 0; j < n_mine && ok; j++) {
            const uint32_t acc = tb + 128u * (uint32_t)(j & 1);
            if (j >= 2) {      // the epilogue of tile j - 2 has read this accumulator
                ok = __all_sync(0xFFFFFFFFu, mbar_wait(&acc_empty[j & 1], (uint32_t)((j >> 1) - 1) & 1u, p.error_flag, 1));
                if (!ok) break;
            }
            for (int kc = 0; kc < 4 && ok; kc++) {
                ok = __all_sync(0xFFFFFFFFu, mbar_wait(&split[kc], (uint32_t)j & 1u, p.error_flag, 1));
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t bh = smem_u32(smem + kc * kBChunkBytes7), bl = smem_u32(smem + kOffBlo7 + kc * kBChunkBytes7);
                    const uint32_t a_hi = tb + 256u + 64u * (uint32_t)kc, a_lo = a_hi + 32u;
                    #pragma unroll
                    for (int k = 0; k < kChunkK / 8; k++) {
                        const uint64_t b_hi = make_desc(bh + k * 32);
                        umma_tf32_ts_n96(acc, a_hi + 8 * k, b_hi, (kc == 0 && k == 0) ? 0u : 1u);
                        if (p.exact_main) {
                            umma_tf32_ts_n96(acc, a_lo + 8 * k, b_hi, 1u);
                            umma_tf32_ts_n96(acc, a_hi + 8 * k, make_desc(bl + k * 32), 1u);
                        }
                    }
                    umma_commit(&slot_free[kc]);
                    if (kc == 3) umma_commit(&acc_full[j & 1]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ===== warps 2..5: operand conversion (thread = position = TMEM lane) =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        bool ok = true;
        uint32_t r[32];
        for (int j = 0; j < n_mine && ok; j++) {
            for (int kc = 0; kc < 4 && ok; kc++) {
                ok = mbar_wait(&a_full[kc], (uint32_t)j & 1u, p.error_flag, 2);
                if (!ok) break;
                const uint8_t* rowp = smem + kOffA7 + kc * kTileBytes + row * 128;
                #pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ (row & 7)) << 4));
                    r[4 * c + 0] = __float_as_uint(v.x); r[4 * c + 1] = __float_as_uint(v.y);
                    r[4 * c + 2] = __float_as_uint(v.z); r[4 * c + 3] = __float_as_uint(v.w);
                }
                if (j >= 1) {      // the MMAs of the previous tile read this TMEM slot
                    ok = mbar_wait(&slot_free[kc], (uint32_t)(j - 1) & 1u, p.error_flag, 5);
                    if (!ok) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t slot = lane_base + 256u + 64u * (uint32_t)kc;
                if (!p.exact_main) {      // single pass: round to nearest TF32 (see version 4)
                    #pragma unroll
                    for (int i = 0; i < 32; i++) r[i] += 0x1000u;
                }
                tmem_st32(slot, r);       // (the stores consume the registers: the shared-memory reads above are complete)
                if (p.exact_main) {
                    #pragma unroll
                    for (int i = 0; i < 32; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) - __uint_as_float(r[i] & 0xFFFFE000u));
                    tmem_st32(slot + 32u, r);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) { mbar_arrive(&a_free[kc]); mbar_arrive(&split[kc]); }
            }
        }
    } else {
        // ===== warps 6..13: epilogue. Once a wait has failed the warp keeps running the barriers without working. =====
        const int e = warp - 6;
        const int quarter = warp & 3, half = e >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* col = reinterpret_cast<float*>(smem + kOffCol7);
        const int t = threadIdx.x - 192;
        bool ok = true;
        for (int j = 0; j < n_mine; j++) {
            const int tile = (int)blockIdx.x + j * (int)gridDim.x;
            const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
            const int q0 = (trem / p.tiles_x) * kBlkY6, p0 = (trem % p.tiles_x) * kBlkX6;
            if (ok) ok = mbar_wait(&acc_full[j & 1], (uint32_t)(j >> 1) & 1u, p.error_flag, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ok) {
                // tap columns [48 half, +48) of this position -> shared memory (the previous tile's gather has finished)
                uint32_t v[48];
                const uint32_t src = lane_base + 128u * (uint32_t)(j & 1) + 48u * (uint32_t)half;
                tmem_ld16_nowait(src, v);
                tmem_ld16_nowait(src + 16u, v + 16);
                tmem_ld16_nowait(src + 32u, v + 32);
                tmem_ld_wait();
                float* mine = col + row * kColStride6 + 48 * half;
                #pragma unroll
                for (int i = 0; i < 48; i++)
                    if (i < 33 || half == 0) mine[i] = __uint_as_float(v[i]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0 && ok) mbar_arrive(&acc_empty[j & 1]);
            named_bar_sync(1, 256);
            // ---- gather: one item = two horizontally adjacent pixels (s0, s0 + 1) of pixel block (qa, pb), row rr
            #pragma unroll 1
            for (int item = t; item < 4 * kBlkY6 * 2 * kBlkX6; item += 256) {
                const int ly = item / (2 * kBlkX6), pr = item - ly * (2 * kBlkX6);
                const int qa = ly >> 2, rr = ly & 3;
                const int pb = pr >> 1, s0 = (pr & 1) * 2;
                const int oy = 4 * (q0 + qa) + rr - 2, ox = 4 * (p0 + pb) + s0 - 2;
                if (!ok || oy < 0 || oy >= p.H || ox < 0 || ox >= p.W) continue;
                float acc0 = 0.f, acc1 = 0.f;
                #pragma unroll
                for (int da = 0; da < 3; da++) {
                    const int ky = rr + 4 * da;
                    if (ky > 8) continue;
                    const float* rowc = col + ((qa + 2 - da) * 16 + pb + 2) * kColStride6 + ky * 9 + s0;
                    #pragma unroll
                    for (int db = 0; db < 3; db++) {
                        const float* c = rowc - db * kColStride6 + 4 * db;      // position pb + 2 - db, tap kx = s0 + 4 db
                        if (s0 + 4 * db <= 8) acc0 += c[0];
                        if (s0 + 4 * db + 1 <= 8) acc1 += c[1];
                    }
                }
                const size_t at = ((size_t)img * p.H + oy) * p.W + ox;
                if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + at) = make_float2(acc0, acc1);
                if (p.out_u8) {
                    uchar2 v;
                    v.x = (uint8_t)(int)rintf(fminf(fmaxf(acc0, 16.f), 235.f));
                    v.y = (uint8_t)(int)rintf(fminf(fmaxf(acc1, 16.f), 235.f));
                    *reinterpret_cast<uchar2*>(p.out_u8 + at) = v;
                }
            }
            named_bar_sync(1, 256);      // the tap columns may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint32_t* box)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EAE_ERR_CUDA; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bdim[5], estride[5];
    uint64_t stride = sizeof(float);
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estride[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bdim,
                    estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return EAE_ERR_CUDA; }
    return 0;
}

// uint8 image [n, H, W] -> boxes of kImgBoxW x kImgBoxH pixels of one image, no swizzle, zero fill outside.
int make_map_u8(CUtensorMap* map, const void* base, uint64_t W, uint64_t H, uint64_t n)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EAE_ERR_CUDA; }
    const cuuint64_t gdim[3] = {W, H, n};
    const cuuint64_t gstride[2] = {W, W * H};
    const cuuint32_t bdim[3] = {(cuuint32_t)kImgBoxW, (cuuint32_t)kImgBoxH, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, bdim, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (uint8 image) failed with CUresult %d", (int)r); return EAE_ERR_CUDA; }
    return 0;
}

uint32_t* g_error_flag = nullptr;   // device word, per process (one device per process in practice)

}  // namespace

int umma_available()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return EAE_ERR_CUDA; }
    cudaDeviceProp prop;
    EAE_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("the tcgen05 path needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
        return EAE_ERR_CUDA;
    }
    if (!encode_tiled_fn()) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EAE_ERR_CUDA; }
    return 0;
}

uint32_t* umma_error_flag_dev() { return g_error_flag; }

int umma_check_error(cudaStream_t st)
{
    if (!g_error_flag) return 0;
    uint32_t flag = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(&flag, g_error_flag, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (flag) {
        cudaMemsetAsync(g_error_flag, 0, 4, st);
        set_error("tcgen05 GEMM pipeline timed out (role mask 0x%x)", flag);
        return EAE_ERR_CUDA;
    }
    return 0;
}

int umma_version()
{
    static int v = 0;
    if (!v) {
        const char* env = getenv("EAE_UMMA_VERSION");
        v = env ? atoi(env) : 7;
        if (v < 1 || v > 7) v = 7;
    }
    return v;
}

int launch_gemm_umma(const GemmPlan& plan, const UmmaWeights& w, const UmmaWeights* gamma, bool exact3x,
                     cudaStream_t st, const GemmPlan* more, int n_more)
{
    if (plan.M == 0) return 0;
    if (n_more > 3) { set_error("gemm_umma: at most four plans per launch"); return EAE_ERR_ARGUMENT; }
    if (plan.Cin % kChunkK != 0 || plan.n_taps < 1 || plan.n_taps > kMaxTaps || !w.hi || (exact3x && !w.lo)) {
        set_error("gemm_umma: bad plan (Cin %d, taps %d)", plan.Cin, plan.n_taps);
        return EAE_ERR_ARGUMENT;
    }
    if (!g_error_flag) {
        EAE_CUDA_OK(cudaMalloc(&g_error_flag, 4));
        EAE_CUDA_OK(cudaMemset(g_error_flag, 0, 4));
    }
    const uint32_t per_img = (uint32_t)(plan.Hg * plan.Wg);
    const uint32_t n_img = plan.M / per_img;
    UmmaParams p;
    memset(&p, 0, sizeof p);
    p.n_taps = plan.n_taps;
    p.kchunks = plan.Cin / kChunkK;
    // Position grids that are wide enough use 16 x 8 tiles; flat (1-row) grids use 128 x 1.
    // (a fused GDN tail stores 16 x 16 tiles, so a one-row grid with fusion keeps the 2-D tiling)
    if (plan.Hg == 1 && !plan.fuse) { p.tile_w = 128; p.tile_h = 1; } else { p.tile_w = 16; p.tile_h = 8; }
    p.tiles_x = (plan.Wg + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (plan.Hg + p.tile_h - 1) / p.tile_h;
    p.Hg = plan.Hg; p.Wg = plan.Wg;
    p.out = plan.out; p.bias = plan.bias; p.xin = plan.in;
    p.Hout = plan.Hout; p.Wout = plan.Wout; p.out_mul = plan.out_mul; p.out_r = plan.out_r; p.out_s = plan.out_s;
    p.out_split = plan.out_split;
    p.mode = plan.mode;
    p.error_flag = g_error_flag;
    // Input planes: natural NHWC (1 plane) or parity-split (4 planes of Hin/2 x Win/2).
    const int planes = plan.in_split ? 4 : 1;
    const int Hp = plan.in_split ? plan.Hin / 2 : plan.Hin, Wp = plan.in_split ? plan.Win / 2 : plan.Win;
    for (int t = 0; t < plan.n_taps; t++) {
        const int dy = plan.taps[t].dy, dx = plan.taps[t].dx;
        UmmaTap& u = p.taps[t];
        if (plan.in_split) {
            if (plan.in_mul != 2) { set_error("gemm_umma: split input needs in_mul 2"); return EAE_ERR_ARGUMENT; }
            u.plane = (dy & 1) * 2 + (dx & 1);
            u.fy = (dy - (dy & 1)) / 2;      // floor(dy / 2)
            u.fx = (dx - (dx & 1)) / 2;
        } else {
            if (plan.in_mul != 1) { set_error("gemm_umma: natural input needs in_mul 1"); return EAE_ERR_ARGUMENT; }
            u.plane = 0; u.fy = dy; u.fx = dx;
        }
        u.w_tap = (int)(plan.taps[t].w_off / ((uint32_t)plan.Cin * kCout));
    }
    CUtensorMap map_a, map_b_hi, map_b_lo;
    const uint64_t bdims[3] = {(uint64_t)plan.Cin, kCout, (uint64_t)w.n_taps};
    const uint32_t bbox[3] = {kChunkK, kCout, 1};
    EAE_TRY(make_map(&map_b_hi, w.hi, 3, bdims, bbox));
    EAE_TRY(make_map(&map_b_lo, exact3x ? w.lo : w.hi, 3, bdims, bbox));
    if (plan.img_u8) {
        if (umma_version() < 3 || plan.n_taps != 1 || plan.Cin != 96 || plan.Hg * 4 != plan.img_H ||
            plan.Wg * 4 != plan.img_W || plan.img_W % 16 != 0) {
            set_error("gemm_umma: the fused k9 s4 convolution needs kernel version 3 and a [n, 4 Hg, 4 Wg] image");
            return EAE_ERR_ARGUMENT;
        }
        map_a = map_b_hi;      // not read
    } else {
        const uint64_t adims[5] = {(uint64_t)plan.Cin, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)planes, n_img};
        const uint32_t abox[5] = {kChunkK, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1, 1};
        EAE_TRY(make_map(&map_a, plan.in, 5, adims, abox));
    }
    const uint32_t grid = n_img * (uint32_t)(p.tiles_x * p.tiles_y);
    // Builds the version-4 parameters of one plan; false if its taps do not fit the union boxes.
    auto build4 = [&](const GemmPlan& plan, UmmaParams4& q) -> bool {
        UmmaParams p;      // (shadows the outer one: taps of THIS plan)
        for (int t = 0; t < plan.n_taps; t++) {
            const int dy = plan.taps[t].dy, dx = plan.taps[t].dx;
            UmmaTap& u = p.taps[t];
            if (plan.in_split) { u.plane = (dy & 1) * 2 + (dx & 1); u.fy = (dy - (dy & 1)) / 2; u.fx = (dx - (dx & 1)) / 2; }
            else { u.plane = 0; u.fy = dy; u.fx = dx; }
            u.w_tap = (int)(plan.taps[t].w_off / ((uint32_t)plan.Cin * kCout));
        }
        memset(&q, 0, sizeof q);
        q.n_taps = plan.n_taps; q.kchunks = plan.Cin / kChunkK;
        q.tiles_x = (plan.Wg + 15) / 16; q.tiles_y = (plan.Hg + 15) / 16;
        q.Hg = plan.Hg; q.Wg = plan.Wg;
        q.out = plan.out; q.bias = plan.bias; q.beta = plan.fuse_beta;
        q.Hout = plan.Hout; q.Wout = plan.Wout; q.out_mul = plan.out_mul; q.out_r = plan.out_r; q.out_s = plan.out_s;
        q.out_split = plan.out_split;
        q.fuse = plan.fuse; q.exact_main = exact3x ? 1 : 0; q.exact_gdn = plan.fuse_single_pass ? 0 : 1;
        q.error_flag = g_error_flag;
        // groups = input planes in use; taps sorted by group
        int fy_min[kMaxGroups4], fx_min[kMaxGroups4], fy_max[kMaxGroups4], fx_max[kMaxGroups4], group_of_plane[4] = {-1, -1, -1, -1};
        int n_groups = 0;
        for (int t = 0; t < plan.n_taps; t++) {
            const UmmaTap& u = p.taps[t];
            int g = group_of_plane[u.plane];
            if (g < 0) {
                g = group_of_plane[u.plane] = n_groups++;
                fy_min[g] = fy_max[g] = u.fy; fx_min[g] = fx_max[g] = u.fx;
            }
            if (u.fy < fy_min[g]) fy_min[g] = u.fy;
            if (u.fy > fy_max[g]) fy_max[g] = u.fy;
            if (u.fx < fx_min[g]) fx_min[g] = u.fx;
            if (u.fx > fx_max[g]) fx_max[g] = u.fx;
        }
        bool fits = true;
        for (int g = 0; g < n_groups; g++) fits = fits && fy_max[g] - fy_min[g] <= kUnionH - 16 && fx_max[g] - fx_min[g] <= kUnionW - 16;
        if (!fits || plan.mode != kEpiBias) return false;
        {
            q.n_groups = n_groups;
            int nt = 0;
            for (int g = 0; g < n_groups; g++) {
                int plane = 0;
                for (int pl = 0; pl < 4; pl++) if (group_of_plane[pl] == g) plane = pl;
                q.groups[g] = UmmaGroup4{plane, fy_min[g], fx_min[g], 0};
                for (int t = 0; t < plan.n_taps; t++) {
                    const UmmaTap& u = p.taps[t];
                    if (group_of_plane[u.plane] != g) continue;
                    q.taps[nt++] = UmmaTap4{u.w_tap, (u.fy - fy_min[g]) * kUnionW + (u.fx - fx_min[g]), g, 0};
                }
                q.taps[nt - 1].last = 1;
                q.groups[g].pad = nt - 1;      // index of the group's last tap in the sorted list
            }
        }
        return true;
    };
    auto eligible4 = [&](const GemmPlan& pl) {
        return umma_version() >= 4 && pl.n_taps > 1 && pl.n_taps <= kMaxTaps && pl.Hg > 1 && pl.Cin == 128 && !pl.img_u8 &&
               pl.in == plan.in && pl.Hg == plan.Hg && pl.Wg == plan.Wg && pl.in_split == plan.in_split && pl.M == plan.M &&
               pl.fuse == plan.fuse && pl.fuse_single_pass == plan.fuse_single_pass;
    };
    UmmaParams4x qx;
    memset(&qx, 0, sizeof qx);
    bool all4 = eligible4(plan) && build4(plan, qx.ph[0]);
    for (int i = 0; i < n_more && all4; i++) all4 = eligible4(more[i]) && build4(more[i], qx.ph[1 + i]);
    if (n_more > 0 && !all4) {      // no common launch: one after the other
        EAE_TRY(launch_gemm_umma(plan, w, gamma, exact3x, st, nullptr, 0));
        for (int i = 0; i < n_more; i++) EAE_TRY(launch_gemm_umma(more[i], w, gamma, exact3x, st, nullptr, 0));
        return 0;
    }
    if (all4) {
        {
            const UmmaParams4& q = qx.ph[0];
            qx.n_phases = 1 + n_more;
            CUtensorMap map_u;
            const uint64_t adims[5] = {(uint64_t)plan.Cin, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)planes, n_img};
            const uint32_t ubox[5] = {kChunkK, (uint32_t)kUnionW, (uint32_t)kUnionH, 1, 1};
            EAE_TRY(make_map(&map_u, plan.in, 5, adims, ubox));
            CUtensorMap map_g_hi = map_b_hi, map_g_lo = map_b_lo;
            if (plan.fuse) {
                if (!gamma || !gamma->hi || !gamma->lo || !plan.fuse_beta) {
                    set_error("gemm_umma: fused GDN needs gamma hi/lo and beta");
                    return EAE_ERR_ARGUMENT;
                }
                const uint64_t gdims[3] = {kCout, kCout, 1};
                EAE_TRY(make_map(&map_g_hi, gamma->hi, 3, gdims, bbox));
                EAE_TRY(make_map(&map_g_lo, gamma->lo, 3, gdims, bbox));
            }
            static bool attr4_done = false;
            if (!attr4_done) {
                EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
                attr4_done = true;
            }
            const uint32_t grid4 = n_img * (uint32_t)(q.tiles_x * q.tiles_y) * (uint32_t)qx.n_phases;
            static int timing4 = -1;
            if (timing4 < 0) { const char* e = getenv("EAE_UMMA_TIMING"); timing4 = e ? atoi(e) : 0; }
            // (EAE_UMMA_TIMING=4: only this kernel, from a buffer that is allocated once - cudaMalloc / cudaFree wait for
            //  every stream of the device, which defeats measurements beside kernels running on other streams)
            static long long* d_times = nullptr;
            static size_t d_times_cap = 0;
            if (timing4) {
                if (d_times_cap < grid4) {
                    if (d_times) cudaFree(d_times);
                    d_times_cap = grid4 > 8192 ? grid4 : 8192;
                    EAE_CUDA_OK(cudaMalloc(&d_times, d_times_cap * 12 * sizeof(long long)));
                }
                EAE_CUDA_OK(cudaMemsetAsync(d_times, 0, (size_t)grid4 * 12 * sizeof(long long), st));
                for (int i = 0; i < qx.n_phases; i++) qx.ph[i].times = d_times;
            }
            gemm_umma4_kernel<<<grid4, kUmmaThreads3, kSmemBytes4, st>>>(map_u, map_b_hi, map_b_lo, map_g_hi, map_g_lo, qx);
            EAE_LAUNCH_OK();
            if (timing4) {
                std::vector<long long> h((size_t)grid4 * 12);
                EAE_CUDA_OK(cudaMemcpyAsync(h.data(), d_times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
                EAE_CUDA_OK(cudaStreamSynchronize(st));
                double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (uint32_t b = 0; b < grid4; b++)
                    for (int j = 1; j < 8; j++) acc[j] += (double)(h[(size_t)b * 12 + j] - h[(size_t)b * 12]);
                std::vector<long long> dur(grid4);
                for (uint32_t b = 0; b < grid4; b++) dur[b] = h[(size_t)b * 12 + 7] - h[(size_t)b * 12];
                std::sort(dur.begin(), dur.end());
                // which SMs ran the CTAs, and when (global timer): makespan against the busy time of the SMs
                {
                    int per_sm[256] = {0};
                    long long t_min = h[8], t_max = h[9];
                    double busy_ns = 0.;
                    for (uint32_t b = 0; b < grid4; b++) {
                        per_sm[h[(size_t)b * 12 + 10] & 255]++;
                        if (h[(size_t)b * 12 + 8] < t_min) t_min = h[(size_t)b * 12 + 8];
                        if (h[(size_t)b * 12 + 9] > t_max) t_max = h[(size_t)b * 12 + 9];
                        busy_ns += (double)(h[(size_t)b * 12 + 9] - h[(size_t)b * 12 + 8]);
                    }
                    int used = 0, most = 0, least = 1 << 30;
                    for (int i = 0; i < 256; i++) if (per_sm[i]) { used++; most = per_sm[i] > most ? per_sm[i] : most; least = per_sm[i] < least ? per_sm[i] : least; }
                    fprintf(stderr, "umma4 grid %u: %d SMs used, %d..%d CTAs per SM, makespan %.1f us, CTA time summed / 148 = %.1f us\n",
                            grid4, used, least, most, (double)(t_max - t_min) * 1e-3, busy_ns * 1e-3 / 148.);
                }
                fprintf(stderr, "umma4 taps %d groups %d fuse %d grid %u: setup %.0f first_union %.0f acc_seen %.0f nrm_seen %.0f "
                                "staged %.0f end %.0f (avg cycles from CTA start, conversion warp 2); CTA duration min %lld "
                                "p10 %lld median %lld p90 %lld max %lld\n",
                        q.n_taps, q.n_groups, q.fuse, grid4, acc[1] / grid4, acc[2] / grid4, acc[4] / grid4, acc[5] / grid4,
                        acc[6] / grid4, acc[7] / grid4, dur[0], dur[grid4 / 10], dur[grid4 / 2], dur[grid4 * 9 / 10], dur[grid4 - 1]);
            }
            return 0;
        }
    }
    // Version 5 pays off for the thin layers WITHOUT a fused GDN (the last layer: 195 -> 152 us per 24 images); with the
    // fused tail (layer 1) two half-size CTAs contend for the tensor pipe and shared-memory bandwidth of the GDN steps and
    // the single GDN stage serialises them: 290 us against 250 us on version 3 (EAE_UMMA_V5_FUSED=1 routes it here anyway).
    static int v5_fused = -1;
    if (v5_fused < 0) { const char* e = getenv("EAE_UMMA_V5_FUSED"); v5_fused = e ? atoi(e) : 0; }
    if (umma_version() >= 5 && plan.n_taps == 1 && p.tile_h == 8 && plan.mode == kEpiBias && p.taps[0].fy == 0 &&
        p.taps[0].fx == 0 && p.taps[0].plane == 0 && (plan.img_u8 ? plan.Cin == 96 : true) && (!plan.fuse || v5_fused)) {
        // ---- version 5: thin layers, 128 positions per CTA, two CTAs per SM ----
        UmmaParams5 q;
        memset(&q, 0, sizeof q);
        q.kchunks = plan.Cin / kChunkK;
        q.conv1 = plan.img_u8 ? 1 : 0;
        q.tiles_x = (plan.Wg + 15) / 16; q.tiles_y = (plan.Hg + 7) / 8;
        q.Hg = plan.Hg; q.Wg = plan.Wg;
        q.out = plan.out; q.bias = plan.bias; q.beta = plan.fuse_beta;
        q.Hout = plan.Hout; q.Wout = plan.Wout; q.out_mul = plan.out_mul; q.out_r = plan.out_r; q.out_s = plan.out_s;
        q.out_split = plan.out_split;
        q.fuse = plan.fuse; q.exact_main = exact3x ? 1 : 0;
        q.error_flag = g_error_flag;
        CUtensorMap map_g_hi = map_b_hi, map_g_lo = map_b_lo, map_img = map_b_hi;
        if (plan.fuse) {
            if (!gamma || !gamma->hi || !gamma->lo || !plan.fuse_beta) {
                set_error("gemm_umma: fused GDN needs gamma hi/lo and beta");
                return EAE_ERR_ARGUMENT;
            }
            const uint64_t gdims[3] = {kCout, kCout, 1};
            EAE_TRY(make_map(&map_g_hi, gamma->hi, 3, gdims, bbox));
            EAE_TRY(make_map(&map_g_lo, gamma->lo, 3, gdims, bbox));
        }
        if (plan.img_u8) EAE_TRY(make_map_u8(&map_img, plan.img_u8, (uint64_t)plan.img_W, (uint64_t)plan.img_H, n_img));
        static bool attr5_done = false;
        if (!attr5_done) {
            EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes5));
            EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma5_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr5_done = true;
        }
        const uint32_t grid5 = n_img * (uint32_t)(q.tiles_x * q.tiles_y);
        if (getenv("EAE_UMMA_TIMING") && atoi(getenv("EAE_UMMA_TIMING")) == 1) {
            // Debug: which SM ran each CTA and when; prints the largest number of CTAs alive at once on one SM.
            long long* d_times = nullptr;
            EAE_CUDA_OK(cudaMalloc(&d_times, (size_t)grid5 * 3 * sizeof(long long)));
            q.times = d_times;
            gemm_umma5_kernel<<<grid5, kThreads5, kSmemBytes5, st>>>(map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, map_img, q);
            EAE_LAUNCH_OK();
            std::vector<long long> h((size_t)grid5 * 3);
            EAE_CUDA_OK(cudaMemcpyAsync(h.data(), d_times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
            EAE_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(d_times);
            int worst = 0;
            double mean_ns = 0.;
            long long t_min = h[1], t_max = h[2];
            for (uint32_t b = 0; b < grid5; b++) {
                int alive = 0;
                for (uint32_t c = 0; c < grid5; c++)
                    if (h[(size_t)c * 3] == h[(size_t)b * 3] && h[(size_t)c * 3 + 1] <= h[(size_t)b * 3 + 1] && h[(size_t)c * 3 + 2] > h[(size_t)b * 3 + 1]) alive++;
                if (alive > worst) worst = alive;
                mean_ns += (double)(h[(size_t)b * 3 + 2] - h[(size_t)b * 3 + 1]);
                if (h[(size_t)b * 3 + 1] < t_min) t_min = h[(size_t)b * 3 + 1];
                if (h[(size_t)b * 3 + 2] > t_max) t_max = h[(size_t)b * 3 + 2];
            }
            fprintf(stderr, "umma5 conv1 %d fuse %d grid %u: up to %d CTAs alive on one SM, mean CTA %.1f us, kernel %.1f us\n", q.conv1,
                    q.fuse, grid5, worst, mean_ns / grid5 * 1e-3, (double)(t_max - t_min) * 1e-3);
            return 0;
        }
        gemm_umma5_kernel<<<grid5, kThreads5, kSmemBytes5, st>>>(map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, map_img, q);
        EAE_LAUNCH_OK();
        return 0;
    }
    if (umma_version() >= 3) {
        UmmaParams2 q;
        memset(&q, 0, sizeof q);
        q.n_taps = p.n_taps; q.kchunks = p.kchunks;
        q.tile_w = p.tile_w; q.tile_h = p.tile_h;
        q.tile_w_log2 = p.tile_w == 128 ? 7 : 4;
        if (p.tile_h == 1) { q.half_da = 0; q.half_db = p.tile_w; } else { q.half_da = p.tile_h; q.half_db = 0; }
        q.tiles_x = (plan.Wg + q.tile_w + q.half_db - 1) / (q.tile_w + q.half_db);
        q.tiles_y = (plan.Hg + q.tile_h + q.half_da - 1) / (q.tile_h + q.half_da);
        q.Hg = p.Hg; q.Wg = p.Wg;
        q.out = p.out; q.bias = p.bias; q.beta = plan.fuse_beta; q.xin = p.xin;
        q.Hout = p.Hout; q.Wout = p.Wout; q.out_mul = p.out_mul; q.out_r = p.out_r; q.out_s = p.out_s;
        q.out_split = p.out_split;
        q.mode = p.mode;
        q.fuse = plan.fuse;
        q.exact_main = exact3x ? 1 : 0;
        q.exact_gdn = plan.fuse_single_pass ? 0 : 1;
        q.cluster = 1;
        q.error_flag = p.error_flag;
        memcpy(q.taps, p.taps, sizeof q.taps);
        const uint32_t grid3 = n_img * (uint32_t)(q.tiles_x * q.tiles_y);
        q.n_tiles = (int)grid3;
        CUtensorMap map_g_hi = map_b_hi, map_g_lo = map_b_lo, map_img = map_b_hi;
        if (plan.fuse) {
            if (!gamma || !gamma->hi || !gamma->lo || !plan.fuse_beta || plan.mode != kEpiBias || p.tile_h == 1) {
                set_error("gemm_umma: fused GDN needs gamma hi/lo, beta, a bias-mode contraction and 2-D tiles");
                return EAE_ERR_ARGUMENT;
            }
            const uint64_t gdims[3] = {kCout, kCout, 1};
            EAE_TRY(make_map(&map_g_hi, gamma->hi, 3, gdims, bbox));
            EAE_TRY(make_map(&map_g_lo, gamma->lo, 3, gdims, bbox));
        }
        if (plan.img_u8) {
            q.conv1 = 1;
            EAE_TRY(make_map_u8(&map_img, plan.img_u8, (uint64_t)plan.img_W, (uint64_t)plan.img_H, n_img));
        }
        static bool attr3_done = false;
        if (!attr3_done) {
            EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes3));
            attr3_done = true;
        }
        static int timing = -1;
        if (timing < 0) { const char* e = getenv("EAE_UMMA_TIMING"); timing = (e && atoi(e) == 1) ? 1 : 0; }
        if (timing) {
            // Debug: per-CTA phase stamps, averaged over the grid, printed to stderr (synchronises the stream).
            long long* d_times = nullptr;
            EAE_CUDA_OK(cudaMalloc(&d_times, (size_t)grid3 * 8 * sizeof(long long)));
            EAE_CUDA_OK(cudaMemsetAsync(d_times, 0, (size_t)grid3 * 8 * sizeof(long long), st));
            q.times = d_times;
            gemm_umma3_kernel<<<grid3, kUmmaThreads3, kSmemBytes3, st>>>(map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, map_img, q);
            EAE_LAUNCH_OK();
            std::vector<long long> h((size_t)grid3 * 8);
            EAE_CUDA_OK(cudaMemcpyAsync(h.data(), d_times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
            EAE_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(d_times);
            for (uint32_t b = 0; b < grid3; b += (grid3 / 3 ? grid3 / 3 : 1))
                fprintf(stderr, "  cta %u raw: %lld | +%lld +%lld +%lld +%lld +%lld +%lld +%lld\n", b, h[(size_t)b * 8],
                        h[(size_t)b * 8 + 1] - h[(size_t)b * 8], h[(size_t)b * 8 + 2] - h[(size_t)b * 8], h[(size_t)b * 8 + 3] - h[(size_t)b * 8],
                        h[(size_t)b * 8 + 4] - h[(size_t)b * 8], h[(size_t)b * 8 + 5] - h[(size_t)b * 8], h[(size_t)b * 8 + 6] - h[(size_t)b * 8],
                        h[(size_t)b * 8 + 7] - h[(size_t)b * 8]);
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t b = 0; b < grid3; b++)
                for (int j = 1; j < 8; j++) acc[j] += (double)(h[(size_t)b * 8 + j] - h[(size_t)b * 8]);
            fprintf(stderr, "umma3 taps %d kchunks %d fuse %d conv1 %d grid %u: setup %.0f first_full %.0f main_issued %.0f acc_seen %.0f "
                            "nrm_seen %.0f staged %.0f end %.0f (avg cycles from CTA start)\n",
                    q.n_taps, q.kchunks, q.fuse, q.conv1, grid3, acc[1] / grid3, acc[2] / grid3, acc[3] / grid3, acc[4] / grid3,
                    acc[5] / grid3, acc[6] / grid3, acc[7] / grid3);
            return 0;
        }
        gemm_umma3_kernel<<<grid3, kUmmaThreads3, kSmemBytes3, st>>>(map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, map_img, q);
        EAE_LAUNCH_OK();
        return 0;
    }
    if (umma_version() == 2) {
        // Cluster size: B tiles are identical for every CTA, so the CTAs of a cluster split each B tile and
        // multicast their slices (env EAE_UMMA_CLUSTER: 1, 2 or 4; default 2 when there are enough tiles).
        static int cs_env = -1;
        if (cs_env < 0) { const char* env = getenv("EAE_UMMA_CLUSTER"); cs_env = env ? atoi(env) : 0; }
        unsigned cs = (cs_env == 1 || cs_env == 2 || cs_env == 4) ? (unsigned)cs_env : 2u;
        if (grid < 2 * cs) cs = 1;
        if (cs > 1) {
            const uint32_t sbox[3] = {kChunkK, kCout / cs, 1};
            EAE_TRY(make_map(&map_b_hi, w.hi, 3, bdims, sbox));
            EAE_TRY(make_map(&map_b_lo, exact3x ? w.lo : w.hi, 3, bdims, sbox));
        }
        UmmaParams2 q;
        memset(&q, 0, sizeof q);
        q.cluster = (int)cs;
        q.n_tiles = (int)grid;
        { static int dbg = -1; if (dbg < 0) { const char* e = getenv("EAE_UMMA_DEBUG"); dbg = e ? atoi(e) : 0; } q.debug = dbg; }
        q.n_taps = p.n_taps; q.kchunks = p.kchunks;
        q.tile_w = p.tile_w; q.tile_h = p.tile_h; q.tiles_x = p.tiles_x; q.tiles_y = p.tiles_y;
        q.Hg = p.Hg; q.Wg = p.Wg;
        q.out = p.out; q.bias = p.bias; q.beta = plan.fuse_beta; q.xin = p.xin;
        q.Hout = p.Hout; q.Wout = p.Wout; q.out_mul = p.out_mul; q.out_r = p.out_r; q.out_s = p.out_s;
        q.out_split = p.out_split;
        q.mode = p.mode;
        q.fuse = plan.fuse;
        q.exact_main = exact3x ? 1 : 0;
        q.error_flag = p.error_flag;
        memcpy(q.taps, p.taps, sizeof q.taps);
        CUtensorMap map_g_hi = map_b_hi, map_g_lo = map_b_lo;
        if (plan.fuse) {
            if (!gamma || !gamma->hi || !gamma->lo || !plan.fuse_beta || plan.mode != kEpiBias) {
                set_error("gemm_umma: fused GDN needs gamma hi/lo, beta and a bias-mode contraction");
                return EAE_ERR_ARGUMENT;
            }
            const uint64_t gdims[3] = {kCout, kCout, 1};
            const uint32_t gbox[3] = {kChunkK, kCout / cs, 1};
            EAE_TRY(make_map(&map_g_hi, gamma->hi, 3, gdims, gbox));
            EAE_TRY(make_map(&map_g_lo, gamma->lo, 3, gdims, gbox));
        }
        static bool attr2_done = false;
        if (!attr2_done) {
            EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
            attr2_done = true;
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(((grid + cs - 1) / cs) * cs);
        cfg.blockDim = dim3(kUmmaThreads2);
        cfg.dynamicSmemBytes = kSmemBytes2;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        EAE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_umma2_kernel, map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, q));
        EAE_LAUNCH_OK();
        return 0;
    }
    if (plan.fuse) { set_error("gemm_umma: fused GDN needs kernel version 2 or 3"); return EAE_ERR_ARGUMENT; }
    if (exact3x) {
        static bool attr_done = false;
        if (!attr_done) {
            EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg<true>::kSmemBytes));
            attr_done = true;
        }
        gemm_umma_kernel<true><<<grid, kUmmaThreads, Cfg<true>::kSmemBytes, st>>>(map_a, map_b_hi, map_b_lo, p);
    } else {
        static bool attr_done = false;
        if (!attr_done) {
            EAE_CUDA_OK(cudaFuncSetAttribute(gemm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg<false>::kSmemBytes));
            attr_done = true;
        }
        gemm_umma_kernel<false><<<grid, kUmmaThreads, Cfg<false>::kSmemBytes, st>>>(map_a, map_b_hi, map_b_lo, p);
    }
    EAE_LAUNCH_OK();
    return 0;
}

int launch_tconv9s4_fused(const float* in, const UmmaWeights& w, uint8_t* out_u8, float* out_f32, uint32_t n, int H, int W,
                          bool exact3x, cudaStream_t st)
{
    if (!n) return 0;
    if (H % 4 != 0 || W % 4 != 0 || !w.hi || (exact3x && !w.lo)) { set_error("tconv9s4: bad arguments"); return EAE_ERR_ARGUMENT; }
    // the gather stores pixel pairs (uchar2 / float2)
    if ((reinterpret_cast<uintptr_t>(out_u8) & 1u) || (reinterpret_cast<uintptr_t>(out_f32) & 7u)) {
        set_error("tconv9s4: the reconstruction buffer must be 2-byte (uint8) / 8-byte (float) aligned");
        return EAE_ERR_ARGUMENT;
    }
    if (!g_error_flag) {
        EAE_CUDA_OK(cudaMalloc(&g_error_flag, 4));
        EAE_CUDA_OK(cudaMemset(g_error_flag, 0, 4));
    }
    const int H1 = H / 4, W1 = W / 4;
    UmmaParams6 q;
    memset(&q, 0, sizeof q);
    q.kchunks = kCout / kChunkK;
    q.tiles_y = (H1 + 1 + kBlkY6 - 1) / kBlkY6;      // pixel blocks 0 .. H1 (the first and the last are half blocks)
    q.tiles_x = (W1 + 1 + kBlkX6 - 1) / kBlkX6;
    q.H = H; q.W = W;
    q.out_u8 = out_u8; q.out_f32 = out_f32;
    q.exact_main = exact3x ? 1 : 0;
    q.error_flag = g_error_flag;
    CUtensorMap map_a, map_b_hi, map_b_lo;
    const uint64_t adims[5] = {(uint64_t)kCout, (uint64_t)W1, (uint64_t)H1, 1, n};
    const uint32_t abox[5] = {kChunkK, 16, 8, 1, 1};
    EAE_TRY(make_map(&map_a, in, 5, adims, abox));
    const uint64_t bdims[3] = {(uint64_t)kCout, kCout, 1};      // [tap column (81 used)][Cin], K-major
    const uint32_t bbox[3] = {kChunkK, 96, 1};
    EAE_TRY(make_map(&map_b_hi, w.hi, 3, bdims, bbox));
    EAE_TRY(make_map(&map_b_lo, exact3x ? w.lo : w.hi, 3, bdims, bbox));
    if (umma_version() >= 7) {
        UmmaParams7 q7;
        memset(&q7, 0, sizeof q7);
        q7.tiles_x = q.tiles_x; q7.tiles_y = q.tiles_y;
        const uint64_t n_tiles = (uint64_t)n * (uint64_t)(q.tiles_x * q.tiles_y);
        if (n_tiles >= (1ull << 31)) { set_error("tconv9s4: batch too large"); return EAE_ERR_ARGUMENT; }
        q7.n_tiles = (int)n_tiles;
        q7.H = H; q7.W = W;
        q7.out_u8 = out_u8; q7.out_f32 = out_f32;
        q7.exact_main = q.exact_main;
        q7.error_flag = g_error_flag;
        static int n_sms = 0;
        if (!n_sms) {
            int dev = 0;
            EAE_CUDA_OK(cudaGetDevice(&dev));
            EAE_CUDA_OK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
            EAE_CUDA_OK(cudaFuncSetAttribute(tconv9s4_umma7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes7));
        }
        const uint32_t grid7 = n_tiles < (uint64_t)n_sms ? (uint32_t)n_tiles : (uint32_t)n_sms;
        tconv9s4_umma7_kernel<<<grid7, kThreads7, kSmemBytes7, st>>>(map_a, map_b_hi, map_b_lo, q7);
        EAE_LAUNCH_OK();
        return 0;
    }
    static bool attr6_done = false;
    if (!attr6_done) {
        EAE_CUDA_OK(cudaFuncSetAttribute(tconv9s4_umma6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes6));
        EAE_CUDA_OK(cudaFuncSetAttribute(tconv9s4_umma6_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr6_done = true;
    }
    const uint64_t grid = (uint64_t)n * (uint64_t)(q.tiles_x * q.tiles_y);
    if (grid >= (1ull << 31)) { set_error("tconv9s4: batch too large"); return EAE_ERR_ARGUMENT; }
    tconv9s4_umma6_kernel<<<(uint32_t)grid, kThreads5, kSmemBytes6, st>>>(map_a, map_b_hi, map_b_lo, q);
    EAE_LAUNCH_OK();
    return 0;
}

}  // namespace eae
