// Constants, parameter blocks and PTX wrappers shared by the tcgen05 kernels (conv_umma.cu).
// Private to conv_umma.cu (one translation unit): everything here sits in its anonymous namespace.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "common.cuh"
#include "conv_plan.cuh"

namespace eae {
namespace {


constexpr int kTileM = 128;
constexpr int kChunkK = 32;                 // fp32 elements per 128-byte swizzle row
constexpr int kTileBytes = kTileM * 128;    // 16 KB: 128 rows x 128 bytes
constexpr uint32_t kTmemCols2 = 512;              // every kernel here owns the whole TMEM of its SM (one CTA per SM)
// A barrier wait gives up after ~2 s. (It was 0.2 s until one bench run in about sixty, the first process on a fresh box,
// reported a time-out of all three main-loop roles of a CTA in the middle of an otherwise normal run, while 4 800 steps of
// scripts/soak.py and every other run passed: with the kernels of the library still being loaded lazily and twelve
// streams of graph instantiations in flight, a running CTA can apparently be held for longer than that. A genuine
// dead-lock is still reported, ten times later.)
constexpr long long kTimeoutCycles = 4000ll * 1000 * 1000;

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
// Bulk tensor stores, shared -> global (bulk async-group completion): the writing threads make their generic-proxy
// shared-memory writes visible to the async proxy (tma_store_fence), synchronise, then ONE thread issues the stores and
// commits the group; before the CTA may exit (or the source is rewritten) that thread waits until the source has been read.
__device__ __forceinline__ void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

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

// Parameters of gemm_umma3_kernel (layer 1 and the standalone GDN / IGDN).
struct UmmaParams3 {
    int n_taps, kchunks;
    int tile_w, tile_h, tiles_x, tiles_y;
    int Hg, Wg;
    float* out;
    const float* bias;
    const float* beta;       // fused GDN / IGDN
    const float* xin;        // standalone GDN / IGDN: the un-squared input
    // standalone IGDN on the decoder's input with the dequantizer fused into the operand load
    // (reconstructing_eae_kodak.py:192, components.py:56-58): when idx_in is set the input is
    // delta[c] * k + mean[c] of the planar int16 indices [n, 128, hw_in] the entropy decoder wrote; `xin` is not read
    const int16_t* idx_in;
    const float* dq_mean;    // [128] or NULL
    const float* dq_delta;   // [128]
    int hw_in;
    int Hout, Wout, out_mul, out_r, out_s, out_split;
    int tma_out;             // fused tiles leave through TMA stores (the kernel's map_out; see OutGeom4)
    int mode;                // EpilogueMode of a standalone launch
    int fuse;                // 0 none, 1 GDN, 2 IGDN after the contraction
    int exact_main;          // 3xTF32 for the main contraction
    int exact_gdn;           // 3xTF32 for the fused norm (versions 3 and 4)
    int tile_w_log2;
    int half_da, half_db;    // version 3: offset of the second 128-row half of a tile in the position grid
    int conv1;               // version 3: A rows are the 9x9 patches (k9 s4) of a uint8 image tile staged by TMA
    long long* times;        // version 3, env EAE_UMMA_TIMING=1: [grid][8] clock64 stamps of the phases of each CTA
    uint32_t* error_flag;
    UmmaTap taps[kMaxTaps];
};


// A operand from tensor memory, B from shared memory.
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

}  // namespace
}  // namespace eae
