// The last layer (transposed k9 s4 convolution) with its col2im gather and the BT.601 cast in the kernel.
// Private to conv_umma.cu (one translation unit): everything here sits in its anonymous namespace.
#pragma once

#include "umma_v4.cuh"

namespace eae {
namespace {

// =================================================================================================
// The LAST layer (conv2d_transpose k9 s4, 128 -> 1, components.py:79-84) with its col2im gather and the
// BT.601 cast (tools.py:61-93) inside the kernel.
//
// Measured on the first tensor version of this layer (profiles/r01_ncu_full_gemm_layers_final.md): the per-position tap matrix [positions, 128] is
// written to HBM (257 MB per 24 images) only to be read back by col2im_k9s4_kernel (another 302 MB + 94 us), for 9 MB
// of pixels. Here a CTA contracts tiles of 8 x 16 positions (thread = position = TMEM lane), dumps the 81 tap columns of its accumulator to shared memory and gathers the pixels of the 6 x 14 pixel
// blocks whose contributing positions all lie inside the tile: pixel row oy = 4 q + r - 2 (block q, r in [0, 4))
// receives position q through ky = r, q - 1 through ky = r + 4 and, for r = 0, q - 2 through ky = 8 - so blocks
// [q0, q0 + 6) need positions [q0 - 2, q0 + 6). Tiles overlap by two positions (65 % of the contracted rows are
// new); positions outside the layer's input are zero-filled by TMA and contribute exact zeros, which is what the
// skip in col2im_k9s4_kernel amounts to. The sum runs in that kernel's order, so the two paths agree bit for bit.
// MMA N = 96 (81 taps used). HBM traffic: the activations once (the overlap is served by L2) + the pixels.
constexpr int kColStride6 = 87;                    // odd: the thread-per-row dump is conflict-free
constexpr int kBlkY6 = 6, kBlkX6 = 14;             // pixel blocks (4 x 4 pixels) a tile completes
constexpr uint32_t kInstrDescN96 = (1u << 4) | (2u << 7) | (2u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ts_n96(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kInstrDescN96), "r"(accumulate)
        : "memory");
}

// =================================================================================================
// The kernel is PERSISTENT and fully pipelined (one CTA per SM, tiles round-robin).
//
// Measured on a one-tile-per-CTA version (two 128-position CTAs per SM): 170 us per 24 images = 13 k cycles per CTA, of which the
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

}  // namespace
}  // namespace eae
