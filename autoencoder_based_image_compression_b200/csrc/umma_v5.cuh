// Version 5: thin layers on two half-size CTAs per SM (EAE_UMMA_VERSION=5, EAE_UMMA_V5_FUSED).
// Private to conv_umma.cu (one translation unit): everything here sits in its anonymous namespace.
#pragma once

#include "umma_v4.cuh"

namespace eae {
namespace {

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

}  // namespace
}  // namespace eae
