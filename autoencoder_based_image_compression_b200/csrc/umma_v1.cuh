// Version 1 of the tcgen05 tap-list GEMM: both operands from shared memory, no fusion (EAE_UMMA_VERSION=1).
// Private to conv_umma.cu (one translation unit): everything here sits in its anonymous namespace.
#pragma once

#include "umma_common.cuh"

namespace eae {
namespace {

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

}  // namespace
}  // namespace eae
