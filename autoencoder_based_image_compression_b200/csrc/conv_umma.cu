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
#include "umma_last_layer.cuh"      // pulls in umma_common / v3 / v4

namespace eae {

namespace {

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
// Output map of a fused tile's TMA stores (OutGeom4::map_out, umma_v3.cuh): *use = false when this output geometry keeps
// the per-warp stores (quantizer, parity-split phases, odd sizes, EAE_NO_TMA_STORE, or a driver that refuses the map).
int make_out_map(CUtensorMap* map, const GemmPlan& plan, uint64_t n_img, bool* use)
{
    const char* env = getenv("EAE_NO_TMA_STORE");      // (read per launch: the tests switch it inside one process)
    const bool disabled = env && atoi(env);
    *use = false;
    if (disabled || !plan.fuse || plan.quant_idx || !plan.out) return 0;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return 0;
    const uint64_t px = (uint64_t)kCout * sizeof(float);      // one pixel: 128 channels
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bdim[5], estride[5] = {1, 1, 1, 1, 1};
    const void* base = plan.out;
    int rank;
    if (plan.out_split) {
        // [n][plane (y & 1) * 2 + (x & 1)][Hout / 2][Wout / 2][128]; a tile row of the position grid alternates between two planes
        if (plan.out_mul != 1 || (plan.Hout & 1) || (plan.Wout & 1) || plan.Hg != plan.Hout || plan.Wg != plan.Wout) return 0;
        const uint64_t plane = (uint64_t)(plan.Hout / 2) * (uint64_t)(plan.Wout / 2) * px;
        rank = 5;
        gdim[0] = kCout; gdim[1] = 4; gdim[2] = (cuuint64_t)(plan.Wout / 2); gdim[3] = (cuuint64_t)(plan.Hout / 2); gdim[4] = n_img;
        gstride[0] = plane; gstride[1] = px; gstride[2] = (uint64_t)(plan.Wout / 2) * px; gstride[3] = 4 * plane;
        bdim[0] = 32; bdim[1] = 2; bdim[2] = 8; bdim[3] = 4; bdim[4] = 1;
    } else {
        // position (a, b) -> pixel (a * out_mul + out_r, b * out_mul + out_s): a strided view from the phase's first pixel
        const int m = plan.out_mul;
        if (m < 1 || plan.out_r >= plan.Hout || plan.out_s >= plan.Wout) return 0;
        const int wv = (plan.Wout - plan.out_s + m - 1) / m, hv = (plan.Hout - plan.out_r + m - 1) / m;
        base = plan.out + ((size_t)plan.out_r * plan.Wout + (size_t)plan.out_s) * kCout;
        rank = 4;
        gdim[0] = kCout; gdim[1] = (cuuint64_t)(wv < plan.Wg ? wv : plan.Wg); gdim[2] = (cuuint64_t)(hv < plan.Hg ? hv : plan.Hg);
        gdim[3] = n_img;
        gstride[0] = (uint64_t)m * px; gstride[1] = (uint64_t)m * (uint64_t)plan.Wout * px;
        gstride[2] = (uint64_t)plan.Hout * (uint64_t)plan.Wout * px;
        bdim[0] = 32; bdim[1] = 16; bdim[2] = 8; bdim[3] = 1;
    }
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bdim,
                          estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    *use = r == CUDA_SUCCESS;
    return 0;
}
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
        if (plan.n_taps != 1 || plan.Cin != 96 || plan.Hg * 4 != plan.img_H ||
            plan.Wg * 4 != plan.img_W || plan.img_W % 16 != 0) {
            set_error("gemm_umma: the fused k9 s4 convolution needs a 1-tap plan and a [n, 4 Hg, 4 Wg] image");
            return EAE_ERR_ARGUMENT;
        }
        map_a = map_b_hi;      // not read
    } else if (plan.dequant_idx) {
        if (plan.n_taps != 1 || plan.Cin != 128 || plan.Hg != 1 || plan.fuse || plan.mode != kEpiIgdn || !plan.dequant_delta ||
            plan.dequant_hw <= 0 || plan.in_split || plan.out_split || plan.out_mul != 1) {
            set_error("gemm_umma: the fused dequantizer needs a standalone IGDN over a flat position grid");
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
        q.idx_out = plan.quant_idx; q.q_mean = plan.quant_mean; q.q_delta = plan.quant_delta; q.q_flag = plan.quant_flag;
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
        return pl.n_taps > 1 && pl.n_taps <= kMaxTaps && pl.Hg > 1 && pl.Cin == 128 && !pl.img_u8 &&
               pl.in == plan.in && pl.Hg == plan.Hg && pl.Wg == plan.Wg && pl.in_split == plan.in_split && pl.M == plan.M &&
               pl.fuse == plan.fuse && pl.fuse_single_pass == plan.fuse_single_pass && pl.fuse_precise == plan.fuse_precise;
    };
    UmmaParams4x qx;
    memset(&qx, 0, sizeof qx);
    bool all4 = eligible4(plan) && build4(plan, qx.ph[0]);
    for (int i = 0; i < n_more && all4; i++) all4 = eligible4(more[i]) && build4(more[i], qx.ph[1 + i]);
    if (plan.quant_idx && (!all4 || plan.out_split || plan.out_mul != 1 || !plan.quant_delta || !plan.quant_flag)) {
        set_error("gemm_umma: the fused quantizer needs a multi-tap layer with a natural NHWC output grid");
        return EAE_ERR_ARGUMENT;
    }
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
            OutMaps4 maps_out;
            memset(&maps_out, 0, sizeof maps_out);
            for (int ph = 0; ph < qx.n_phases; ph++) {
                bool use = false;
                EAE_TRY(make_out_map(&maps_out.m[ph], ph ? more[ph - 1] : plan, n_img, &use));
                qx.ph[ph].tma_out = use ? 1 : 0;
            }
            // four instantiations: {MUFU, IEEE} normalisation x {fp32 pixels, quantizer} store
            typedef void (*Kernel4)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                    const OutMaps4, const UmmaParams4x);
            static const Kernel4 kernels4[4] = {gemm_umma4_kernel<false, false>, gemm_umma4_kernel<true, false>,
                                                gemm_umma4_kernel<false, true>, gemm_umma4_kernel<true, true>};
            static bool attr4_done = false;
            if (!attr4_done) {
                for (Kernel4 k : kernels4) EAE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
                attr4_done = true;
            }
            const Kernel4 kernel4 = kernels4[(plan.fuse && plan.fuse_precise ? 1 : 0) + (plan.quant_idx ? 2 : 0)];
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
                    EAE_CUDA_OK(cudaMalloc(&d_times, d_times_cap * kStamps4 * sizeof(long long)));
                }
                EAE_CUDA_OK(cudaMemsetAsync(d_times, 0, (size_t)grid4 * kStamps4 * sizeof(long long), st));
                for (int i = 0; i < qx.n_phases; i++) qx.ph[i].times = d_times;
            }
            kernel4<<<grid4, kUmmaThreads4, kSmemBytes4, st>>>(map_u, map_b_hi, map_b_lo, map_g_hi, map_g_lo, maps_out, qx);
            EAE_LAUNCH_OK();
            if (timing4) {
                std::vector<long long> h((size_t)grid4 * kStamps4);
                EAE_CUDA_OK(cudaMemcpyAsync(h.data(), d_times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
                EAE_CUDA_OK(cudaStreamSynchronize(st));
                double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (uint32_t b = 0; b < grid4; b++)
                    for (int j = 1; j < 8; j++) acc[j] += (double)(h[(size_t)b * kStamps4 + j] - h[(size_t)b * kStamps4]);
                std::vector<long long> dur(grid4);
                for (uint32_t b = 0; b < grid4; b++) dur[b] = h[(size_t)b * kStamps4 + 7] - h[(size_t)b * kStamps4];
                std::sort(dur.begin(), dur.end());
                // which SMs ran the CTAs, and when (global timer): makespan against the busy time of the SMs
                {
                    int per_sm[256] = {0};
                    long long t_min = h[8], t_max = h[9];
                    double busy_ns = 0.;
                    for (uint32_t b = 0; b < grid4; b++) {
                        per_sm[h[(size_t)b * kStamps4 + 10] & 255]++;
                        if (h[(size_t)b * kStamps4 + 8] < t_min) t_min = h[(size_t)b * kStamps4 + 8];
                        if (h[(size_t)b * kStamps4 + 9] > t_max) t_max = h[(size_t)b * kStamps4 + 9];
                        busy_ns += (double)(h[(size_t)b * kStamps4 + 9] - h[(size_t)b * kStamps4 + 8]);
                    }
                    int used = 0, most = 0, least = 1 << 30;
                    for (int i = 0; i < 256; i++) if (per_sm[i]) { used++; most = per_sm[i] > most ? per_sm[i] : most; least = per_sm[i] < least ? per_sm[i] : least; }
                    fprintf(stderr, "umma4 grid %u: %d SMs used, %d..%d CTAs per SM, makespan %.1f us, CTA time summed / 148 = %.1f us\n",
                            grid4, used, least, most, (double)(t_max - t_min) * 1e-3, busy_ns * 1e-3 / 148.);
                }
                {
                    // one conversion (set 0) and one MMA issue in the steady state of the main loop (iteration 8)
                    double cv[6] = {0, 0, 0, 0, 0, 0}, mm[4] = {0, 0, 0, 0};
                    uint32_t nb = 0;
                    for (uint32_t b = 0; b < grid4; b++) {
                        const long long* t = &h[(size_t)b * kStamps4];
                        if (!t[12] || !t[18]) continue;
                        nb++;
                        for (int j = 0; j < 6; j++) cv[j] += (double)(t[12 + j] - t[12]);
                        for (int j = 0; j < 4; j++) mm[j] += (double)(t[18 + j] - t[12]);
                    }
                    if (nb)
                        fprintf(stderr, "umma4 iteration 8 (cycles from the conversion's loop top): union seen %.0f, rows read %.0f, slot free %.0f, "
                                        "slot written %.0f, arrived %.0f | MMA warp: loop top %.0f, slot seen %.0f, weights seen %.0f, issued %.0f\n",
                                cv[1] / nb, cv[2] / nb, cv[3] / nb, cv[4] / nb, cv[5] / nb, mm[0] / nb, mm[1] / nb, mm[2] / nb, mm[3] / nb);
                }
                if (q.fuse) {
                    // the fused tail, cycles from the moment conversion warp 2 saw the accumulators complete
                    double cv[11] = {0}, mm[8] = {0};
                    uint32_t nb = 0;
                    for (uint32_t b = 0; b < grid4; b++) {
                        const long long* t = &h[(size_t)b * kStamps4];
                        if (!t[4] || !t[24 + 10]) continue;
                        nb++;
                        for (int j = 0; j < 11; j++) cv[j] += (double)(t[24 + j] - t[4]);
                        for (int j = 0; j < 8; j++) mm[j] += (double)(t[24 + 12 + j] - t[4]);
                    }
                    if (nb)
                        fprintf(stderr, "umma4 tail (cycles from accumulators seen): conversions %.0f %.0f, x0 copied %.0f, conversions %.0f %.0f, "
                                        "norm 0 seen %.0f, half 0 normalised %.0f, barrier %.0f, stored %.0f, half 1 staged %.0f | MMA steps issued "
                                        "%.0f %.0f %.0f %.0f %.0f %.0f %.0f %.0f\n",
                                cv[0] / nb, cv[1] / nb, cv[2] / nb, cv[3] / nb, cv[4] / nb, cv[5] / nb, cv[6] / nb, cv[7] / nb, cv[8] / nb, cv[10] / nb,
                                mm[0] / nb, mm[1] / nb, mm[2] / nb, mm[3] / nb, mm[4] / nb, mm[5] / nb, mm[6] / nb, mm[7] / nb);
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
    {
        UmmaParams3 q;
        memset(&q, 0, sizeof q);
        q.n_taps = p.n_taps; q.kchunks = p.kchunks;
        q.tile_w = p.tile_w; q.tile_h = p.tile_h;
        q.tile_w_log2 = p.tile_w == 128 ? 7 : 4;
        if (p.tile_h == 1) { q.half_da = 0; q.half_db = p.tile_w; } else { q.half_da = p.tile_h; q.half_db = 0; }
        q.tiles_x = (plan.Wg + q.tile_w + q.half_db - 1) / (q.tile_w + q.half_db);
        q.tiles_y = (plan.Hg + q.tile_h + q.half_da - 1) / (q.tile_h + q.half_da);
        q.Hg = p.Hg; q.Wg = p.Wg;
        q.out = p.out; q.bias = p.bias; q.beta = plan.fuse_beta; q.xin = p.xin;
        q.idx_in = plan.dequant_idx; q.dq_mean = plan.dequant_mean; q.dq_delta = plan.dequant_delta; q.hw_in = plan.dequant_hw;
        q.Hout = p.Hout; q.Wout = p.Wout; q.out_mul = p.out_mul; q.out_r = p.out_r; q.out_s = p.out_s;
        q.out_split = p.out_split;
        q.mode = p.mode;
        q.fuse = plan.fuse;
        q.exact_main = exact3x ? 1 : 0;
        q.exact_gdn = plan.fuse_single_pass ? 0 : 1;
        q.error_flag = p.error_flag;
        memcpy(q.taps, p.taps, sizeof q.taps);
        const uint32_t grid3 = n_img * (uint32_t)(q.tiles_x * q.tiles_y);
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
        CUtensorMap map_out = map_b_hi;
        {
            bool use = false;
            EAE_TRY(make_out_map(&map_out, plan, n_img, &use));
            q.tma_out = use ? 1 : 0;
        }
        typedef void (*Kernel3)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                const CUtensorMap, const CUtensorMap, const UmmaParams3);
        static const Kernel3 kernels3[3] = {gemm_umma3_kernel<false, false>, gemm_umma3_kernel<true, false>,
                                            gemm_umma3_kernel<false, true>};
        static bool attr3_done = false;
        if (!attr3_done) {
            for (Kernel3 k : kernels3) EAE_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes3));
            attr3_done = true;
        }
        const Kernel3 kernel3 = kernels3[plan.dequant_idx ? 2 : (plan.fuse && plan.fuse_precise ? 1 : 0)];
        static int timing = -1;
        if (timing < 0) { const char* e = getenv("EAE_UMMA_TIMING"); timing = (e && atoi(e) == 1) ? 1 : 0; }
        if (timing) {
            // Debug: per-CTA phase stamps, averaged over the grid, printed to stderr (synchronises the stream).
            long long* d_times = nullptr;
            EAE_CUDA_OK(cudaMalloc(&d_times, (size_t)grid3 * kStamps3 * sizeof(long long)));
            EAE_CUDA_OK(cudaMemsetAsync(d_times, 0, (size_t)grid3 * kStamps3 * sizeof(long long), st));
            q.times = d_times;
            kernel3<<<grid3, kUmmaThreads3, kSmemBytes3, st>>>(map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, map_img, map_out, q);
            EAE_LAUNCH_OK();
            std::vector<long long> h((size_t)grid3 * kStamps3);
            EAE_CUDA_OK(cudaMemcpyAsync(h.data(), d_times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
            EAE_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(d_times);
            for (uint32_t b = 0; b < grid3; b += (grid3 / 3 ? grid3 / 3 : 1))
                fprintf(stderr, "  cta %u raw: %lld | +%lld +%lld +%lld +%lld +%lld +%lld +%lld\n", b, h[(size_t)b * kStamps3],
                        h[(size_t)b * kStamps3 + 1] - h[(size_t)b * kStamps3], h[(size_t)b * kStamps3 + 2] - h[(size_t)b * kStamps3], h[(size_t)b * kStamps3 + 3] - h[(size_t)b * kStamps3],
                        h[(size_t)b * kStamps3 + 4] - h[(size_t)b * kStamps3], h[(size_t)b * kStamps3 + 5] - h[(size_t)b * kStamps3], h[(size_t)b * kStamps3 + 6] - h[(size_t)b * kStamps3],
                        h[(size_t)b * kStamps3 + 7] - h[(size_t)b * kStamps3]);
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t b = 0; b < grid3; b++)
                for (int j = 1; j < 8; j++) acc[j] += (double)(h[(size_t)b * kStamps3 + j] - h[(size_t)b * kStamps3]);
            fprintf(stderr, "umma3 taps %d kchunks %d fuse %d conv1 %d grid %u: setup %.0f first_full %.0f main_issued %.0f acc_seen %.0f "
                            "nrm_seen %.0f staged %.0f end %.0f (avg cycles from CTA start)\n",
                    q.n_taps, q.kchunks, q.fuse, q.conv1, grid3, acc[1] / grid3, acc[2] / grid3, acc[3] / grid3, acc[4] / grid3,
                    acc[5] / grid3, acc[6] / grid3, acc[7] / grid3);
            if (q.fuse && q.tma_out) {      // the store issuer (producer thread): cycles from CTA start
                double is[6] = {0, 0, 0, 0, 0, 0};
                for (uint32_t b = 0; b < grid3; b++)
                    for (int j = 0; j < 6; j++) is[j] += (double)(h[(size_t)b * kStamps3 + 8 + j] - h[(size_t)b * kStamps3]);
                fprintf(stderr, "umma3 store issuer: rounds seen ready %.0f %.0f %.0f %.0f, last stores issued %.0f, sources read %.0f\n",
                        is[0] / grid3, is[1] / grid3, is[2] / grid3, is[3] / grid3, is[4] / grid3, is[5] / grid3);
            }
            return 0;
        }
        kernel3<<<grid3, kUmmaThreads3, kSmemBytes3, st>>>(map_a, map_b_hi, map_b_lo, map_g_hi, map_g_lo, map_img, map_out, q);
        EAE_LAUNCH_OK();
        return 0;
    }
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
    const int tiles_y = (H1 + 1 + kBlkY6 - 1) / kBlkY6;      // pixel blocks 0 .. H1 (the first and the last are half blocks)
    const int tiles_x = (W1 + 1 + kBlkX6 - 1) / kBlkX6;
    CUtensorMap map_a, map_b_hi, map_b_lo;
    const uint64_t adims[5] = {(uint64_t)kCout, (uint64_t)W1, (uint64_t)H1, 1, n};
    const uint32_t abox[5] = {kChunkK, 16, 8, 1, 1};
    EAE_TRY(make_map(&map_a, in, 5, adims, abox));
    const uint64_t bdims[3] = {(uint64_t)kCout, kCout, 1};      // [tap column (81 used)][Cin], K-major
    const uint32_t bbox[3] = {kChunkK, 96, 1};
    EAE_TRY(make_map(&map_b_hi, w.hi, 3, bdims, bbox));
    EAE_TRY(make_map(&map_b_lo, exact3x ? w.lo : w.hi, 3, bdims, bbox));
    UmmaParams7 q7;
    memset(&q7, 0, sizeof q7);
    q7.tiles_x = tiles_x; q7.tiles_y = tiles_y;
    const uint64_t n_tiles = (uint64_t)n * (uint64_t)(tiles_x * tiles_y);
    if (n_tiles >= (1ull << 31)) { set_error("tconv9s4: batch too large"); return EAE_ERR_ARGUMENT; }
    q7.n_tiles = (int)n_tiles;
    q7.H = H; q7.W = W;
    q7.out_u8 = out_u8; q7.out_f32 = out_f32;
    q7.exact_main = exact3x ? 1 : 0;
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

namespace {

// Debug hook: sqrt_rn_norm / div_rn_norm and the packed norm_apply2 against the compiler's IEEE sqrt.rn / div.rn / mul.rn.
//   square roots: EVERY float in [2^-20, 2^40] (a norm + beta lies in [2e-5, ~1e6]), each also through the packed GDN and
//   IGDN normalisation with two pseudo-random x
//   quotients:    n_pairs pseudo-random pairs, |a| in [2^-30, 2^30] (either sign), b in [2^-10, 2^20]
__global__ void norm_arith_check_kernel(uint64_t n_pairs, unsigned long long* __restrict__ bad)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad_sqrt = 0, bad_div = 0;
    const uint32_t lo = 0x35800000u /* 2^-20 */, hi = 0x53800000u /* 2^40 */;
    for (uint64_t b = lo + tid; b <= hi; b += nth) {
        const float n = __uint_as_float((uint32_t)b);
        const float s = __fsqrt_rn(n);
        if (__float_as_uint(sqrt_rn_norm(n)) != __float_as_uint(s)) bad_sqrt++;
        // the packed (fp32x2) normalisation of the fused tails against the compiler's IEEE operations: two pseudo-random
        // x per n, |x| in [2^-30, 2^30], either sign, as a GDN (x / sqrt(n)) and as an IGDN (x * sqrt(n))
        uint64_t z = b * 0x9E3779B97F4A7C15ull + 0x7654321ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        const uint32_t u0 = (uint32_t)z, u1 = (uint32_t)(z >> 32);
        const float x0 = __uint_as_float((u0 & 0x807FFFFFu) | ((127u - 30u + (u0 >> 23) % 61u) << 23));
        const float x1 = __uint_as_float((u1 & 0x807FFFFFu) | ((127u - 30u + (u1 >> 23) % 61u) << 23));
        const float2 g = norm_apply2<true>(x0, x1, n, n, 1), ig = norm_apply2<true>(x0, x1, n, n, 2);
        if (__float_as_uint(g.x) != __float_as_uint(__fdiv_rn(x0, s)) || __float_as_uint(g.y) != __float_as_uint(__fdiv_rn(x1, s))) bad_div++;
        if (__float_as_uint(ig.x) != __float_as_uint(__fmul_rn(x0, s)) || __float_as_uint(ig.y) != __float_as_uint(__fmul_rn(x1, s))) bad_sqrt++;
    }
    for (uint64_t i = tid; i < n_pairs; i += nth) {
        uint64_t z = i * 0x9E3779B97F4A7C15ull + 0x1234567ull;      // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        const uint32_t ua = (uint32_t)z, ub = (uint32_t)(z >> 32);
        const uint32_t ea = 127u - 30u + (ua >> 23) % 61u, eb = 127u - 10u + (ub >> 23) % 31u;
        const float a = __uint_as_float((ua & 0x807FFFFFu) | (ea << 23)), b = __uint_as_float((ub & 0x007FFFFFu) | (eb << 23));
        if (__float_as_uint(div_rn_norm(a, b)) != __float_as_uint(__fdiv_rn(a, b))) bad_div++;
    }
    if (bad_sqrt) atomicAdd(bad, bad_sqrt);
    if (bad_div) atomicAdd(bad + 1, bad_div);
}

}  // namespace
}  // namespace eae

extern "C" int eae_debug_check_norm_arithmetic(uint64_t n_pairs, uint64_t* sqrt_mismatches, uint64_t* div_mismatches)
{
    using namespace eae;
    if (!sqrt_mismatches || !div_mismatches) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    unsigned long long* d = nullptr;
    EAE_CUDA_OK(cudaMalloc(&d, 16));
    EAE_CUDA_OK(cudaMemset(d, 0, 16));
    norm_arith_check_kernel<<<148 * 8, 256>>>(n_pairs, d);
    unsigned long long h[2] = {0, 0};
    const cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    EAE_CUDA_OK(e);
    *sqrt_mismatches = h[0]; *div_mismatches = h[1];
    return 0;
}
