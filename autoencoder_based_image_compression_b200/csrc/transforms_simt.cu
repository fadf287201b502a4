// fp32 CUDA-core path of the transforms (EAE_MATH_FP32_SIMT): the tap-list implicit GEMM of
// conv_plan.cuh with FFMA accumulation, plus the thin-layer helpers (im2col for the 1->128 k9 s4
// convolution, col2im for the 128->1 k9 s4 transposed convolution) and the quantize / dequantize
// kernels that sit between the transforms and the lossless coder.
//
// This is the exact-fp32 parity mode and the numerical yardstick for the tcgen05 path
// (conv_umma.cu). 128 x 128 output tile per CTA, 8 x 8 outputs per thread, K consumed 16 channels
// of one tap at a time through double-buffered shared memory.
#include "common.cuh"
#include "conv_plan.cuh"
#include "internal.cuh"
#include "transforms.cuh"

namespace eae {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, kGemmThreads = 256;
constexpr int kAStride = BM + 4;

__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_simt_kernel(const __grid_constant__ GemmPlan p)
{
    __shared__ __align__(16) float As[2][BK][kAStride];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const uint32_t m0 = blockIdx.x * BM;
    const int tx = tid & 15, ty = tid >> 4;

    // ---- A-operand gather coordinates: this thread loads 4 channels of rows a_row and a_row + 64.
    const int a_row = tid >> 2, a_c4 = tid & 3;
    int iy0[2], ix0[2];
    size_t img_off[2];
    bool row_ok[2];
    #pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint32_t m = m0 + a_row + 64 * r;
        row_ok[r] = m < p.M;
        const uint32_t mm = row_ok[r] ? m : 0;
        const uint32_t per_img = (uint32_t)(p.Hg * p.Wg);
        const uint32_t img = mm / per_img, rem = mm % per_img;
        iy0[r] = (int)(rem / p.Wg) * p.in_mul;
        ix0[r] = (int)(rem % p.Wg) * p.in_mul;
        img_off[r] = (size_t)img * p.Hin * p.Win;
    }
    const int b_k = tid >> 5, b_n4 = tid & 31;

    const int kpt = p.Cin / BK;             // k-steps per tap
    const int total = p.n_taps * kpt;

    float4 ra[2], rb[2];
    auto load_regs = [&](int step) {
        const int t = step / kpt, c0 = (step - t * kpt) * BK;
        const Tap tap = p.taps[t];
        #pragma unroll
        for (int r = 0; r < 2; r++) {
            const int iy = iy0[r] + tap.dy, ix = ix0[r] + tap.dx;
            const bool ok = row_ok[r] && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                size_t pix;
                if (p.in_split)   // img_off is a multiple of Hin * Win = 4 planes of (Hin/2) * (Win/2)
                    pix = img_off[r] + ((size_t)((iy & 1) * 2 + (ix & 1)) * (size_t)(p.Hin / 2) + (size_t)(iy >> 1)) *
                                           (size_t)(p.Win / 2) + (size_t)(ix >> 1);
                else
                    pix = img_off[r] + (size_t)iy * p.Win + ix;
                v = __ldg(reinterpret_cast<const float4*>(p.in + pix * p.Cin + c0 + a_c4 * 4));
            }
            if (p.mode != kEpiBias) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
            ra[r] = v;
            rb[r] = __ldg(reinterpret_cast<const float4*>(
                        p.w + tap.w_off + (size_t)(c0 + b_k + 8 * r) * BN + b_n4 * 4));
        }
    };
    auto store_smem = [&](int buf) {
        #pragma unroll
        for (int r = 0; r < 2; r++) {
            As[buf][a_c4 * 4 + 0][a_row + 64 * r] = ra[r].x;
            As[buf][a_c4 * 4 + 1][a_row + 64 * r] = ra[r].y;
            As[buf][a_c4 * 4 + 2][a_row + 64 * r] = ra[r].z;
            As[buf][a_c4 * 4 + 3][a_row + 64 * r] = ra[r].w;
            *reinterpret_cast<float4*>(&Bs[buf][b_k + 8 * r][b_n4 * 4]) = rb[r];
        }
    };

    float acc[8][8];
    #pragma unroll
    for (int i = 0; i < 8; i++)
        #pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    load_regs(0);
    store_smem(0);
    __syncthreads();
    int buf = 0;
    #pragma unroll 1
    for (int step = 0; step < total; step++) {
        const bool more = step + 1 < total;
        if (more) load_regs(step + 1);
        #pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            #pragma unroll
            for (int i = 0; i < 8; i++)
                #pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store_smem(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue
    float4 bias0 = make_float4(0.f, 0.f, 0.f, 0.f), bias1 = bias0;
    if (p.bias) {
        bias0 = __ldg(reinterpret_cast<const float4*>(p.bias + tx * 4));
        bias1 = __ldg(reinterpret_cast<const float4*>(p.bias + 64 + tx * 4));
    }
    const uint32_t per_img = (uint32_t)(p.Hg * p.Wg);
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        const uint32_t img = m / per_img, rem = m % per_img;
        const int a = (int)(rem / p.Wg), b = (int)(rem % p.Wg);
        const int oy = a * p.out_mul + p.out_r, ox = b * p.out_mul + p.out_s;
        const size_t opix = p.out_split
            ? (((size_t)img * 4 + (size_t)((oy & 1) * 2 + (ox & 1))) * (size_t)(p.Hout / 2) + (size_t)(oy >> 1)) *
                  (size_t)(p.Wout / 2) + (size_t)(ox >> 1)
            : ((size_t)img * p.Hout + (size_t)oy) * p.Wout + (size_t)ox;
        float4 v0 = make_float4(acc[i][0] + bias0.x, acc[i][1] + bias0.y, acc[i][2] + bias0.z, acc[i][3] + bias0.w);
        float4 v1 = make_float4(acc[i][4] + bias1.x, acc[i][5] + bias1.y, acc[i][6] + bias1.z, acc[i][7] + bias1.w);
        if (p.mode != kEpiBias) {
            // GDN / IGDN: the input pixel is the output pixel (1 tap, in_mul = out_mul = 1).
            const float* xin = p.in + (((size_t)img * p.Hin + a) * p.Win + b) * p.Cin;
            const float4 x0 = *reinterpret_cast<const float4*>(xin + tx * 4);
            const float4 x1 = *reinterpret_cast<const float4*>(xin + 64 + tx * 4);
            if (p.mode == kEpiGdn) {
                v0 = make_float4(__fdiv_rn(x0.x, __fsqrt_rn(v0.x)), __fdiv_rn(x0.y, __fsqrt_rn(v0.y)),
                                 __fdiv_rn(x0.z, __fsqrt_rn(v0.z)), __fdiv_rn(x0.w, __fsqrt_rn(v0.w)));
                v1 = make_float4(__fdiv_rn(x1.x, __fsqrt_rn(v1.x)), __fdiv_rn(x1.y, __fsqrt_rn(v1.y)),
                                 __fdiv_rn(x1.z, __fsqrt_rn(v1.z)), __fdiv_rn(x1.w, __fsqrt_rn(v1.w)));
            } else {
                v0 = make_float4(__fmul_rn(x0.x, __fsqrt_rn(v0.x)), __fmul_rn(x0.y, __fsqrt_rn(v0.y)),
                                 __fmul_rn(x0.z, __fsqrt_rn(v0.z)), __fmul_rn(x0.w, __fsqrt_rn(v0.w)));
                v1 = make_float4(__fmul_rn(x1.x, __fsqrt_rn(v1.x)), __fmul_rn(x1.y, __fsqrt_rn(v1.y)),
                                 __fmul_rn(x1.z, __fsqrt_rn(v1.z)), __fmul_rn(x1.w, __fsqrt_rn(v1.w)));
            }
        }
        float* o = p.out + opix * kCout;
        *reinterpret_cast<float4*>(o + tx * 4) = v0;
        *reinterpret_cast<float4*>(o + 64 + tx * 4) = v1;
    }
}

// ---- conv1 helper: uint8 [n, H, W] -> im2col rows [n * H/4 * W/4, 96] fp32 (k = ky * 9 + kx; k >= 81
// is zero padding). TF SAME for k9 s4: 2 before, 3 after (total 5).
__global__ void im2col_k9s4_kernel(const uint8_t* __restrict__ img, float* __restrict__ out,
                                   uint32_t n, int H, int W)
{
    const int H1 = H / 4, W1 = W / 4;
    const uint64_t total = (uint64_t)n * H1 * W1 * kIm2colK;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = (uint32_t)(t % kIm2colK);
        const uint64_t m = t / kIm2colK;
        float v = 0.f;
        if (k < 81) {
            const int b = (int)(m % W1), a = (int)((m / W1) % H1);
            const uint64_t im = m / ((uint64_t)W1 * H1);
            const int iy = 4 * a + (int)(k / 9) - 2, ix = 4 * b + (int)(k % 9) - 2;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = (float)img[(im * H + iy) * W + ix];
        }
        out[t] = v;
    }
}

// ---- tconv3 helper: P [n, H/4, W/4, 128] (column ky * 9 + kx = contribution of that input pixel
// through filter tap (ky, kx)) -> out[oy, ox] = sum over (a, ky) with 4a + ky - 2 = oy, same in x.
// Then tls.cast_bt601 (tools.py:61-93): clip to [16, 235], round half to even, uint8.
__global__ void col2im_k9s4_kernel(const float* __restrict__ P, uint8_t* __restrict__ out_u8,
                                   float* __restrict__ out_f32, uint32_t n, int H, int W)
{
    const int H1 = H / 4, W1 = W / 4;
    const uint64_t total = (uint64_t)n * H * W;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const int ox = (int)(t % W), oy = (int)((t / W) % H);
        const uint64_t im = t / ((uint64_t)W * H);
        float acc = 0.f;
        // ky = oy + 2 - 4a, increasing ky <=> decreasing a
        const int a_hi = (oy + 2) >> 2, b_hi = (ox + 2) >> 2;
        #pragma unroll
        for (int da = 0; da < 3; da++) {
            const int a = a_hi - da, ky = oy + 2 - 4 * a;
            if (a < 0 || a >= H1 || ky > 8) continue;
            #pragma unroll
            for (int db = 0; db < 3; db++) {
                const int b = b_hi - db, kx = ox + 2 - 4 * b;
                if (b < 0 || b >= W1 || kx > 8) continue;
                acc += __ldg(P + ((im * H1 + a) * W1 + b) * kCout + ky * 9 + kx);
            }
        }
        if (out_f32) out_f32[t] = acc;
        if (out_u8) out_u8[t] = (uint8_t)(int)rintf(fminf(fmaxf(acc, 16.f), 235.f));
    }
}

// ---- quantizer between the transforms and the coder (reconstructing_eae_kodak.py:170-192,
// tools.py:927-929, compression.py:142): c = y - mean; k = rint(c / delta) (int16, planar
// [n, C, hw] for the coder); q_off = delta * k + mean (fp32 NHWC, the decoder input).
// One CTA: 32 pixels x 128 maps through a shared-memory transpose.
__global__ void __launch_bounds__(256)
quantize_to_planar_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                          const float* __restrict__ delta, int16_t* __restrict__ idx_planar,
                          float* __restrict__ q_off, uint32_t hw, uint32_t* __restrict__ flag)
{
    __shared__ int16_t tile[kCout][34];
    const uint32_t img = blockIdx.y, p0 = blockIdx.x * 32;
    const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
    const float d = __ldg(delta + c), mu = mean ? __ldg(mean + c) : 0.f;
    bool bad = false;
    for (int i = half; i < 32; i += 2) {
        const uint32_t pix = p0 + i;
        int16_t k = 0;
        if (pix < hw) {
            const size_t at = ((size_t)img * hw + pix) * kCout + c;
            const float r = rintf(__fdiv_rn(__fsub_rn(y[at], mu), d));
            if (!(fabsf(r) < 32768.f)) bad = true; else k = (int16_t)(int)r;
            if (q_off) q_off[at] = __fadd_rn(__fmul_rn(d, (float)k), mu);
        }
        tile[c][i] = k;
    }
    if (bad) atomicOr(flag, 1u);
    __syncthreads();
    // 256 threads: each writes one int16 of 8 (map, pixel) rows per pass
    const int px = threadIdx.x & 31, row = threadIdx.x >> 5;
    for (int cc = row; cc < kCout; cc += 8) {
        const uint32_t pix = p0 + px;
        if (pix < hw) idx_planar[((size_t)img * kCout + cc) * hw + pix] = tile[cc][px];
    }
}

// idx planar [n, C, hw] -> q_off NHWC fp32 = delta[c] * k + mean[c]
__global__ void __launch_bounds__(256)
dequantize_from_planar_kernel(const int16_t* __restrict__ idx_planar, const float* __restrict__ mean,
                              const float* __restrict__ delta, float* __restrict__ q_off, uint32_t hw)
{
    __shared__ int16_t tile[kCout][34];
    const uint32_t img = blockIdx.y, p0 = blockIdx.x * 32;
    const int px = threadIdx.x & 31, row = threadIdx.x >> 5;
    for (int cc = row; cc < kCout; cc += 8) {
        const uint32_t pix = p0 + px;
        tile[cc][px] = pix < hw ? idx_planar[((size_t)img * kCout + cc) * hw + pix] : (int16_t)0;
    }
    __syncthreads();
    const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
    const float d = __ldg(delta + c), mu = mean ? __ldg(mean + c) : 0.f;
    for (int i = half; i < 32; i += 2) {
        const uint32_t pix = p0 + i;
        if (pix < hw)
            q_off[((size_t)img * hw + pix) * kCout + c] = __fadd_rn(__fmul_rn(d, (float)tile[c][i]), mu);
    }
}

}  // namespace

int launch_gemm_simt(const GemmPlan& plan, cudaStream_t st)
{
    if (plan.M == 0) return 0;
    if (plan.Cin % BK != 0 || plan.n_taps < 1 || plan.n_taps > kMaxTaps) {
        set_error("gemm_simt: bad plan (Cin %d, taps %d)", plan.Cin, plan.n_taps);
        return EAE_ERR_ARGUMENT;
    }
    gemm_simt_kernel<<<ceil_div_u32(plan.M, BM), kGemmThreads, 0, st>>>(plan);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_im2col_k9s4(const uint8_t* img, float* out, uint32_t n, int H, int W, cudaStream_t st)
{
    const uint64_t total = (uint64_t)n * (H / 4) * (W / 4) * kIm2colK;
    if (!total) return 0;
    const uint64_t blocks = (total + 255) / 256;
    im2col_k9s4_kernel<<<(uint32_t)(blocks > 148ull * 64 ? 148ull * 64 : blocks), 256, 0, st>>>(img, out, n, H, W);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_col2im_k9s4(const float* P, uint8_t* out_u8, float* out_f32, uint32_t n, int H, int W,
                       cudaStream_t st)
{
    const uint64_t total = (uint64_t)n * H * W;
    if (!total) return 0;
    const uint64_t blocks = (total + 255) / 256;
    col2im_k9s4_kernel<<<(uint32_t)(blocks > 148ull * 64 ? 148ull * 64 : blocks), 256, 0, st>>>(
        P, out_u8, out_f32, n, H, W);
    EAE_LAUNCH_OK();
    return 0;
}

// (see prefer_max_shared in common.cuh)
static void glue_carveouts()
{
    static uint64_t seen = 0;
    if (!first_use_on_device(&seen)) return;
    prefer_max_shared(quantize_to_planar_kernel); prefer_max_shared(dequantize_from_planar_kernel);
    prefer_max_shared(col2im_k9s4_kernel); prefer_max_shared(im2col_k9s4_kernel);
}

int launch_quantize_to_planar(const float* y, const float* mean, const float* delta, int16_t* idx_planar,
                              float* q_off, uint32_t n, uint32_t hw, uint32_t* flag, cudaStream_t st)
{
    if (!n || !hw) return 0;
    glue_carveouts();
    if (n > 65535u) { set_error("quantize: batch %u exceeds grid.y", n); return EAE_ERR_ARGUMENT; }
    quantize_to_planar_kernel<<<dim3(ceil_div_u32(hw, 32), n), 256, 0, st>>>(y, mean, delta, idx_planar,
                                                                            q_off, hw, flag);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_dequantize_from_planar(const int16_t* idx_planar, const float* mean, const float* delta,
                                  float* q_off, uint32_t n, uint32_t hw, cudaStream_t st)
{
    if (!n || !hw) return 0;
    glue_carveouts();
    if (n > 65535u) { set_error("dequantize: batch %u exceeds grid.y", n); return EAE_ERR_ARGUMENT; }
    dequantize_from_planar_kernel<<<dim3(ceil_div_u32(hw, 32), n), 256, 0, st>>>(idx_planar, mean, delta,
                                                                                q_off, hw);
    EAE_LAUNCH_OK();
    return 0;
}

}  // namespace eae
