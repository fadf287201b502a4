// Launchers of the transform kernels (transforms_simt.cu, conv_umma.cu), internal to the library.
#pragma once

#include "common.cuh"
#include "conv_plan.cuh"

namespace eae {

constexpr int kIm2colK = 96;  // 81 taps of the 9x9 kernel, zero-padded to a multiple of 32

// fp32 FFMA tap-list GEMM (always available).
int launch_gemm_simt(const GemmPlan& plan, cudaStream_t st);
// tcgen05 / TMEM / TMA tap-list GEMM; exact3x = hi/lo split of both operands (3 MMAs per product).
// w_lo: the low parts of the weights, same layout as plan.w (only read when exact3x).
// Weights of one layer for the tensor path: K-major [taps][128 out][Cin], tf32-rounded high part and
// (for exact3x) the tf32-rounded remainder.
struct UmmaWeights {
    const float* hi;
    const float* lo;
    int n_taps;
};
// gamma: K-major hi/lo of the GDN weights when plan.fuse != 0 (kernel version 2), else NULL.
// more / n_more: up to three further plans over the same input and weight array (the other output phases of a transposed
// convolution); kernel version 4 runs them with `plan` as ONE grid, otherwise they are launched one after the other.
int launch_gemm_umma(const GemmPlan& plan, const UmmaWeights& w, const UmmaWeights* gamma, bool exact3x,
                     cudaStream_t st, const GemmPlan* more = nullptr, int n_more = 0);
// The last layer in one launch (kernel version 6): conv2d_transpose k9 s4 of `in` [n, H/4, W/4, 128] through the
// K-major tap matrix `w` ([128 tap columns, 81 used][128 in]) + the col2im gather + the BT.601 cast / float output.
int launch_tconv9s4_fused(const float* in, const UmmaWeights& w, uint8_t* out_u8, float* out_f32, uint32_t n, int H, int W,
                          bool exact3x, cudaStream_t st);
// Can launch_gemm_umma fuse the quantizer (GemmPlan::quant_idx) into this k5 s2 convolution's store?
inline bool umma_can_fuse_quantizer(int Hg) { return Hg > 1; }
// Reads and clears the device-side timeout flag of the tensor path (synchronises `st`).
int umma_check_error(cudaStream_t st);
// The device word behind umma_check_error (NULL before the first tensor-path launch); kernels may read and clear it.
uint32_t* umma_error_flag_dev();
// 0 if the tcgen05 path can run on the current device (sm_100), else an error code with message.
int umma_available();

int launch_im2col_k9s4(const uint8_t* img, float* out, uint32_t n, int H, int W, cudaStream_t st);
int launch_col2im_k9s4(const float* P, uint8_t* out_u8, float* out_f32, uint32_t n, int H, int W,
                       cudaStream_t st);
int launch_quantize_to_planar(const float* y, const float* mean, const float* delta, int16_t* idx_planar,
                              float* q_off, uint32_t n, uint32_t hw, uint32_t* flag, cudaStream_t st);
int launch_dequantize_from_planar(const int16_t* idx_planar, const float* mean, const float* delta,
                                  float* q_off, uint32_t n, uint32_t hw, cudaStream_t st);

}  // namespace eae
