// "Tap-list implicit GEMM": the one contraction shape every layer of the EAE reduces to.
//
//   out[m, :] = epilogue( sum over taps t, channels c of  A_t[m, c] * W_t[c, :] )
//
// m runs over a regular grid (img, a, b) of Hg x Wg positions per image;
// A_t[m, c] = in[img, a * in_mul + dy_t, b * in_mul + dx_t, c]   (zero outside the input: TF SAME),
// the result goes to output pixel (a * out_mul + out_r, b * out_mul + out_s).
//
//   conv k5 s2 (components.py:126-136)      : 25 taps, in_mul 2, (dy, dx) = (ky - 1, kx - 1), out_mul 1
//   conv2d_transpose k5 s2 (:63-76)         : 4 launches (output phases r, s), in_mul 1, out_mul 2,
//                                             taps ky = r + 1 - 2 dy in [0, 5)  (exact adjoint, SAME)
//   GDN / IGDN (tfutils.py:363-397,480-509) : 1 tap, A = in^2, W = gamma[j, i], epilogue in (/ or *) sqrt(acc + beta)
//   (activations that feed a stride-2 convolution are stored parity-split so that each of its taps is
//    a dense box of one parity plane: see in_split / out_split)
//   conv k9 s4, Cin = 1 (:119-123)          : 1 tap over the im2col matrix [pixels, 96] (81 used)
//   conv2d_transpose k9 s4, Cout = 1 (:79-84): 1 tap, W = [128, 81 -> 128]; col2im kernel finishes it
#pragma once

#include <stdint.h>

namespace eae {

constexpr int kMaxTaps = 25;
constexpr int kCout = 128;  // every contraction here has 128 output columns (padded for the last layer)

enum EpilogueMode : int {
    kEpiBias = 0,   // out = acc + bias            (bias may be NULL)
    kEpiGdn = 1,    // out = in / sqrt(acc + beta) (A operand squared on load)
    kEpiIgdn = 2,   // out = in * sqrt(acc + beta)
};

struct Tap {
    int dy, dx;
    uint32_t w_off;  // element offset of this tap's [Cin][128] slice in the weight array
};

struct GemmPlan {
    const float* in;     // [n, Hin, Win, Cin]
    const float* w;      // per tap [Cin][128], cout contiguous
    const float* bias;   // [128]: bias (mode 0) or beta (modes 1, 2); may be NULL in mode 0
    float* out;          // [n, Hout, Wout, 128]
    int Hin, Win, Cin;   // Cin % 16 == 0
    int Hg, Wg;          // position grid per image
    int in_mul;
    int Hout, Wout;
    int out_mul, out_r, out_s;
    int mode;
    int in_split;        // input stored parity-split: [n][(y&1)*2+(x&1)][Hin/2][Win/2][Cin]
    int out_split;       // output stored parity-split: [n][(y&1)*2+(x&1)][Hout/2][Wout/2][128]
    int fuse;            // tensor path only: 0 none, 1 GDN, 2 IGDN applied to (acc + bias) before the store
    const float* fuse_beta;
    int fuse_single_pass;    //   the fused norm contracts in single-pass TF32 (squares rounded to nearest) instead of 3xTF32
    int fuse_precise;        //   the normalisation uses IEEE sqrt and division (tfutils.py:394-397) instead of the MUFU forms
    const uint8_t* img_u8;   // tensor path only: the layer is the k9 s4 convolution of this uint8 image [n, img_H, img_W]
    int img_H, img_W;        //   (A rows are gathered from the pixels inside the kernel; `in` is not read)
    // tensor path, multi-tap layers only: the quantizer of the latent fused into the store (umma_v3.cuh, OutGeom4):
    // planar int16 indices [n, 128, Hg * Wg] instead of the fp32 output; `out` is then not written
    int16_t* quant_idx;
    const float* quant_mean;     // [128] or NULL (zeros)
    const float* quant_delta;    // [128]
    uint32_t* quant_flag;        // bit 0: an index does not fit int16
    // tensor path, standalone IGDN (mode kEpiIgdn, flat position grid) only: the dequantizer fused into the operand
    // load: the input is delta[c] * k + mean[c] of the planar int16 indices [n images, 128, dequant_hw]; `in` is not read
    const int16_t* dequant_idx;
    const float* dequant_mean;   // [128] or NULL (zeros)
    const float* dequant_delta;  // [128]
    int dequant_hw;              // positions per image
    int n_taps;
    uint32_t M;          // n * Hg * Wg
    Tap taps[kMaxTaps];
};

}  // namespace eae
