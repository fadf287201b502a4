// Launchers shared between translation units of libeae_b200.so (not part of the C ABI).
#pragma once

#include "common.cuh"

namespace eae {

// ---- coder.cu ----
uint32_t coder_capacity_bits(uint32_t size, uint32_t L);
int launch_encode_streams(const int16_t* idx_planar, uint32_t n_streams, uint32_t size,
                          const double* table_dev, uint32_t table_rows, uint32_t L,
                          const uint8_t* skip_mask_dev, uint8_t* bac_slots, uint8_t* byp_slots,
                          uint32_t slot_bytes, uint32_t* bac_bits, uint32_t* byp_bits, uint32_t* err,
                          cudaStream_t st, uint32_t lanes = 0, void* scratch = nullptr,
                          const uint64_t* qtable_dev = nullptr, const uint8_t* row_flags_dev = nullptr);
// Bytes of `scratch` (truncated-unary bit strings + bin counts); NULL scratch = stream-ordered allocation.
size_t coder_encode_scratch_bytes(uint32_t n_streams, uint32_t size, uint32_t L);
int launch_decode_streams(int16_t* idx_planar_out, uint32_t n_streams, uint32_t size,
                          const double* table_dev, uint32_t table_rows, uint32_t L,
                          const uint8_t* skip_mask_dev, const uint8_t* bac_base, const uint64_t* bac_off,
                          const uint32_t* bac_bits, const uint8_t* byp_base, const uint64_t* byp_off,
                          const uint32_t* byp_bits, uint32_t* err, cudaStream_t st, uint32_t lanes = 0,
                          const uint64_t* qtable_dev = nullptr, const uint8_t* row_flags_dev = nullptr);
// Fixed-point multipliers and validity flags of a probability table (coder v3); see prepare_table_kernel.
int launch_prepare_table(const double* table_dev, uint32_t rows, uint32_t L, uint64_t* qtable_dev,
                         uint8_t* row_flags_dev, cudaStream_t st);
int launch_slot_offsets(uint64_t* off, uint32_t n, uint32_t slot_bytes, cudaStream_t st);
int launch_transpose_i16(const int16_t* in, int16_t* out, uint32_t batch, uint32_t rows, uint32_t cols,
                         cudaStream_t st);
// (planar streams: stream image * C + map holds its hw symbols contiguously)
int launch_histograms(const int16_t* idx_planar_dev, uint32_t n_images, uint32_t hw, uint32_t C,
                      int per_image, int32_t* mn, int32_t* mx, unsigned long long* abs_sum,
                      unsigned long long* hist, uint32_t cap, bool only_minmax, cudaStream_t st);

// ---- glue.cu ----
// out[r, c] = delta[c] * rint(data[r, c] / delta[c])
int launch_quantize_per_map(const float* data, float* out, uint64_t n_rows, uint32_t C,
                            const float* delta_dev, cudaStream_t st);
// idx = int16(rint(x)); *flag |= 1 when |rint(x)| >= 32768
int launch_cast_float_to_int16(const float* data, int16_t* out, uint64_t n, uint32_t* flag_dev,
                               cudaStream_t st);
// idx[r, c] = int16(rint(q[r, c] / delta[c])), flag bit0 on int16 overflow
int launch_rescale_to_int16(const float* q, int16_t* out, uint64_t n_rows, uint32_t C,
                            const float* delta_dev, uint32_t* flag_dev, cudaStream_t st);
// flag bit1 when q[r, c] != float(idx[r, c]) * delta[c] (bitwise, NaN-aware like assert_equal)
int launch_check_rescaled(const float* q, const int16_t* idx, uint64_t n_rows, uint32_t C,
                          const float* delta_dev, uint32_t* flag_dev, cudaStream_t st);
int launch_cast_bt601(const float* data, uint8_t* out, uint64_t n, cudaStream_t st);
int launch_sse_u8(const uint8_t* a, const uint8_t* b, uint64_t n, unsigned long long* sse_dev,
                  cudaStream_t st);

}  // namespace eae
