/*
 * eae_b200.h — C ABI of libeae_b200.so, the B200-native (sm_100a) implementation of the entropy
 * autoencoder (EAE) codec hot path of thierrydumas/autoencoder_based_image_compression.
 *
 * Plain pointers and sizes only; no C++ / torch / numpy types. Every function returns an int:
 *   0        success
 *   1..4     the reference's error_code (kodak_tensorflow/lossless/c++/source/utils.h:12-19):
 *            1 capacity, 2 resource, 3 precision, 4 probability
 *   -1       a required pointer is NULL        (reference: std::invalid_argument, compression.cpp:9-12)
 *   -2       truncated unary length is 0       (reference: std::out_of_range from vector::at)
 *   -3       bad argument (shape / size / enum)
 *   -4       value not representable as int16  (reference: AssertionError, tools/tools.py:126-133)
 *   -5       lossless round trip altered data  (reference: AssertionError, compression.py:146-153)
 *   -10      CUDA runtime error, no usable sm_100 device, or out of memory; see eae_last_error()
 * There is NO CPU fallback anywhere behind this ABI: without a CUDA device every compute entry
 * point returns -10.
 *
 * "Reference" citations are relative to the kodak_tensorflow/ directory of the reference repository.
 *
 * Memory spaces: functions suffixed _host take host pointers and perform the H2D / D2H copies
 * themselves on the given stream and synchronise it before returning. Functions suffixed _dev take
 * device pointers, are asynchronous on the given stream and never synchronise. `stream` is a
 * cudaStream_t passed as void* (NULL = the legacy default stream).
 */
#ifndef EAE_B200_H
#define EAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAE_SUCCESS 0
#define EAE_ERR_CAPACITY 1
#define EAE_ERR_RESOURCE 2
#define EAE_ERR_PRECISION 3
#define EAE_ERR_PROBABILITY 4
#define EAE_ERR_NULL (-1)
#define EAE_ERR_UNARY_LENGTH (-2)
#define EAE_ERR_ARGUMENT (-3)
#define EAE_ERR_INT16_RANGE (-4)
#define EAE_ERR_ROUND_TRIP (-5)
#define EAE_ERR_CUDA (-10)

#define EAE_NB_MAPS 128   /* eae/graph/constants.py:42-44 */
#define EAE_STRIDE_PROD 16 /* eae/graph/constants.py:59 */

/* ------------------------------------------------------------------------------------------ */
/* Runtime                                                                                     */

/* Thread-local description of the last failure (CUDA error string, argument that was rejected). */
const char* eae_last_error(void);
/* ABI version, bumped on any signature change. */
int eae_abi_version(void);
/* Number of visible CUDA devices (0 when none: every compute call then returns EAE_ERR_CUDA). */
int eae_device_count(void);
/* SM count / compute capability of `device`. */
int eae_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* Cumulative number of kernels this library has launched in this process (bench: gpu_launches). */
uint64_t eae_launch_count(void);

/* Selects the device used by the entry points that take no codec (coder / glue); codecs carry their own. */
int eae_set_device(int device);
/* How host threads wait for the current device in the _host entry points: 1 = sleep on an interrupt
 * (cudaDeviceScheduleBlockingSync), 0 = the driver's default (spins while cores outnumber contexts). Several host
 * threads per GPU and one process per GPU (the pipelined codec on an 8-GPU box) oversubscribe the cores when every
 * waiting thread spins; a blocked thread costs a wake-up of some tens of microseconds per call instead. */
int eae_set_blocking_sync(int on);

/* Optional per-kernel-class device timing (CUDA events on the launching stream around every launch of
 * a class). Classes: 0 gemm_conv (k5 s2 convolutions), 1 gemm_tconv (k5 s2 transposed convolutions),
 * 2 gemm_gdn, 3 gemm_thin (first / last layer contractions), 4 im2col, 5 col2im, 6 quantize,
 * 7 dequantize, 8 coder_encode, 9 coder_decode, 10 pack. eae_profile_read synchronises the device. */
#define EAE_PROFILE_CLASSES 11
int eae_profile_enable(int on);
int eae_profile_reset(void);
int eae_profile_read(int kernel_class, uint64_t* launches, double* total_ms);
const char* eae_profile_name(int kernel_class);

/* Pinned host memory and device memory helpers for callers without a CUDA binding. */
void* eae_host_alloc(size_t bytes);
void eae_host_free(void* p);
void* eae_device_alloc(size_t bytes);
void eae_device_free(void* p);
int eae_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int eae_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int eae_stream_create(void** stream);
int eae_stream_destroy(void* stream);
int eae_stream_synchronize(void* stream);
/* CUDA events for device-side timing on the launching stream. */
int eae_event_create(void** event);
int eae_event_destroy(void* event);
int eae_event_record(void* event, void* stream);
int eae_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on `stop` */

/* ------------------------------------------------------------------------------------------ */
/* Lossless coder                                                                              */

/*
 * Replaces: uint32_t compress_lossless(size, array_input, array_output, truncated_unary_length,
 *           probabilities)   lossless/c++/source/compression.h:41-45, compression.cpp:3-65
 * (the only FFI entry of the reference, bound by lossless/interface_cython.pyx:6-11, 54-58).
 * Encodes `size` int16 symbols of one map on the GPU (signed UEG0: truncated-unary prefix through
 * the 16-bit binary arithmetic coder with probabilities[i] = P(bin i == 0), Exp-Golomb-0 suffix and
 * sign as bypass bits), counts bac + bypass bits, decodes on the GPU from those two buffers into
 * array_output. Host pointers. C++ exceptions of the reference become the return codes above.
 */
int eae_compress_lossless(uint32_t size, const int16_t* array_input, int16_t* array_output,
                          uint8_t truncated_unary_length, const double* probabilities,
                          uint32_t* nb_bits);

/* Same, additionally returning the two byte buffers (bit i of a buffer is bit (i & 7) of byte
 * i >> 3, lossless/c++/source/Bitstream.cpp:36-58). Each buffer must hold
 * eae_coder_capacity_bytes(size, L) bytes. array_output may be NULL (no decode). */
int eae_encode_map_host(uint32_t size, const int16_t* array_input, uint8_t truncated_unary_length,
                        const double* probabilities, uint8_t* bac_bytes, uint32_t* bac_bits,
                        uint8_t* bypass_bytes, uint32_t* bypass_bits);
/* Standalone decoder the reference lacks (compression.cpp:51-63 as its own entry point). */
int eae_decode_map_host(uint32_t size, int16_t* array_output, uint8_t truncated_unary_length,
                        const double* probabilities, const uint8_t* bac_bytes, uint32_t bac_bits,
                        const uint8_t* bypass_bytes, uint32_t bypass_bits);
/* ceil(size * max(32, L) / 8): bytes of each per-map buffer (compression.cpp:24). */
uint32_t eae_coder_capacity_bytes(uint32_t size, uint32_t truncated_unary_length);

/*
 * Replaces the per-map loop of lossless.compression.compress_lossless_maps
 * (lossless/compression.py:67-81): all `nb_maps` maps of one [h, w, nb_maps] int16 latent (HWC as in
 * the reference; map i is ref_hwc[:, :, i] flattened row-major) are coded as independent streams
 * with row i of `table` (float64 [nb_maps, L]). Maps flagged in skip_mask (uint8 [nb_maps], may be
 * NULL) are copied through uncoded and get nb_bits 0 (the reference's "exception" map,
 * compression.py:68-75, whose cost the Python host computes from eae_histogram_maps_host).
 * rec_hwc receives the DECODED maps. nb_bits_each_map: uint32 [nb_maps].
 * Returns the first non-zero per-stream error code, if any.
 */
int eae_compress_lossless_maps_host(const int16_t* ref_hwc, uint32_t h, uint32_t w, uint32_t nb_maps,
                                    const double* table, uint32_t truncated_unary_length,
                                    const uint8_t* skip_mask, int16_t* rec_hwc,
                                    uint32_t* nb_bits_each_map, void* stream);

/*
 * Replaces lossless.compression.rescale_compress_lossless_maps (lossless/compression.py:84-154)
 * minus the table load: k = int16(round(q / delta)) per map (error -4 if |round| >= 32768),
 * code/decode as above, verify q == float32(k) * delta bit-for-bit (error -5). `entropy_bits_skip`
 * (may be NULL) receives, for maps in skip_mask, nothing — the host adds the entropy estimate.
 * idx_hwc_out (may be NULL) receives the int16 indices.
 */
int eae_rescale_compress_lossless_maps_host(const float* centered_quantized_hwc, uint32_t h, uint32_t w,
                                            uint32_t nb_maps, const float* bin_widths,
                                            const double* table, uint32_t truncated_unary_length,
                                            const uint8_t* skip_mask, int16_t* idx_hwc_out,
                                            uint32_t* nb_bits_each_map, void* stream);

/*
 * Per-map symbol histograms (tools/tools.py count_symbols :322-388 on integer indices): for each of
 * the nb_maps maps of an [n_images, h, w, nb_maps] int16 array, reduced over (images?, h, w).
 * If per_image != 0 there are n_images * nb_maps histograms (rate_3d, :931-989), else nb_maps
 * histograms over all images (stats.py). For histogram j: min_out[j], max_out[j] (int32) and counts
 * hist_out[j * hist_cap + (symbol - min)] (uint64). Returns -3 with *needed_cap set if a range
 * exceeds hist_cap. abs_sum_out (uint64, may be NULL): sum |k| (count_nb_deads, :294-320).
 */
int eae_histogram_maps_host(const int16_t* idx_nhwc, uint32_t n_images, uint32_t h, uint32_t w,
                            uint32_t nb_maps, int per_image, int32_t* min_out, int32_t* max_out,
                            uint64_t* hist_out, uint32_t hist_cap, uint32_t* needed_cap,
                            uint64_t* abs_sum_out, void* stream);

/* The same on device-resident PLANAR streams (stream image * nb_maps + map holds its `size` symbols contiguously: the
 * layout the codec and the coder primitives below use), asynchronous, warp-level (one warp per stream, 16-byte loads,
 * counts in shared memory). Two calls: with hist_dev == NULL it writes min_dev / max_dev (and abs_sum_dev if given);
 * with hist_dev != NULL it reads min_dev and writes the counts (hist_cap bins per histogram, zeroed first; symbols
 * beyond min + hist_cap are not counted). Number of histograms: per_image ? n_images * nb_maps : nb_maps. */
int eae_histogram_streams_dev(const int16_t* idx_planar_dev, uint32_t n_images, uint32_t size, uint32_t nb_maps,
                              int per_image, int32_t* min_dev, int32_t* max_dev, uint64_t* abs_sum_dev,
                              uint64_t* hist_dev, uint32_t hist_cap, void* stream);

/* ---- device-level coder primitives (planar streams: stream s holds `size` int16 at s * size) ---- */

/* Bytes of one per-stream slot in the scratch arenas (capacity rounded up to 16). */
uint32_t eae_coder_slot_bytes(uint32_t size, uint32_t truncated_unary_length);

/* [N, h, w, C] int16 (or one [h, w, C]) -> planar [N * C, h * w] and back. */
int eae_nhwc_to_planar_i16_dev(const int16_t* nhwc, int16_t* planar, uint32_t n_images, uint32_t hw,
                               uint32_t nb_maps, void* stream);
int eae_planar_to_nhwc_i16_dev(const int16_t* planar, int16_t* nhwc, uint32_t n_images, uint32_t hw,
                               uint32_t nb_maps, void* stream);

/*
 * One GPU lane per stream. Stream s uses table row (s % table_rows); streams whose row is flagged
 * in skip_mask_dev (may be NULL) are not coded (bits 0). Outputs per stream: bac_bits, bypass_bits,
 * err (uint32 each), bytes in bac_slots / bypass_slots at s * slot_bytes.
 */
int eae_encode_streams_dev(const int16_t* idx_planar, uint32_t n_streams, uint32_t size,
                           const double* table_dev, uint32_t table_rows, uint32_t truncated_unary_length,
                           const uint8_t* skip_mask_dev, uint8_t* bac_slots, uint8_t* bypass_slots,
                           uint32_t slot_bytes, uint32_t* bac_bits, uint32_t* bypass_bits,
                           uint32_t* err, void* stream);
/* Decoder: stream s reads bac bytes at bac_base + bac_off[s], bypass at byp_base + byp_off[s]
 * (uint64 byte offsets, need not be aligned). */
int eae_decode_streams_dev(int16_t* idx_planar_out, uint32_t n_streams, uint32_t size,
                           const double* table_dev, uint32_t table_rows, uint32_t truncated_unary_length,
                           const uint8_t* skip_mask_dev, const uint8_t* bac_base, const uint64_t* bac_off,
                           const uint32_t* bac_bits, const uint8_t* byp_base, const uint64_t* byp_off,
                           const uint32_t* bypass_bits, uint32_t* err, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* numpy glue of the hot path (tools/tools.py)                                                 */

/* tls.quantize_per_map (:883-929): out = delta[c] * rint(data / delta[c]), fp32, half-to-even.
 * data/out: [n_rows, nb_maps] (any leading shape flattened). -3 if a bin width is <= 0. */
int eae_quantize_per_map_host(const float* data, float* out, uint64_t n_rows, uint32_t nb_maps,
                              const float* bin_widths, void* stream);
/* tls.cast_float_to_int16 (:95-133): rint then int16; -4 if |rint(x)| >= 32768. */
int eae_cast_float_to_int16_host(const float* data, int16_t* out, uint64_t n, void* stream);
/* tls.cast_bt601 (:61-93): rint(clip(x, 16, 235)) -> uint8. */
int eae_cast_bt601_host(const float* data, uint8_t* out, uint64_t n, void* stream);
/* tls.psnr_2d (:831-881) numerator: sum over pixels of (a - b)^2 as an exact uint64. */
int eae_sum_squared_error_u8_host(const uint8_t* a, const uint8_t* b, uint64_t n, uint64_t* sse,
                                  void* stream);

/* tls.count_nb_deads (:294-320): nb_deads[i] = number of maps j with sum |data[i, :, :, j]| == 0.
 * data: float32 [n, hw, nb_maps]; nb_deads: uint32 [n]. */
int eae_count_nb_deads_host(const float* data, uint32_t n, uint64_t hw, uint32_t nb_maps,
                            uint32_t* nb_deads, void* stream);

/*
 * The counting part of lossless.stats.save_statistics (lossless/stats.py:13-68, 70-134, 136-195, 197-241,
 * 243-320) over a calibration set of latents y: float32 [n_rows, nb_maps] (an NHWC batch, flattened).
 *   sum_out  float64 [nb_maps]: per-map sum (map_mean = sum / n_rows, :311)
 *   min_out, max_out float32 [nb_maps]
 *   unit_hist_out (may be NULL) uint64 [nb_maps, unit_cap]: numpy.histogram of map c with unit bins from
 *     floor(min) to ceil(max) (compute_probabilities_intervals with size_interval 1, used by
 *     find_index_map_exception); needed_cap (may be NULL) receives the widest range; EAE_ERR_ARGUMENT if it
 *     exceeds unit_cap
 *   abs_counts_out (may be NULL) uint64 [nb_maps, L + 1]: occurrences of a = min(|rint((y - mean) / delta)|, L)
 *     in float32 arithmetic, from which count_binary_decisions follows: zeros[j] = counts[j],
 *     ones[j] = sum of counts[j + 1 ..]
 */
int eae_latent_statistics_host(const float* y, uint64_t n_rows, uint32_t nb_maps, double* sum_out,
                               float* min_out, float* max_out, uint64_t* unit_hist_out, uint32_t unit_cap,
                               uint32_t* needed_cap, const float* mean, const float* delta,
                               uint32_t truncated_unary_length, uint64_t* abs_counts_out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Transforms                                                                                  */

/* Weights in TensorFlow variable layouts (eae/graph/EntropyAutoencoder.py:108-224). Host pointers,
 * copied at creation. gamma_3/beta_3/gamma_4/beta_4 may be NULL iff are_bin_widths_learned != 0. */
typedef struct eae_weights {
    const float* weights_1; /* [9,9,1,128]   kh,kw,in,out */
    const float* biases_1;  /* [128] */
    const float* gamma_1;   /* [128,128]     [in j, out i] */
    const float* beta_1;    /* [128] */
    const float* weights_2; /* [5,5,128,128] */
    const float* biases_2;
    const float* gamma_2;
    const float* beta_2;
    const float* weights_3; /* [5,5,128,128] */
    const float* biases_3;
    const float* gamma_3;
    const float* beta_3;
    const float* gamma_4;
    const float* beta_4;
    const float* weights_4; /* [5,5,128,128] kh,kw,out,in (conv2d_transpose filter) */
    const float* biases_4;
    const float* gamma_5;
    const float* beta_5;
    const float* weights_5; /* [5,5,128,128] kh,kw,out,in */
    const float* biases_5;
    const float* gamma_6;
    const float* beta_6;
    const float* weights_6; /* [9,9,1,128]   kh,kw,out,in ; no bias (components.py:79-84) */
} eae_weights_t;

/* Arithmetic of the 128->128 contractions. */
#define EAE_MATH_FP32_SIMT 0 /* fp32 FFMA on CUDA cores (exact-fp32 parity mode, always available) */
#define EAE_MATH_TF32X3 1    /* tcgen05 kind::tf32, hi/lo split of both operands, 3 MMAs, fp32 accumulate */
#define EAE_MATH_TF32 2      /* tcgen05 kind::tf32, single pass (throughput mode; index mismatches reported) */
#define EAE_MATH_MIXED 3     /* analysis transform (it decides the quantization indices) as EAE_MATH_TF32X3; synthesis
                              * transform single pass TF32 with 3xTF32 IGDN norms (its bar is PSNR within 0.01 dB) */

typedef struct eae_codec eae_codec_t;

/* Replaces EntropyAutoencoder.__init__/.initialization (EntropyAutoencoder.py:36-251, 440-463) and
 * IsolatedDecoder.__init__/.initialization (IsolatedDecoder.py:21-129) for inference: one handle
 * holds both transforms on `device`, its own stream-ordered workspace and the weights re-laid-out
 * for the kernels. */
int eae_codec_create(eae_codec_t** codec, const eae_weights_t* weights, int are_bin_widths_learned,
                     int device);
int eae_codec_destroy(eae_codec_t* codec);
int eae_codec_set_math(eae_codec_t* codec, int math_mode);
int eae_codec_get_math(const eae_codec_t* codec);
/* GPU threads per coded stream in the fused pipeline: 0 (default) = one warp per stream while they fit, the
 * lowest batch latency; 1, 2 or 4 pack 32, 16 or 8 streams per warp, which costs latency but far fewer issue
 * slots — the choice when batches are pipelined on several CUDA streams. */
int eae_codec_set_coder_lanes(eae_codec_t* codec, uint32_t lanes);

/* Replaces eae.batching.encode_mini_batches' sess.run(node_y) (eae/batching.py:56-100,
 * eae/graph/components.py:86-142): uint8 [n, h, w, 1] -> float32 [n, h/16, w/16, 128]. */
int eae_encode_host(eae_codec_t* codec, const uint8_t* luminances, uint32_t n, uint32_t h, uint32_t w,
                    float* y_out, void* stream);
int eae_encode_dev(eae_codec_t* codec, const uint8_t* luminances_dev, uint32_t n, uint32_t h, uint32_t w,
                   float* y_out_dev, void* stream);
/* Replaces eae.batching.decode_mini_batches' sess.run(node_reconstruction) + tls.cast_bt601
 * (eae/batching.py:11-54, components.py:11-84): float32 [n, h/16, w/16, 128] -> uint8 [n, h, w, 1]. */
int eae_decode_host(eae_codec_t* codec, const float* quantized_y, uint32_t n, uint32_t h, uint32_t w,
                    uint8_t* reconstruction_out, void* stream);
int eae_decode_dev(eae_codec_t* codec, const float* quantized_y_dev, uint32_t n, uint32_t h, uint32_t w,
                   uint8_t* reconstruction_out_dev, void* stream);
/* The decoder graph's float32 output node_reconstruction itself (components.py:79-84), before
 * tls.cast_bt601: float32 [n, h/16, w/16, 128] -> float32 [n, h, w, 1]. */
int eae_decode_float_host(eae_codec_t* codec, const float* quantized_y, uint32_t n, uint32_t h, uint32_t w,
                          float* reconstruction_out, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused codec: encode -> quantize -> lossless code -> container, and back                     */

/* Coding parameters of one operating point (reconstructing_eae_kodak.py:170-217). Host pointers. */
typedef struct eae_coding_params {
    const float* map_mean;    /* [128] or NULL (zeros) */
    const float* bin_widths;  /* [128] test bin widths (multiplier * learned widths) */
    const double* table;      /* [128, L] P(bin j == 0) per map */
    uint32_t truncated_unary_length;
} eae_coding_params_t;

/*
 * Container produced by eae_compress_* for a batch of n images (all little-endian):
 *   u32 magic 'EAEB', u32 version, u32 n, u32 h, u32 w, u32 nb_maps, u32 L, u32 reserved
 *   n * nb_maps * { u32 bac_bits, u32 bypass_bits }                      stream table
 *   for each stream in order: ceil(bac_bits/8) bytes, then ceil(bypass_bits/8) bytes   payload
 * Stream s = image * nb_maps + map. Every map is really coded (no "exception" shortcut).
 */
uint64_t eae_container_bound(uint32_t n, uint32_t h, uint32_t w, uint32_t truncated_unary_length);

/* Per-batch statistics (device-reduced): total bits per map over the batch (what numpy.mean(rate, axis=1) reduces,
 * reconstructing_eae_kodak.py:810-815), total bits, and the number of dead maps (tools.py:294-320). */
typedef struct eae_batch_stats {
    uint64_t bits_per_map[EAE_NB_MAPS];
    uint64_t total_bits;
    uint64_t nb_dead_maps; /* summed over images */
} eae_batch_stats_t;

/* images uint8 [n,h,w] (host) -> container bytes (host). *container_bytes receives the size. */
int eae_compress_host(eae_codec_t* codec, const eae_coding_params_t* params, const uint8_t* luminances,
                      uint32_t n, uint32_t h, uint32_t w, uint8_t* container, uint64_t container_cap,
                      uint64_t* container_bytes, eae_batch_stats_t* stats, void* stream);
/* container bytes (host) -> reconstruction uint8 [n,h,w] (host). */
int eae_decompress_host(eae_codec_t* codec, const eae_coding_params_t* params, const uint8_t* container,
                        uint64_t container_bytes, uint8_t* reconstruction, uint64_t reconstruction_cap,
                        void* stream);
/* Same pipeline with the images / container / reconstruction resident in device memory. Asynchronous: the calls
 * return once the work is enqueued on `stream`, so they can only report host-side errors. What happens on the device
 * is reported by eae_codec_poll_status:
 *   - container_bytes_dev (uint64 on the device) receives the size the container NEEDS. If it exceeds container_cap
 *     the streams that do not fit are not written (the header and the stream table always are) and the status says
 *     container_overflow: size the buffer with eae_container_bound, or poll and retry.
 *   - an index that does not fit int16 (tools.py:126-133), a coder error (codes 1..4 of compression.cpp:32-62) and a
 *     time-out of the tensor pipeline are recorded in the codec and reported by the poll.
 * eae_decompress_dev reads at most container_bytes bytes at container_dev (the exact size of the container, or the
 * capacity of the buffer when the size is only known on the device). The stream table is validated on the device: a
 * stream whose bit counts exceed the coder capacity (compression.cpp:24) or whose payload would run past
 * container_bytes is decoded as an empty stream (resource error 2 for that stream) and the status says
 * container_invalid; nothing outside [container_dev, container_dev + container_bytes) is read. */
int eae_compress_dev(eae_codec_t* codec, const eae_coding_params_t* params, const uint8_t* luminances_dev,
                     uint32_t n, uint32_t h, uint32_t w, uint8_t* container_dev, uint64_t container_cap,
                     uint64_t* container_bytes_dev, eae_batch_stats_t* stats_dev, void* stream);
int eae_decompress_dev(eae_codec_t* codec, const eae_coding_params_t* params, const uint8_t* container_dev,
                       uint64_t container_bytes, uint32_t n, uint32_t h, uint32_t w, uint8_t* reconstruction_dev,
                       void* stream);

/* Running totals over steps: when set (device pointer, caller-owned, zeroed by the caller; NULL to stop), every
 * eae_compress_* step of this codec also ADDS its statistics to *acc_dev with atomic adds, so several codecs (pipeline
 * slots) may share one accumulator and a multi-GPU job reduces it ONCE at the end, as the reference does
 * (reconstructing_eae_kodak.py:810-815), instead of once per batch. */
int eae_codec_set_stats_accumulator(eae_codec_t* codec, eae_batch_stats_t* acc_dev);

/* What the device-resident steps of this codec recorded since the previous poll (sticky), plus the decoder's error of
 * the LAST eae_decompress_dev step. Enqueues a one-CTA kernel on `stream`, waits for the stream, clears the record.
 * Returns 0, or the code the corresponding _host entry point would have returned (EAE_ERR_INT16_RANGE, 1..4,
 * EAE_ERR_ARGUMENT for a container that did not fit, EAE_ERR_CAPACITY / EAE_ERR_RESOURCE for an invalid container,
 * EAE_ERR_CUDA for a tensor-pipeline time-out) with eae_last_error set. `status` may be NULL. */
typedef struct eae_codec_status {
    uint32_t int16_overflow;      /* 1: an index did not fit int16 */
    uint32_t container_overflow;  /* 1: a container needed more than container_cap bytes */
    uint32_t container_invalid;   /* 1: a stream table failed the device-side validation of eae_decompress_dev */
    uint32_t coder_error;         /* first coder error code (1..4), 0 if none */
    uint32_t tensor_timeout_mask; /* role mask of a timed-out tcgen05 pipeline, 0 if none */
} eae_codec_status_t;
int eae_codec_poll_status(eae_codec_t* codec, void* stream, eae_codec_status_t* status);

/* Debug / parity hooks: the int16 indices [n, nb_maps, h/16 * w/16] (planar) produced by the last
 * eae_compress_* / eae_decompress_* on this codec, copied to host on the stream of that step (synchronises it). */
int eae_last_indices_host(eae_codec_t* codec, int16_t* idx_planar_out, uint64_t n_elems);

/* Debug hook: the correctly rounded square root and quotient the fused GDN / IGDN normalisation uses
 * (tfutils.py:394-397, 506-509) against the compiler's sqrt.rn / div.rn: every float n in [2^-20, 2^40] for the
 * square root - each also through the packed (fp32x2) x / sqrt(n) and x * sqrt(n) of the fused tails with two
 * pseudo-random x -, n_pairs pseudo-random pairs for the scalar quotient. Both counts must be 0. */
int eae_debug_check_norm_arithmetic(uint64_t n_pairs, uint64_t* sqrt_mismatches, uint64_t* div_mismatches);

#ifdef __cplusplus
}
#endif
#endif /* EAE_B200_H */
