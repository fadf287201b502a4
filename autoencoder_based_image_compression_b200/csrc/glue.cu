// Element-wise glue of the hot path (kodak_tensorflow/tools/tools.py) as HBM-bound kernels:
// quantize_per_map (:883-929), cast_float_to_int16 (:95-133), cast_bt601 (:61-93), the rescale +
// round-trip check of rescale_compress_lossless_maps (lossless/compression.py:134-153) and the
// squared-error sum of psnr_2d (:831-881). All fp32 steps use the explicitly rounded intrinsics so
// that no FMA contraction or reciprocal substitution changes a result: x / delta is a true IEEE
// division and rint is round-half-to-even, exactly as numpy computes them.
#include <memory>
#include <string.h>

#include "common.cuh"
#include "internal.cuh"

namespace eae {

namespace {

constexpr int kThreads = 256;

inline uint32_t grid_for(uint64_t n, uint32_t per_thread = 1)
{
    uint64_t blocks = (n + (uint64_t)kThreads * per_thread - 1) / ((uint64_t)kThreads * per_thread);
    const uint64_t cap = 148ull * 32ull;
    return (uint32_t)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

__global__ void quantize_per_map_kernel(const float* __restrict__ data, float* __restrict__ out,
                                        uint64_t n, uint32_t C, const float* __restrict__ delta)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const float d = __ldg(delta + (uint32_t)(t % C));
        out[t] = __fmul_rn(d, rintf(__fdiv_rn(data[t], d)));
    }
}

__global__ void cast_float_to_int16_kernel(const float* __restrict__ data, int16_t* __restrict__ out,
                                           uint64_t n, uint32_t* __restrict__ flag)
{
    bool bad = false;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const float r = rintf(data[t]);
        // tools.py:126-128 assert_array_less(|r|, 32768): NaN fails the assertion as well.
        if (!(fabsf(r) < 32768.f)) { bad = true; out[t] = 0; }
        else out[t] = (int16_t)(int)r;
    }
    if (bad) atomicOr(flag, 1u);
}

__global__ void rescale_to_int16_kernel(const float* __restrict__ q, int16_t* __restrict__ out,
                                        uint64_t n, uint32_t C, const float* __restrict__ delta,
                                        uint32_t* __restrict__ flag)
{
    bool bad = false;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const float d = __ldg(delta + (uint32_t)(t % C));
        const float r = rintf(__fdiv_rn(q[t], d));
        if (!(fabsf(r) < 32768.f)) { bad = true; out[t] = 0; }
        else out[t] = (int16_t)(int)r;
    }
    if (bad) atomicOr(flag, 1u);
}

__global__ void check_rescaled_kernel(const float* __restrict__ q, const int16_t* __restrict__ idx,
                                      uint64_t n, uint32_t C, const float* __restrict__ delta,
                                      uint32_t* __restrict__ flag)
{
    bool bad = false;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const float d = __ldg(delta + (uint32_t)(t % C));
        const float rec = __fmul_rn((float)idx[t], d);
        const float x = q[t];
        // numpy.testing.assert_equal: equal values, or NaN in the same place; +0 == -0 passes.
        if (!(rec == x || (rec != rec && x != x))) bad = true;
    }
    if (bad) atomicOr(flag, 2u);
}

__global__ void cast_bt601_kernel(const float* __restrict__ data, uint8_t* __restrict__ out, uint64_t n)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        out[t] = (uint8_t)(int)rintf(fminf(fmaxf(data[t], 16.f), 235.f));
    }
}

// alive[img * C + c] = 1 as soon as one coefficient of the map is non-zero (NaN counts as non-zero,
// as numpy's sum(|x|) == 0 would be False).
__global__ void alive_maps_kernel(const float* __restrict__ data, uint64_t n, uint64_t hw, uint32_t C,
                                  uint32_t* __restrict__ alive)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        if (!(data[t] == 0.f)) alive[(t / (hw * C)) * C + (t % C)] = 1u;
    }
}

__global__ void sse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint64_t n,
                              unsigned long long* __restrict__ sse)
{
    unsigned long long acc = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const int d = (int)a[t] - (int)b[t];
        acc += (unsigned long long)(d * d);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sse, acc);
}

}  // namespace

int launch_quantize_per_map(const float* data, float* out, uint64_t n_rows, uint32_t C,
                            const float* delta_dev, cudaStream_t st)
{
    const uint64_t n = n_rows * C;
    if (!n) return 0;
    quantize_per_map_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(data, out, n, C, delta_dev);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_cast_float_to_int16(const float* data, int16_t* out, uint64_t n, uint32_t* flag_dev,
                               cudaStream_t st)
{
    if (!n) return 0;
    cast_float_to_int16_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(data, out, n, flag_dev);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_rescale_to_int16(const float* q, int16_t* out, uint64_t n_rows, uint32_t C,
                            const float* delta_dev, uint32_t* flag_dev, cudaStream_t st)
{
    const uint64_t n = n_rows * C;
    if (!n) return 0;
    rescale_to_int16_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(q, out, n, C, delta_dev, flag_dev);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_check_rescaled(const float* q, const int16_t* idx, uint64_t n_rows, uint32_t C,
                          const float* delta_dev, uint32_t* flag_dev, cudaStream_t st)
{
    const uint64_t n = n_rows * C;
    if (!n) return 0;
    check_rescaled_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(q, idx, n, C, delta_dev, flag_dev);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_cast_bt601(const float* data, uint8_t* out, uint64_t n, cudaStream_t st)
{
    if (!n) return 0;
    cast_bt601_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(data, out, n);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_sse_u8(const uint8_t* a, const uint8_t* b, uint64_t n, unsigned long long* sse_dev,
                  cudaStream_t st)
{
    if (!n) return 0;
    sse_u8_kernel<<<grid_for(n, 8), kThreads, 0, st>>>(a, b, n, sse_dev);
    EAE_LAUNCH_OK();
    return 0;
}

}  // namespace eae

using namespace eae;

extern "C" int eae_quantize_per_map_host(const float* data, float* out, uint64_t n_rows, uint32_t nb_maps,
                                         const float* bin_widths, void* stream)
{
    if (!data || !out || !bin_widths) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (nb_maps == 0) { set_error("nb_maps is 0"); return EAE_ERR_ARGUMENT; }
    for (uint32_t c = 0; c < nb_maps; c++) {
        if (!(bin_widths[c] > 0.f)) {   // tools.py:924-925
            set_error("A quantization bin width is not strictly positive.");
            return EAE_ERR_ARGUMENT;
        }
    }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = n_rows * nb_maps;
    DevBuf d, o, bw;
    EAE_TRY(d.alloc(n * 4));
    EAE_TRY(o.alloc(n * 4));
    EAE_TRY(bw.alloc((size_t)nb_maps * 4));
    EAE_CUDA_OK(cudaMemcpyAsync(d.p, data, n * 4, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(bw.p, bin_widths, (size_t)nb_maps * 4, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_quantize_per_map(d.as<float>(), o.as<float>(), n_rows, nb_maps, bw.as<float>(), st));
    EAE_CUDA_OK(cudaMemcpyAsync(out, o.p, n * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int eae_cast_float_to_int16_host(const float* data, int16_t* out, uint64_t n, void* stream)
{
    if (!data || !out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf d, o, f;
    EAE_TRY(d.alloc(n * 4));
    EAE_TRY(o.alloc(n * 2));
    EAE_TRY(f.alloc(4));
    EAE_CUDA_OK(cudaMemsetAsync(f.p, 0, 4, st));
    EAE_CUDA_OK(cudaMemcpyAsync(d.p, data, n * 4, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_cast_float_to_int16(d.as<float>(), o.as<int16_t>(), n, f.as<uint32_t>(), st));
    uint32_t flag = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(&flag, f.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(out, o.p, n * 2, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (flag & 1u) {
        set_error("The rounded array elements cannot be represented as 16-bit signed integers.");
        return EAE_ERR_INT16_RANGE;
    }
    return 0;
}

extern "C" int eae_cast_bt601_host(const float* data, uint8_t* out, uint64_t n, void* stream)
{
    if (!data || !out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf d, o;
    EAE_TRY(d.alloc(n * 4));
    EAE_TRY(o.alloc(n));
    EAE_CUDA_OK(cudaMemcpyAsync(d.p, data, n * 4, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_cast_bt601(d.as<float>(), o.as<uint8_t>(), n, st));
    EAE_CUDA_OK(cudaMemcpyAsync(out, o.p, n, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int eae_sum_squared_error_u8_host(const uint8_t* a, const uint8_t* b, uint64_t n, uint64_t* sse,
                                             void* stream)
{
    if (!a || !b || !sse) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf da, db, acc;
    EAE_TRY(da.alloc(n));
    EAE_TRY(db.alloc(n));
    EAE_TRY(acc.alloc(8));
    EAE_CUDA_OK(cudaMemsetAsync(acc.p, 0, 8, st));
    EAE_CUDA_OK(cudaMemcpyAsync(da.p, a, n, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(db.p, b, n, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_sse_u8(da.as<uint8_t>(), db.as<uint8_t>(), n, acc.as<unsigned long long>(), st));
    unsigned long long v = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(&v, acc.p, 8, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    *sse = v;
    return 0;
}

extern "C" int eae_count_nb_deads_host(const float* data, uint32_t n, uint64_t hw, uint32_t nb_maps,
                                       uint32_t* nb_deads, void* stream)
{
    if (!data || !nb_deads) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (nb_maps == 0) { set_error("nb_maps is 0"); return EAE_ERR_ARGUMENT; }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t total = (uint64_t)n * hw * nb_maps;
    DevBuf d, alive;
    EAE_TRY(d.alloc(total * 4));
    EAE_TRY(alive.alloc((size_t)n * nb_maps * 4));
    EAE_CUDA_OK(cudaMemsetAsync(alive.p, 0, (size_t)n * nb_maps * 4, st));
    EAE_CUDA_OK(cudaMemcpyAsync(d.p, data, total * 4, cudaMemcpyHostToDevice, st));
    if (total) {
        alive_maps_kernel<<<grid_for(total, 4), kThreads, 0, st>>>(d.as<float>(), total, hw, nb_maps, alive.as<uint32_t>());
        EAE_LAUNCH_OK();
    }
    std::unique_ptr<uint32_t[]> h(new uint32_t[(size_t)n * nb_maps]);
    EAE_CUDA_OK(cudaMemcpyAsync(h.get(), alive.p, (size_t)n * nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n; i++) {
        uint32_t dead = 0;
        for (uint32_t c = 0; c < nb_maps; c++) dead += h[(size_t)i * nb_maps + c] ? 0u : 1u;
        nb_deads[i] = dead;
    }
    return 0;
}

extern "C" int eae_rescale_compress_lossless_maps_host(const float* cq_hwc, uint32_t h, uint32_t w,
                                                       uint32_t nb_maps, const float* bin_widths,
                                                       const double* table, uint32_t L,
                                                       const uint8_t* skip_mask, int16_t* idx_hwc_out,
                                                       uint32_t* nb_bits_each_map, void* stream)
{
    if (!cq_hwc || !bin_widths || !table || !nb_bits_each_map) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (L == 0) { set_error("truncated unary length is 0"); return EAE_ERR_UNARY_LENGTH; }
    if (L > 255 || nb_maps == 0 || (uint64_t)h * w == 0 || (uint64_t)h * w > 0x7FFFFFFFu / 64u) {
        set_error("bad shape / unary length"); return EAE_ERR_ARGUMENT;
    }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t size = h * w;
    const size_t n = (size_t)size * nb_maps;
    const uint32_t slot = eae_coder_slot_bytes(size, L);
    DevBuf q, bw, idx, planar, rec_planar, rec, flag, tbl, skip, bac, byp, bacb, bypb, err, off;
    EAE_TRY(q.alloc(n * 4));
    EAE_TRY(bw.alloc((size_t)nb_maps * 4));
    EAE_TRY(idx.alloc(n * 2));
    EAE_TRY(planar.alloc(n * 2));
    EAE_TRY(rec_planar.alloc(n * 2));
    EAE_TRY(rec.alloc(n * 2));
    EAE_TRY(flag.alloc(4));
    EAE_TRY(tbl.alloc((size_t)nb_maps * L * 8));
    EAE_TRY(bac.alloc((size_t)nb_maps * slot));
    EAE_TRY(byp.alloc((size_t)nb_maps * slot));
    EAE_TRY(bacb.alloc((size_t)nb_maps * 4));
    EAE_TRY(bypb.alloc((size_t)nb_maps * 4));
    EAE_TRY(err.alloc((size_t)nb_maps * 4));
    EAE_TRY(off.alloc((size_t)nb_maps * 8));
    if (skip_mask) {
        EAE_TRY(skip.alloc(nb_maps));
        EAE_CUDA_OK(cudaMemcpyAsync(skip.p, skip_mask, nb_maps, cudaMemcpyHostToDevice, st));
    }
    const uint8_t* skip_dev = skip_mask ? skip.as<uint8_t>() : nullptr;
    EAE_CUDA_OK(cudaMemsetAsync(flag.p, 0, 4, st));
    EAE_CUDA_OK(cudaMemcpyAsync(q.p, cq_hwc, n * 4, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(bw.p, bin_widths, (size_t)nb_maps * 4, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(tbl.p, table, (size_t)nb_maps * L * 8, cudaMemcpyHostToDevice, st));
    // compression.py:142: ref_int16 = cast_float_to_int16(q / tiled_bin_widths)
    EAE_TRY(launch_rescale_to_int16(q.as<float>(), idx.as<int16_t>(), size, nb_maps, bw.as<float>(),
                                    flag.as<uint32_t>(), st));
    EAE_TRY(launch_transpose_i16(idx.as<int16_t>(), planar.as<int16_t>(), 1, size, nb_maps, st));
    EAE_CUDA_OK(cudaMemcpyAsync(rec_planar.p, planar.p, n * 2, cudaMemcpyDeviceToDevice, st));
    EAE_TRY(launch_encode_streams(planar.as<int16_t>(), nb_maps, size, tbl.as<double>(), nb_maps, L, skip_dev,
                                  bac.as<uint8_t>(), byp.as<uint8_t>(), slot, bacb.as<uint32_t>(),
                                  bypb.as<uint32_t>(), err.as<uint32_t>(), st));
    std::unique_ptr<uint32_t[]> hb(new uint32_t[3 * (size_t)nb_maps]);
    uint32_t hflag = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(&hflag, flag.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get(), bacb.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get() + nb_maps, bypb.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get() + 2 * (size_t)nb_maps, err.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (hflag & 1u) {
        set_error("The rounded array elements cannot be represented as 16-bit signed integers.");
        return EAE_ERR_INT16_RANGE;
    }
    for (uint32_t i = 0; i < nb_maps; i++) {
        if (hb[2 * (size_t)nb_maps + i]) {
            set_error("Error of type %u during the encoding.", hb[2 * (size_t)nb_maps + i]);
            return (int)hb[2 * (size_t)nb_maps + i];
        }
        nb_bits_each_map[i] = hb[i] + hb[nb_maps + i];
    }
    EAE_TRY(launch_slot_offsets(off.as<uint64_t>(), nb_maps, slot, st));
    EAE_TRY(launch_decode_streams(rec_planar.as<int16_t>(), nb_maps, size, tbl.as<double>(), nb_maps, L, skip_dev,
                                  bac.as<uint8_t>(), off.as<uint64_t>(), bacb.as<uint32_t>(), byp.as<uint8_t>(),
                                  off.as<uint64_t>(), bypb.as<uint32_t>(), err.as<uint32_t>(), st));
    EAE_TRY(launch_transpose_i16(rec_planar.as<int16_t>(), rec.as<int16_t>(), 1, nb_maps, size, st));
    // compression.py:146-153: reconstruction = rec_int16 * delta must equal the input exactly.
    EAE_TRY(launch_check_rescaled(q.as<float>(), rec.as<int16_t>(), size, nb_maps, bw.as<float>(),
                                  flag.as<uint32_t>(), st));
    EAE_CUDA_OK(cudaMemcpyAsync(&hflag, flag.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get(), err.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    if (idx_hwc_out) EAE_CUDA_OK(cudaMemcpyAsync(idx_hwc_out, idx.p, n * 2, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < nb_maps; i++) {
        if (hb[i]) { set_error("Error of type %u during the decoding.", hb[i]); return (int)hb[i]; }
    }
    if (hflag & 2u) {
        set_error("The lossless compression has altered the centered quantized data.");
        return EAE_ERR_ROUND_TRIP;
    }
    return 0;
}

// =================================================================================================
// Latent statistics (kodak_tensorflow/lossless/stats.py): the counting part of save_statistics over a
// calibration set of latents. The float64 epilogues (probabilities, divergences) stay on the host, in
// numpy's order, in kodak_tensorflow/lossless/stats.py of this package.
namespace eae {
namespace {

// Order-preserving map float -> uint32 (for atomicMin / atomicMax on floats of either sign).
__device__ __forceinline__ uint32_t float_key(float x)
{
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float key_float(uint32_t k)
{
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// Pass 1: per-map sum (float64), minimum and maximum. One thread owns one map of a slab of rows.
__global__ void __launch_bounds__(256)
stats_moments_kernel(const float* __restrict__ y, uint64_t n_rows, uint32_t C, double* __restrict__ sum,
                     uint32_t* __restrict__ min_key, uint32_t* __restrict__ max_key, uint64_t rows_per_block)
{
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_block;
    const uint64_t r1 = r0 + rows_per_block < n_rows ? r0 + rows_per_block : n_rows;
    for (uint32_t c = threadIdx.x; c < C; c += blockDim.x) {
        double s = 0.;
        float mn = INFINITY, mx = -INFINITY;
        for (uint64_t r = r0; r < r1; r++) {
            const float v = __ldg(y + r * C + c);
            s += (double)v;
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        if (r1 > r0) {
            atomicAdd(&sum[c], s);
            atomicMin(&min_key[c], float_key(mn));
            atomicMax(&max_key[c], float_key(mx));
        }
    }
}

// Pass 2a: unit-interval histogram per map: bin floor(y - left[c]), the right edge belongs to the last
// bin (numpy.histogram, stats.py:118-127 with size_interval 1).
__global__ void __launch_bounds__(256)
stats_unit_hist_kernel(const float* __restrict__ y, uint64_t n_elems, uint32_t C, const double* __restrict__ left,
                       const uint32_t* __restrict__ nb_bins, unsigned long long* __restrict__ hist, uint32_t cap)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_elems;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(t % C);
        const double d = floor((double)__ldg(y + t) - left[c]);
        uint32_t b = d < 0. ? 0u : (uint32_t)d;
        if (b >= nb_bins[c]) b = nb_bins[c] - 1u;
        atomicAdd(&hist[(size_t)c * cap + b], 1ull);
    }
}

// Pass 2b: counts of a = min(|rint((y - mean[c]) / delta[c])|, L) per map (L + 1 bins), privatised per CTA in
// shared memory [L + 1][C].
__global__ void __launch_bounds__(256)
stats_abs_index_kernel(const float* __restrict__ y, uint64_t n_rows, uint32_t C, const float* __restrict__ mean,
                       const float* __restrict__ delta, uint32_t L, unsigned long long* __restrict__ counts,
                       uint64_t rows_per_block)
{
    extern __shared__ uint32_t sh_cnt[];
    for (uint32_t i = threadIdx.x; i < (L + 1u) * C; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_block;
    const uint64_t r1 = r0 + rows_per_block < n_rows ? r0 + rows_per_block : n_rows;
    const uint64_t e1 = r1 * C;
    for (uint64_t t = r0 * C + threadIdx.x; t < e1; t += blockDim.x) {
        const uint32_t c = (uint32_t)(t % C);
        // numpy float32 arithmetic of stats.py:42-44 / tools.py:927-929: (y - mean) / delta, round half to even
        const float k = rintf(__fdiv_rn(__fsub_rn(__ldg(y + t), __ldg(mean + c)), __ldg(delta + c)));
        const float a = fabsf(k);
        const uint32_t bin = a < (float)L ? (uint32_t)a : L;
        atomicAdd(&sh_cnt[bin * C + c], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (L + 1u) * C; i += blockDim.x) {
        const uint32_t v = sh_cnt[i];
        if (v) atomicAdd(&counts[(size_t)(i % C) * (L + 1u) + i / C], (unsigned long long)v);
    }
}

}  // namespace
}  // namespace eae

extern "C" int eae_latent_statistics_host(const float* y, uint64_t n_rows, uint32_t nb_maps, double* sum_out,
                                          float* min_out, float* max_out, uint64_t* unit_hist_out,
                                          uint32_t unit_cap, uint32_t* needed_cap, const float* mean,
                                          const float* delta, uint32_t L, uint64_t* abs_counts_out, void* stream)
{
    if (!y || !sum_out || !min_out || !max_out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (abs_counts_out && (!mean || !delta)) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (nb_maps == 0 || n_rows == 0) { set_error("empty latent array"); return EAE_ERR_ARGUMENT; }
    if (abs_counts_out && (L == 0 || L > 255)) { set_error("truncated unary length %u not in [1, 255]", L); return L == 0 ? EAE_ERR_UNARY_LENGTH : EAE_ERR_ARGUMENT; }
    if (abs_counts_out && n_rows > 0xFFFFFFFFull * 8) { set_error("too many rows"); return EAE_ERR_ARGUMENT; }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n_elems = n_rows * nb_maps;
    DevBuf dy, dsum, dmin, dmax;
    EAE_TRY(dy.alloc(n_elems * 4));
    EAE_TRY(dsum.alloc((size_t)nb_maps * 8));
    EAE_TRY(dmin.alloc((size_t)nb_maps * 4));
    EAE_TRY(dmax.alloc((size_t)nb_maps * 4));
    EAE_CUDA_OK(cudaMemcpyAsync(dy.p, y, n_elems * 4, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemsetAsync(dsum.p, 0, (size_t)nb_maps * 8, st));
    EAE_CUDA_OK(cudaMemsetAsync(dmin.p, 0xFF, (size_t)nb_maps * 4, st));
    EAE_CUDA_OK(cudaMemsetAsync(dmax.p, 0, (size_t)nb_maps * 4, st));
    const uint32_t blocks = (uint32_t)(n_rows < 148ull * 8 ? n_rows : 148ull * 8);
    const uint64_t rows_per_block = (n_rows + blocks - 1) / blocks;
    stats_moments_kernel<<<blocks, 256, 0, st>>>(dy.as<float>(), n_rows, nb_maps, dsum.as<double>(), dmin.as<uint32_t>(),
                                                 dmax.as<uint32_t>(), rows_per_block);
    EAE_LAUNCH_OK();
    std::unique_ptr<uint32_t[]> kmin(new uint32_t[nb_maps]), kmax(new uint32_t[nb_maps]);
    EAE_CUDA_OK(cudaMemcpyAsync(sum_out, dsum.p, (size_t)nb_maps * 8, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(kmin.get(), dmin.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(kmax.get(), dmax.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    uint32_t need = 1;
    std::unique_ptr<double[]> left(new double[nb_maps]);
    std::unique_ptr<uint32_t[]> nbins(new uint32_t[nb_maps]);
    for (uint32_t c = 0; c < nb_maps; c++) {
        min_out[c] = key_float(kmin[c]);
        max_out[c] = key_float(kmax[c]);
        left[c] = floor((double)min_out[c]);                       // stats.py:103-104
        const double width = ceil((double)max_out[c]) - left[c];
        nbins[c] = width >= 1. ? (width > 4e9 ? 0xFFFFFFFFu : (uint32_t)width) : 1u;
        if (nbins[c] > need) need = nbins[c];
    }
    if (needed_cap) *needed_cap = need;
    if (unit_hist_out) {
        if (need > unit_cap) { set_error("unit histogram range %u exceeds unit_cap %u", need, unit_cap); return EAE_ERR_ARGUMENT; }
        DevBuf dleft, dnb, dhist;
        EAE_TRY(dleft.alloc((size_t)nb_maps * 8));
        EAE_TRY(dnb.alloc((size_t)nb_maps * 4));
        EAE_TRY(dhist.alloc((size_t)nb_maps * unit_cap * 8));
        EAE_CUDA_OK(cudaMemcpyAsync(dleft.p, left.get(), (size_t)nb_maps * 8, cudaMemcpyHostToDevice, st));
        EAE_CUDA_OK(cudaMemcpyAsync(dnb.p, nbins.get(), (size_t)nb_maps * 4, cudaMemcpyHostToDevice, st));
        EAE_CUDA_OK(cudaMemsetAsync(dhist.p, 0, (size_t)nb_maps * unit_cap * 8, st));
        stats_unit_hist_kernel<<<grid_for(n_elems, 4), kThreads, 0, st>>>(dy.as<float>(), n_elems, nb_maps, dleft.as<double>(),
                                                                         dnb.as<uint32_t>(),
                                                                         dhist.as<unsigned long long>(), unit_cap);
        EAE_LAUNCH_OK();
        EAE_CUDA_OK(cudaMemcpyAsync(unit_hist_out, dhist.p, (size_t)nb_maps * unit_cap * 8, cudaMemcpyDeviceToHost, st));
        EAE_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (abs_counts_out) {
        DevBuf dmean, ddelta, dcnt;
        EAE_TRY(dmean.alloc((size_t)nb_maps * 4));
        EAE_TRY(ddelta.alloc((size_t)nb_maps * 4));
        EAE_TRY(dcnt.alloc((size_t)nb_maps * (L + 1u) * 8));
        EAE_CUDA_OK(cudaMemcpyAsync(dmean.p, mean, (size_t)nb_maps * 4, cudaMemcpyHostToDevice, st));
        EAE_CUDA_OK(cudaMemcpyAsync(ddelta.p, delta, (size_t)nb_maps * 4, cudaMemcpyHostToDevice, st));
        EAE_CUDA_OK(cudaMemsetAsync(dcnt.p, 0, (size_t)nb_maps * (L + 1u) * 8, st));
        const size_t smem = (size_t)(L + 1u) * nb_maps * 4;
        if (smem > 200u * 1024u) { set_error("nb_maps * (L + 1) too large for the counting kernel"); return EAE_ERR_ARGUMENT; }
        EAE_CUDA_OK(cudaFuncSetAttribute(stats_abs_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stats_abs_index_kernel<<<blocks, 256, smem, st>>>(dy.as<float>(), n_rows, nb_maps, dmean.as<float>(), ddelta.as<float>(), L,
                                                          dcnt.as<unsigned long long>(), rows_per_block);
        EAE_LAUNCH_OK();
        EAE_CUDA_OK(cudaMemcpyAsync(abs_counts_out, dcnt.p, (size_t)nb_maps * (L + 1u) * 8, cudaMemcpyDeviceToHost, st));
        EAE_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}
