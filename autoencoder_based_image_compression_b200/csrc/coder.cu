// Lossless coder of the quantized feature maps on the GPU, bit-exact with the reference C++ coder
// (kodak_tensorflow/lossless/c++/source/{LosslessCoder,BinaryArithmeticCoder,Bitstream}.cpp).
//
// Parallel decomposition: the arithmetic coder is strictly sequential inside a stream (one stream =
// one feature map of one image, compression.py:67-81), so ONE GPU LANE owns one stream and a warp
// advances 32 streams in lock-step. The encoder's main loop is flattened to "one prefix bin per
// iteration" so that lanes whose symbols need different numbers of bins stay converged.
// The split point uses the reference's arithmetic literally: FP64 multiply (round-to-nearest, never
// fused) followed by floor (BinaryArithmeticCoder.cpp:154).
#include <memory>
#include <stdlib.h>

#include "coder_core.cuh"
#include "common.cuh"
#include "internal.cuh"

namespace eae {

namespace {

// One stream per `lanes` consecutive threads (only the first of them works). lanes = 32 gives every
// stream its own warp: no divergence and the most warps in flight, which is what a small batch needs
// because the coder is latency-bound; lanes = 1 packs 32 streams per warp for the best issue efficiency
// when there are far more streams than warp slots. launch_* pick `lanes` from the stream count.
__global__ void __launch_bounds__(64, 8)
encode_streams_kernel(const int16_t* __restrict__ idx, uint32_t n_streams, uint32_t size,
                      const double* __restrict__ table, uint32_t table_rows, uint32_t L,
                      const uint8_t* __restrict__ skip_mask, uint8_t* __restrict__ bac_slots,
                      uint8_t* __restrict__ byp_slots, uint32_t slot_bytes, uint32_t cap_bits,
                      uint32_t* __restrict__ bac_bits, uint32_t* __restrict__ byp_bits,
                      uint32_t* __restrict__ err, uint32_t lanes)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t % lanes) return;
    const uint32_t s = t / lanes;
    if (s >= n_streams) return;
    const uint32_t row = s % table_rows;
    if (skip_mask && skip_mask[row]) {
        bac_bits[s] = 0; byp_bits[s] = 0; err[s] = 0;
        return;
    }
    core::BitSink bac, byp;
    bac.init(bac_slots + (size_t)s * slot_bytes, cap_bits);
    byp.init(byp_slots + (size_t)s * slot_bytes, cap_bits);
    const uint32_t e = core::encode_stream(idx + (size_t)s * size, size, table + (size_t)row * L, L, bac, byp);
    bac.flush();
    byp.flush();
    bac_bits[s] = bac.nbits;
    byp_bits[s] = byp.nbits;
    err[s] = e;
}

__global__ void __launch_bounds__(64, 8)
decode_streams_kernel(int16_t* __restrict__ out, uint32_t n_streams, uint32_t size,
                      const double* __restrict__ table, uint32_t table_rows, uint32_t L,
                      const uint8_t* __restrict__ skip_mask, const uint8_t* __restrict__ bac_base,
                      const uint64_t* __restrict__ bac_off, const uint32_t* __restrict__ bac_bits,
                      const uint8_t* __restrict__ byp_base, const uint64_t* __restrict__ byp_off,
                      const uint32_t* __restrict__ byp_bits, uint32_t* __restrict__ err, uint32_t lanes)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t % lanes) return;
    const uint32_t s = t / lanes;
    if (s >= n_streams) return;
    const uint32_t row = s % table_rows;
    if (skip_mask && skip_mask[row]) { err[s] = 0; return; }
    core::BitSource bac, byp;
    bac.init(bac_base + bac_off[s], bac_bits[s]);
    byp.init(byp_base + byp_off[s], byp_bits[s]);
    err[s] = core::decode_stream(out + (size_t)s * size, size, table + (size_t)row * L, L, bac, byp);
}

// Threads per stream: one warp per stream until that would exceed ~32 warps per SM, then halve.
inline uint32_t lanes_per_stream(uint32_t n_streams, uint32_t requested)
{
    static int forced = -1;   // env EAE_CODER_LANES (1, 2, 4, ... 32) overrides the heuristic
    if (forced < 0) {
        const char* env = getenv("EAE_CODER_LANES");
        forced = env ? atoi(env) : 0;
    }
    if (forced >= 1 && forced <= 32 && (forced & (forced - 1)) == 0) return (uint32_t)forced;
    if (requested >= 1 && requested <= 32 && (requested & (requested - 1)) == 0) return requested;
    uint32_t lanes = 32;
    while (lanes > 1 && (uint64_t)n_streams * lanes / 32 > 148ull * 32ull) lanes >>= 1;
    return lanes;
}

// ------------------------------------------------------------------------------------------------
// Layout: [n, hw, C] int16 <-> planar [n, C, hw], 32x32 tiles through shared memory.
__global__ void transpose_i16_kernel(const int16_t* __restrict__ in, int16_t* __restrict__ out,
                                     uint32_t rows, uint32_t cols)
{
    // in: [batch][rows][cols] -> out: [batch][cols][rows]
    __shared__ int16_t tile[32][33];
    const size_t base = (size_t)blockIdx.z * rows * cols;
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint32_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[base + (size_t)r * cols + c];
    }
    __syncthreads();
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint32_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[base + (size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------------------------------------
// Histograms of int16 symbols per map (tools.py count_symbols :322-388 on indices).
// hist id of element (img, pix, map): per_image ? img * C + map : map.
__global__ void minmax_abs_kernel(const int16_t* __restrict__ idx, uint64_t n_elems, uint32_t hw,
                                  uint32_t C, int per_image, int32_t* __restrict__ mn,
                                  int32_t* __restrict__ mx, unsigned long long* __restrict__ abs_sum)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_elems;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t map = (uint32_t)(t % C);
        const uint64_t img = t / ((uint64_t)hw * C);
        const uint64_t j = per_image ? img * C + map : map;
        const int v = idx[t];
        atomicMin(&mn[j], v);
        atomicMax(&mx[j], v);
        if (abs_sum && v != 0) atomicAdd(&abs_sum[j], (unsigned long long)(v < 0 ? -v : v));
    }
}

__global__ void hist_kernel(const int16_t* __restrict__ idx, uint64_t n_elems, uint32_t hw, uint32_t C,
                            int per_image, const int32_t* __restrict__ mn,
                            unsigned long long* __restrict__ hist, uint32_t cap)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_elems;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t map = (uint32_t)(t % C);
        const uint64_t img = t / ((uint64_t)hw * C);
        const uint64_t j = per_image ? img * C + map : map;
        const uint32_t b = (uint32_t)((int)idx[t] - mn[j]);
        if (b < cap) atomicAdd(&hist[j * cap + b], 1ull);
    }
}

__global__ void fill_i32_kernel(int32_t* p, uint64_t n, int32_t v)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (uint64_t)gridDim.x * blockDim.x) p[t] = v;
}

// Byte offsets of the per-stream slots (decoder input when reading straight from the slot arenas).
__global__ void slot_offsets_kernel(uint64_t* off, uint32_t n, uint32_t slot_bytes)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) off[s] = (uint64_t)s * slot_bytes;
}

}  // namespace

int launch_transpose_i16(const int16_t* in, int16_t* out, uint32_t batch, uint32_t rows, uint32_t cols,
                         cudaStream_t st)
{
    if (batch == 0 || rows == 0 || cols == 0) return 0;
    if (batch > 65535u) { set_error("transpose: batch %u exceeds grid.z", batch); return EAE_ERR_ARGUMENT; }
    dim3 grid(ceil_div_u32(cols, 32), ceil_div_u32(rows, 32), batch);
    transpose_i16_kernel<<<grid, dim3(32, 8), 0, st>>>(in, out, rows, cols);
    EAE_LAUNCH_OK();
    return 0;
}

// ---- internal launchers shared with codec.cu ----------------------------------------------------
uint32_t coder_capacity_bits(uint32_t size, uint32_t L)
{
    // compression.cpp:24 and Bitstream.cpp:3-7 (uint32 arithmetic, rounded up to a whole byte)
    uint32_t bits = size * (L > 32u ? L : 32u);
    uint32_t r = bits % 8u;
    return r ? bits + 8u - r : bits;
}

int launch_encode_streams(const int16_t* idx_planar, uint32_t n_streams, uint32_t size,
                          const double* table_dev, uint32_t table_rows, uint32_t L,
                          const uint8_t* skip_mask_dev, uint8_t* bac_slots, uint8_t* byp_slots,
                          uint32_t slot_bytes, uint32_t* bac_bits, uint32_t* byp_bits, uint32_t* err,
                          cudaStream_t st, uint32_t lanes_req)
{
    if (n_streams == 0) return 0;
    const uint32_t lanes = lanes_per_stream(n_streams, lanes_req);
    encode_streams_kernel<<<ceil_div_u32((uint64_t)n_streams * lanes, 64), 64, 0, st>>>(
        idx_planar, n_streams, size, table_dev, table_rows, L, skip_mask_dev, bac_slots, byp_slots,
        slot_bytes, coder_capacity_bits(size, L), bac_bits, byp_bits, err, lanes);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_decode_streams(int16_t* idx_planar_out, uint32_t n_streams, uint32_t size,
                          const double* table_dev, uint32_t table_rows, uint32_t L,
                          const uint8_t* skip_mask_dev, const uint8_t* bac_base, const uint64_t* bac_off,
                          const uint32_t* bac_bits, const uint8_t* byp_base, const uint64_t* byp_off,
                          const uint32_t* byp_bits, uint32_t* err, cudaStream_t st, uint32_t lanes_req)
{
    if (n_streams == 0) return 0;
    const uint32_t lanes = lanes_per_stream(n_streams, lanes_req);
    decode_streams_kernel<<<ceil_div_u32((uint64_t)n_streams * lanes, 64), 64, 0, st>>>(
        idx_planar_out, n_streams, size, table_dev, table_rows, L, skip_mask_dev, bac_base, bac_off,
        bac_bits, byp_base, byp_off, byp_bits, err, lanes);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_slot_offsets(uint64_t* off, uint32_t n, uint32_t slot_bytes, cudaStream_t st)
{
    if (n == 0) return 0;
    slot_offsets_kernel<<<ceil_div_u32(n, 256), 256, 0, st>>>(off, n, slot_bytes);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_histograms(const int16_t* idx_nhwc_dev, uint32_t n_images, uint32_t hw, uint32_t C,
                      int per_image, int32_t* mn, int32_t* mx, unsigned long long* abs_sum,
                      unsigned long long* hist, uint32_t cap, bool only_minmax, cudaStream_t st)
{
    const uint64_t n_elems = (uint64_t)n_images * hw * C;
    const uint64_t n_hist = per_image ? (uint64_t)n_images * C : C;
    if (n_elems == 0) return 0;
    const uint32_t grid = (uint32_t)((n_elems + 255) / 256 < 148u * 16u ? (n_elems + 255) / 256 : 148u * 16u);
    if (only_minmax) {
        fill_i32_kernel<<<ceil_div_u32(n_hist, 256), 256, 0, st>>>(mn, n_hist, 32767);
        EAE_LAUNCH_OK();
        fill_i32_kernel<<<ceil_div_u32(n_hist, 256), 256, 0, st>>>(mx, n_hist, -32768);
        EAE_LAUNCH_OK();
        if (abs_sum) EAE_CUDA_OK(cudaMemsetAsync(abs_sum, 0, n_hist * sizeof(unsigned long long), st));
        minmax_abs_kernel<<<grid, 256, 0, st>>>(idx_nhwc_dev, n_elems, hw, C, per_image, mn, mx, abs_sum);
        EAE_LAUNCH_OK();
    } else {
        EAE_CUDA_OK(cudaMemsetAsync(hist, 0, n_hist * cap * sizeof(unsigned long long), st));
        hist_kernel<<<grid, 256, 0, st>>>(idx_nhwc_dev, n_elems, hw, C, per_image, mn, hist, cap);
        EAE_LAUNCH_OK();
    }
    return 0;
}

}  // namespace eae

// =================================================================================================
// C ABI
using namespace eae;

extern "C" uint32_t eae_coder_capacity_bytes(uint32_t size, uint32_t L)
{
    return coder_capacity_bits(size, L) >> 3;
}

extern "C" uint32_t eae_coder_slot_bytes(uint32_t size, uint32_t L)
{
    uint32_t b = (coder_capacity_bits(size, L) >> 3);
    return ((b + 15u) / 16u) * 16u + 16u;
}

extern "C" int eae_nhwc_to_planar_i16_dev(const int16_t* nhwc, int16_t* planar, uint32_t n_images,
                                          uint32_t hw, uint32_t nb_maps, void* stream)
{
    if (!nhwc || !planar) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    return launch_transpose_i16(nhwc, planar, n_images, hw, nb_maps, (cudaStream_t)stream);
}

extern "C" int eae_planar_to_nhwc_i16_dev(const int16_t* planar, int16_t* nhwc, uint32_t n_images,
                                          uint32_t hw, uint32_t nb_maps, void* stream)
{
    if (!nhwc || !planar) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(require_device());
    return launch_transpose_i16(planar, nhwc, n_images, nb_maps, hw, (cudaStream_t)stream);
}

extern "C" int eae_encode_streams_dev(const int16_t* idx_planar, uint32_t n_streams, uint32_t size,
                                      const double* table_dev, uint32_t table_rows, uint32_t L,
                                      const uint8_t* skip_mask_dev, uint8_t* bac_slots,
                                      uint8_t* bypass_slots, uint32_t slot_bytes, uint32_t* bac_bits,
                                      uint32_t* bypass_bits, uint32_t* err, void* stream)
{
    if (!idx_planar || !table_dev || !bac_slots || !bypass_slots || !bac_bits || !bypass_bits || !err) {
        set_error("NULL pointer"); return EAE_ERR_NULL;
    }
    if (L == 0 || L > 255) { set_error("truncated unary length %u not in [1, 255]", L); return L == 0 ? EAE_ERR_UNARY_LENGTH : EAE_ERR_ARGUMENT; }
    if (table_rows == 0 || slot_bytes < eae_coder_slot_bytes(size, L) || (slot_bytes & 15u)) {
        set_error("bad table_rows / slot_bytes"); return EAE_ERR_ARGUMENT;
    }
    EAE_TRY(require_device());
    return launch_encode_streams(idx_planar, n_streams, size, table_dev, table_rows, L, skip_mask_dev,
                                 bac_slots, bypass_slots, slot_bytes, bac_bits, bypass_bits, err,
                                 (cudaStream_t)stream);
}

extern "C" int eae_decode_streams_dev(int16_t* idx_planar_out, uint32_t n_streams, uint32_t size,
                                      const double* table_dev, uint32_t table_rows, uint32_t L,
                                      const uint8_t* skip_mask_dev, const uint8_t* bac_base,
                                      const uint64_t* bac_off, const uint32_t* bac_bits,
                                      const uint8_t* byp_base, const uint64_t* byp_off,
                                      const uint32_t* bypass_bits, uint32_t* err, void* stream)
{
    if (!idx_planar_out || !table_dev || !bac_base || !bac_off || !bac_bits || !byp_base || !byp_off ||
        !bypass_bits || !err) {
        set_error("NULL pointer"); return EAE_ERR_NULL;
    }
    if (L == 0 || L > 255) { set_error("truncated unary length %u not in [1, 255]", L); return L == 0 ? EAE_ERR_UNARY_LENGTH : EAE_ERR_ARGUMENT; }
    if (table_rows == 0) { set_error("table_rows is 0"); return EAE_ERR_ARGUMENT; }
    EAE_TRY(require_device());
    return launch_decode_streams(idx_planar_out, n_streams, size, table_dev, table_rows, L, skip_mask_dev,
                                 bac_base, bac_off, bac_bits, byp_base, byp_off, bypass_bits, err,
                                 (cudaStream_t)stream);
}

namespace {

// Shared body of the host-side single/multi map entry points. `planar_in` host planar
// [n_streams, size]. Any of the outputs may be NULL.
struct HostCoderRun {
    DevBuf idx, out, table, skip, bac, byp, bacb, bypb, err, off;
    uint32_t slot = 0;

    int encode(const int16_t* planar_in_dev, uint32_t n_streams, uint32_t size, const double* table_host,
               uint32_t rows, uint32_t L, const uint8_t* skip_host, cudaStream_t st)
    {
        slot = eae_coder_slot_bytes(size, L);
        EAE_TRY(table.alloc((size_t)rows * L * sizeof(double)));
        EAE_CUDA_OK(cudaMemcpyAsync(table.p, table_host, (size_t)rows * L * sizeof(double),
                                    cudaMemcpyHostToDevice, st));
        if (skip_host) {
            EAE_TRY(skip.alloc(rows));
            EAE_CUDA_OK(cudaMemcpyAsync(skip.p, skip_host, rows, cudaMemcpyHostToDevice, st));
        }
        EAE_TRY(bac.alloc((size_t)n_streams * slot));
        EAE_TRY(byp.alloc((size_t)n_streams * slot));
        EAE_TRY(bacb.alloc((size_t)n_streams * 4));
        EAE_TRY(bypb.alloc((size_t)n_streams * 4));
        EAE_TRY(err.alloc((size_t)n_streams * 4));
        return launch_encode_streams(planar_in_dev, n_streams, size, table.as<double>(), rows, L,
                                     skip_host ? skip.as<uint8_t>() : nullptr, bac.as<uint8_t>(),
                                     byp.as<uint8_t>(), slot, bacb.as<uint32_t>(), bypb.as<uint32_t>(),
                                     err.as<uint32_t>(), st);
    }
    int decode_from_slots(int16_t* planar_out_dev, uint32_t n_streams, uint32_t size, uint32_t rows,
                          uint32_t L, bool have_skip, cudaStream_t st)
    {
        EAE_TRY(off.alloc((size_t)n_streams * 8));
        EAE_TRY(launch_slot_offsets(off.as<uint64_t>(), n_streams, slot, st));
        return launch_decode_streams(planar_out_dev, n_streams, size, table.as<double>(), rows, L,
                                     have_skip ? skip.as<uint8_t>() : nullptr, bac.as<uint8_t>(),
                                     off.as<uint64_t>(), bacb.as<uint32_t>(), byp.as<uint8_t>(),
                                     off.as<uint64_t>(), bypb.as<uint32_t>(), err.as<uint32_t>(), st);
    }
};

int first_error(const uint32_t* e, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) if (e[i]) return (int)e[i];
    return 0;
}

int check_unary_length(uint32_t L)
{
    if (L == 0) { set_error("truncated unary length is 0"); return EAE_ERR_UNARY_LENGTH; }
    if (L > 255) { set_error("truncated unary length %u exceeds 255", L); return EAE_ERR_ARGUMENT; }
    return 0;
}

}  // namespace

extern "C" int eae_encode_map_host(uint32_t size, const int16_t* in, uint8_t L, const double* probs,
                                   uint8_t* bac_bytes, uint32_t* bac_bits, uint8_t* byp_bytes,
                                   uint32_t* byp_bits)
{
    if (!in || !probs || !bac_bits || !byp_bits) { set_error("One of the pointers is NULL."); return EAE_ERR_NULL; }
    EAE_TRY(check_unary_length(L));
    EAE_TRY(require_device());
    cudaStream_t st = nullptr;
    HostCoderRun run;
    EAE_TRY(run.idx.alloc((size_t)size * 2));
    EAE_CUDA_OK(cudaMemcpyAsync(run.idx.p, in, (size_t)size * 2, cudaMemcpyHostToDevice, st));
    EAE_TRY(run.encode(run.idx.as<int16_t>(), 1, size, probs, 1, L, nullptr, st));
    uint32_t h[3];
    EAE_CUDA_OK(cudaMemcpyAsync(&h[0], run.bacb.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(&h[1], run.bypb.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(&h[2], run.err.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (h[2]) { set_error("Error of type %u during the encoding.", h[2]); return (int)h[2]; }
    *bac_bits = h[0];
    *byp_bits = h[1];
    if (bac_bytes) EAE_CUDA_OK(cudaMemcpy(bac_bytes, run.bac.p, (h[0] + 7) >> 3, cudaMemcpyDeviceToHost));
    if (byp_bytes) EAE_CUDA_OK(cudaMemcpy(byp_bytes, run.byp.p, (h[1] + 7) >> 3, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int eae_decode_map_host(uint32_t size, int16_t* out, uint8_t L, const double* probs,
                                   const uint8_t* bac_bytes, uint32_t bac_bits, const uint8_t* byp_bytes,
                                   uint32_t byp_bits)
{
    if (!out || !probs || !bac_bytes || !byp_bytes) { set_error("One of the pointers is NULL."); return EAE_ERR_NULL; }
    EAE_TRY(check_unary_length(L));
    EAE_TRY(require_device());
    cudaStream_t st = nullptr;
    DevBuf dout, table, bac, byp, meta, err;
    const uint32_t nb = (bac_bits + 7) >> 3, nr = (byp_bits + 7) >> 3;
    EAE_TRY(dout.alloc((size_t)size * 2));
    EAE_TRY(table.alloc((size_t)L * 8));
    EAE_TRY(bac.alloc(nb + 8));
    EAE_TRY(byp.alloc(nr + 8));
    EAE_TRY(meta.alloc(32));
    EAE_TRY(err.alloc(4));
    uint64_t zero_off = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(table.p, probs, (size_t)L * 8, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(bac.p, bac_bytes, nb, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(byp.p, byp_bytes, nr, cudaMemcpyHostToDevice, st));
    uint8_t* m = meta.as<uint8_t>();
    EAE_CUDA_OK(cudaMemcpyAsync(m, &zero_off, 8, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(m + 8, &bac_bits, 4, cudaMemcpyHostToDevice, st));
    EAE_CUDA_OK(cudaMemcpyAsync(m + 12, &byp_bits, 4, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_decode_streams(dout.as<int16_t>(), 1, size, table.as<double>(), 1, L, nullptr,
                                  bac.as<uint8_t>(), (const uint64_t*)m, (const uint32_t*)(m + 8),
                                  byp.as<uint8_t>(), (const uint64_t*)m, (const uint32_t*)(m + 12),
                                  err.as<uint32_t>(), st));
    uint32_t e = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(&e, err.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(out, dout.p, (size_t)size * 2, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (e) { set_error("Error of type %u during the decoding.", e); return (int)e; }
    return 0;
}

extern "C" int eae_compress_lossless(uint32_t size, const int16_t* in, int16_t* out, uint8_t L,
                                     const double* probs, uint32_t* nb_bits)
{
    // compression.cpp:9-12
    if (!in || !out || !probs || !nb_bits) { set_error("One of the three pointers is NULL."); return EAE_ERR_NULL; }
    EAE_TRY(check_unary_length(L));
    EAE_TRY(require_device());
    cudaStream_t st = nullptr;
    HostCoderRun run;
    EAE_TRY(run.idx.alloc((size_t)size * 2));
    EAE_TRY(run.out.alloc((size_t)size * 2));
    EAE_CUDA_OK(cudaMemcpyAsync(run.idx.p, in, (size_t)size * 2, cudaMemcpyHostToDevice, st));
    EAE_TRY(run.encode(run.idx.as<int16_t>(), 1, size, probs, 1, L, nullptr, st));
    uint32_t h[3];
    EAE_CUDA_OK(cudaMemcpyAsync(&h[0], run.bacb.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(&h[1], run.bypb.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(&h[2], run.err.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (h[2]) { set_error("Error of type %u during the encoding.", h[2]); return (int)h[2]; }
    *nb_bits = h[0] + h[1];   // compression.cpp:49
    EAE_TRY(run.decode_from_slots(run.out.as<int16_t>(), 1, size, 1, L, false, st));
    uint32_t e = 0;
    EAE_CUDA_OK(cudaMemcpyAsync(&e, run.err.p, 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(out, run.out.p, (size_t)size * 2, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (e) { set_error("Error of type %u during the decoding.", e); return (int)e; }
    return 0;
}

extern "C" int eae_compress_lossless_maps_host(const int16_t* ref_hwc, uint32_t h, uint32_t w,
                                               uint32_t nb_maps, const double* table, uint32_t L,
                                               const uint8_t* skip_mask, int16_t* rec_hwc,
                                               uint32_t* nb_bits_each_map, void* stream)
{
    if (!ref_hwc || !table || !rec_hwc || !nb_bits_each_map) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_unary_length(L));
    if (nb_maps == 0 || (uint64_t)h * w == 0 || (uint64_t)h * w > 0x7FFFFFFFu / 64u) {
        set_error("bad map shape %ux%ux%u", h, w, nb_maps); return EAE_ERR_ARGUMENT;
    }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t size = h * w;
    const size_t n_elems = (size_t)size * nb_maps;
    HostCoderRun run;
    DevBuf hwc, planar, rec_planar;
    EAE_TRY(hwc.alloc(n_elems * 2));
    EAE_TRY(planar.alloc(n_elems * 2));
    EAE_TRY(rec_planar.alloc(n_elems * 2));
    EAE_CUDA_OK(cudaMemcpyAsync(hwc.p, ref_hwc, n_elems * 2, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_transpose_i16(hwc.as<int16_t>(), planar.as<int16_t>(), 1, size, nb_maps, st));
    // Skipped maps are passed through unchanged (compression.py:75).
    EAE_CUDA_OK(cudaMemcpyAsync(rec_planar.p, planar.p, n_elems * 2, cudaMemcpyDeviceToDevice, st));
    EAE_TRY(run.encode(planar.as<int16_t>(), nb_maps, size, table, nb_maps, L, skip_mask, st));
    std::unique_ptr<uint32_t[]> hb(new uint32_t[3 * (size_t)nb_maps]);
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get(), run.bacb.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get() + nb_maps, run.bypb.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get() + 2 * (size_t)nb_maps, run.err.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    int e = first_error(hb.get() + 2 * (size_t)nb_maps, nb_maps);
    if (e) { set_error("Error of type %d during the encoding.", e); return e; }
    for (uint32_t i = 0; i < nb_maps; i++) nb_bits_each_map[i] = hb[i] + hb[nb_maps + i];
    EAE_TRY(run.decode_from_slots(rec_planar.as<int16_t>(), nb_maps, size, nb_maps, L, skip_mask != nullptr, st));
    EAE_TRY(launch_transpose_i16(rec_planar.as<int16_t>(), hwc.as<int16_t>(), 1, nb_maps, size, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hb.get(), run.err.p, (size_t)nb_maps * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(rec_hwc, hwc.p, n_elems * 2, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    e = first_error(hb.get(), nb_maps);
    if (e) { set_error("Error of type %d during the decoding.", e); return e; }
    return 0;
}

extern "C" int eae_histogram_maps_host(const int16_t* idx_nhwc, uint32_t n_images, uint32_t h, uint32_t w,
                                       uint32_t nb_maps, int per_image, int32_t* min_out, int32_t* max_out,
                                       uint64_t* hist_out, uint32_t hist_cap, uint32_t* needed_cap,
                                       uint64_t* abs_sum_out, void* stream)
{
    if (!idx_nhwc || !min_out || !max_out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (n_images == 0 || nb_maps == 0 || (uint64_t)h * w == 0 || (uint64_t)h * w > 0xFFFFFFFFull) {
        set_error("bad shape"); return EAE_ERR_ARGUMENT;
    }
    EAE_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t hw = h * w;
    const uint64_t n_elems = (uint64_t)n_images * hw * nb_maps;
    const uint64_t n_hist = per_image ? (uint64_t)n_images * nb_maps : nb_maps;
    DevBuf idx, mn, mx, as, hist;
    EAE_TRY(idx.alloc(n_elems * 2));
    EAE_TRY(mn.alloc(n_hist * 4));
    EAE_TRY(mx.alloc(n_hist * 4));
    EAE_TRY(as.alloc(n_hist * 8));
    EAE_CUDA_OK(cudaMemcpyAsync(idx.p, idx_nhwc, n_elems * 2, cudaMemcpyHostToDevice, st));
    EAE_TRY(launch_histograms(idx.as<int16_t>(), n_images, hw, nb_maps, per_image, mn.as<int32_t>(),
                              mx.as<int32_t>(), as.as<unsigned long long>(), nullptr, 0, true, st));
    EAE_CUDA_OK(cudaMemcpyAsync(min_out, mn.p, n_hist * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaMemcpyAsync(max_out, mx.p, n_hist * 4, cudaMemcpyDeviceToHost, st));
    if (abs_sum_out) EAE_CUDA_OK(cudaMemcpyAsync(abs_sum_out, as.p, n_hist * 8, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    uint32_t need = 1;
    for (uint64_t j = 0; j < n_hist; j++) {
        uint32_t r = (uint32_t)(max_out[j] - min_out[j] + 1);
        if (r > need) need = r;
    }
    if (needed_cap) *needed_cap = need;
    if (!hist_out) return 0;
    if (need > hist_cap) { set_error("histogram range %u exceeds hist_cap %u", need, hist_cap); return EAE_ERR_ARGUMENT; }
    EAE_TRY(hist.alloc(n_hist * hist_cap * 8));
    EAE_TRY(launch_histograms(idx.as<int16_t>(), n_images, hw, nb_maps, per_image, mn.as<int32_t>(),
                              nullptr, nullptr, hist.as<unsigned long long>(), hist_cap, false, st));
    EAE_CUDA_OK(cudaMemcpyAsync(hist_out, hist.p, n_hist * hist_cap * 8, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
