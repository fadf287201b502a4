// Lossless coder of the quantized feature maps on the GPU, bit-exact with the reference C++ coder
// (kodak_tensorflow/lossless/c++/source/{LosslessCoder,BinaryArithmeticCoder,Bitstream}.cpp).
//
// Parallel decomposition: the arithmetic coder is strictly sequential inside a stream (one stream =
// one feature map of one image, compression.py:67-81); everything else is not. A parallel pass (one warp per
// stream) binarises the symbols and writes the complete bypass stream, then ONE GPU LANE per stream runs the
// arithmetic coder over the stream's bin string, one branch-free step per bin, 32 streams per warp in lock-step.
// The split point is the reference's arithmetic: FP64 multiply (round-to-nearest, never fused) followed by floor
// (BinaryArithmeticCoder.cpp:154), or a fixed-point form proven equal for every range (prepare_table_kernel).
#include <memory>
#include <stdlib.h>

#include "coder_core.cuh"
#include "common.cuh"
#include "internal.cuh"

namespace eae {

namespace {

// =================================================================================================
//   encode = binarize_streams_kernel (parallel: one warp per stream writes the truncated-unary bit string
//            of the stream and its complete bypass stream)
//          + encode_streams3_kernel  (sequential: one thread per stream, one branch-free step per bin)
//   decode = decode_streams3_kernel  (one thread per stream: arithmetic decoding of all prefixes, then the
//            bypass pass over the same symbols)
//
// The per-bin step has no data-dependent loop and no symbol logic, so the streams of a warp stay converged: the
// kernels are bounded by the dependent chain of one step times the number of bins of the longest stream, and leave
// the SMs free for the transforms of other batches.
//
// Streams that share a warp are chosen to be the SAME feature map of different images (similar statistics,
// hence similar bin counts): slot j -> stream (j % group) * table_rows + j / group.
__global__ void __launch_bounds__(128)
binarize_streams_kernel(const int16_t* __restrict__ idx, uint32_t n_streams, uint32_t size,
                        uint32_t table_rows, uint32_t L, const uint8_t* __restrict__ skip_mask,
                        uint32_t* __restrict__ nbins, uint32_t* __restrict__ ubits, uint32_t uwords,
                        uint8_t* __restrict__ byp_slots, uint32_t slot_bytes, uint32_t* __restrict__ byp_bits)
{
    extern __shared__ uint32_t bin_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ub_words = L + 2u;          // 32 symbols x L bins + a carried partial word
    uint32_t* ub = bin_smem + warp * (ub_words + 34u);
    uint32_t* bb = ub + ub_words;              // 32 symbols x 32 bypass bits + a carried partial word
    const uint32_t s = blockIdx.x * (blockDim.x >> 5) + warp;
    if (s >= n_streams) return;
    if (skip_mask && skip_mask[s % table_rows]) {
        if (lane == 0) { nbins[s] = 0; byp_bits[s] = 0; }
        return;
    }
    for (uint32_t j = lane; j < ub_words; j += 32) ub[j] = 0;
    for (uint32_t j = lane; j < 34u; j += 32) bb[j] = 0;
    __syncwarp();
    const int16_t* src = idx + (size_t)s * size;
    uint32_t* uout = ubits + (size_t)s * uwords;
    uint32_t* bout = reinterpret_cast<uint32_t*>(byp_slots + (size_t)s * slot_bytes);
    uint32_t ucarry = 0, bcarry = 0;           // bits waiting in word 0 of the staging buffers
    uint32_t uw_out = 0, bw_out = 0;           // whole words written so far
    uint32_t total_bins = 0, total_byp = 0;
    for (uint32_t base = 0; base < size; base += 32) {
        const uint32_t i = base + lane;
        const bool have = i < size;
        const int v = have ? (int)__ldg(src + i) : 0;
        const uint32_t a = (uint32_t)(v < 0 ? -v : v);
        const uint32_t ones = have ? (a < L ? a : L) : 0u;
        const uint32_t nb = have ? ones + (a < L ? 1u : 0u) : 0u;
        uint32_t code = 0, cnt = 0;
        if (have) core::bypass_code(v, a, L, code, cnt);
        uint32_t un = nb, bn = cnt;            // inclusive scans over the warp
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tu = __shfl_up_sync(0xFFFFFFFFu, un, o), tb = __shfl_up_sync(0xFFFFFFFFu, bn, o);
            if ((int)lane >= o) { un += tu; bn += tb; }
        }
        const uint32_t utot = __shfl_sync(0xFFFFFFFFu, un, 31), btot = __shfl_sync(0xFFFFFFFFu, bn, 31);
        {   // `ones` one-bits from this symbol's first bin; its terminating zero is already there
            uint32_t pos = ucarry + un - nb, rem = ones;
            while (rem) {
                const uint32_t b = pos & 31u, c = (32u - b) < rem ? (32u - b) : rem;
                atomicOr(&ub[pos >> 5], (c == 32u ? 0xFFFFFFFFu : ((1u << c) - 1u)) << b);
                pos += c; rem -= c;
            }
        }
        if (cnt) {
            const uint32_t pos = bcarry + bn - cnt, b = pos & 31u;
            atomicOr(&bb[pos >> 5], code << b);
            if (b + cnt > 32u) atomicOr(&bb[(pos >> 5) + 1u], code >> (32u - b));
        }
        __syncwarp();
        const uint32_t ubits_now = ucarry + utot, ufull = ubits_now >> 5;
        const uint32_t bbits_now = bcarry + btot, bfull = bbits_now >> 5;
        for (uint32_t j = lane; j < ufull; j += 32) uout[uw_out + j] = ub[j];
        for (uint32_t j = lane; j < bfull; j += 32) bout[bw_out + j] = bb[j];
        const uint32_t ulast = ub[ufull], blast = bb[bfull];
        __syncwarp();
        for (uint32_t j = lane; j <= ufull; j += 32) ub[j] = 0;
        for (uint32_t j = lane; j <= bfull; j += 32) bb[j] = 0;
        __syncwarp();
        if (lane == 0) { ub[0] = ulast; bb[0] = blast; }
        __syncwarp();
        uw_out += ufull; ucarry = ubits_now & 31u;
        bw_out += bfull; bcarry = bbits_now & 31u;
        total_bins += utot; total_byp += btot;
    }
    if (lane == 0) {
        if (ucarry) uout[uw_out] = ub[0];
        if (bcarry) bout[bw_out] = bb[0];
        nbins[s] = total_bins;
        byp_bits[s] = total_byp;
    }
}

__device__ __forceinline__ uint32_t slot_to_stream(uint32_t j, uint32_t group, uint32_t table_rows)
{
    return group ? (j % group) * table_rows + j / group : j;
}

// =================================================================================================
// The sequential passes: the branch-free step of coder_core.cuh ("fast formulation").
// Per table row, `row_flags` says whether every probability is valid (bit 0: the loop then needs no error
// test at all) and whether the 48-bit fixed-point multipliers in `qtable` reproduce floor(p * range) for every
// range (bit 1, established exhaustively by prepare_table_kernel): rows without bit 0 take the lean loop,
// rows without bit 1 the FP64 multiply. Without a prepared table (row_flags == NULL) validity is checked
// here and the multiply is FP64.
__global__ void __launch_bounds__(256)
prepare_table_kernel(const double* __restrict__ table, uint32_t L, uint64_t* __restrict__ qtable,
                     uint8_t* __restrict__ row_flags)
{
    __shared__ uint32_t bad_p, bad_q;
    const uint32_t row = blockIdx.x;
    if (threadIdx.x == 0) { bad_p = 0; bad_q = 0; }
    __syncthreads();
    const double* prow = table + (size_t)row * L;
    for (uint32_t j = 0; j < L; j++) {
        const double p = prow[j];
        const bool ok = p > 0.0 && p < 1.0;
        const core::MulFp64 a{p};
        const core::MulFixed48 b{core::fixed48_of(p)};
        if (threadIdx.x == 0) { qtable[(size_t)row * L + j] = b.q; if (!ok) bad_p = 1; }
        if (!ok) continue;
        uint32_t diff = 0;
        for (uint32_t r = threadIdx.x; r <= 0xFFFFu; r += blockDim.x) diff |= a(r) ^ b(r);
        if (diff) bad_q = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) row_flags[row] = (uint8_t)((bad_p ? 0u : 1u) | ((bad_p || bad_q) ? 0u : 2u));
}

__device__ __forceinline__ bool row_is_valid(const double* prow, uint32_t L)
{
    bool ok = true;
    for (uint32_t j = 0; j < L; j++) { const double p = __ldg(prow + j); ok = ok && p > 0.0 && p < 1.0; }
    return ok;
}

// Fast encoder loop of one stream. The body of the inner loop is one basic block: the bits released by bin
// g - 1 enter the sink while the interval arithmetic of bin g runs (two independent dependent chains that the
// scheduler interleaves). Returns the largest number of bits one bin released.
template <typename Mul, typename T>
__device__ __forceinline__ uint32_t encode_bins_fast(const T* __restrict__ mrow, const uint32_t* __restrict__ uw,
                                                     uint32_t nb, uint32_t L, core::BacState& st, core::FastSink& bac)
{
    uint32_t k = 0, ev = 0, ec = 0, worst = 0;
    Mul mul{__ldg(mrow)};
    uint32_t wnext = nb ? __ldg(uw) : 0u;
    for (uint32_t g0 = 0; g0 < nb; g0 += 32u) {
        uint32_t w = wnext;
        if (g0 + 32u < nb) wnext = __ldg(uw + (g0 >> 5) + 1u);
        const uint32_t cnt = nb - g0 < 32u ? nb - g0 : 32u;
        #pragma unroll 1
        for (uint32_t b = 0; b < cnt; b++) {
            bac.put(ev, ec < 32u ? ec : 0u);
            const uint32_t bit = w & 1u;
            w >>= 1;
            k = (bit && k + 1u < L) ? k + 1u : 0u;
            const Mul next{__ldg(mrow + k)};     // for the next bin: off the dependent chain
            core::fast_encode_arith(st, bit, mul, ev, ec);
            worst = ec > worst ? ec : worst;
            mul = next;
        }
    }
    bac.put(ev, ec < 32u ? ec : 0u);
    return worst;
}

// The lean loop: rows with an invalid probability (errors in the reference's order) and streams in which one
// bin released more than 31 bits.
__device__ __noinline__ uint32_t encode_bins_lean(const double* __restrict__ prow, const uint32_t* __restrict__ uw,
                                                  uint32_t nb, uint32_t L, uint8_t* slot, uint32_t cap_bits,
                                                  uint32_t* nbits_out)
{
    core::BacState st = {0u, core::kRangeMax, 0u};
    core::BitSink bac;
    bac.init(slot, cap_bits);
    uint32_t e = 0, k = 0;
    for (uint32_t g = 0; g < nb && !e; g++) {
        const uint32_t bit = (__ldg(uw + (g >> 5)) >> (g & 31u)) & 1u;
        e = core::lean_encode_bin(st, bac, bit, __ldg(prow + k));
        k = (bit && k + 1u < L) ? k + 1u : 0u;
    }
    if (!e) e = core::bac_finish(st, bac);
    bac.flush();
    *nbits_out = bac.nbits;
    return e;
}

__global__ void __launch_bounds__(256)
encode_streams3_kernel(const uint32_t* __restrict__ nbins, const uint32_t* __restrict__ ubits, uint32_t uwords,
                       uint32_t n_streams, const double* __restrict__ table, const uint64_t* __restrict__ qtable,
                       const uint8_t* __restrict__ row_flags, uint32_t table_rows, uint32_t L,
                       const uint8_t* __restrict__ skip_mask, uint8_t* __restrict__ bac_slots,
                       uint32_t slot_bytes, uint32_t cap_bits, uint32_t* __restrict__ bac_bits,
                       uint32_t* __restrict__ err, uint32_t lanes, uint32_t group)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t % lanes) return;
    const uint32_t j = t / lanes;
    if (j >= n_streams) return;
    const uint32_t s = slot_to_stream(j, group, table_rows);
    const uint32_t row = s % table_rows;
    if (skip_mask && skip_mask[row]) { bac_bits[s] = 0; err[s] = 0; return; }
    const double* prow = table + (size_t)row * L;
    const uint32_t* uw = ubits + (size_t)s * uwords;
    const uint32_t nb = nbins[s];
    uint8_t* slot = bac_slots + (size_t)s * slot_bytes;
    const uint32_t flags = row_flags ? row_flags[row] : (row_is_valid(prow, L) ? 1u : 0u);
    uint32_t nbits = 0, e = 0;
    bool lean = !(flags & 1u);
    if (!lean) {
        core::BacState st = {0u, core::kRangeMax, 0u};
        core::FastSink bac;
        bac.init(slot, cap_bits);
        const uint32_t worst = (flags & 2u)
            ? encode_bins_fast<core::MulFixed48>(qtable + (size_t)row * L, uw, nb, L, st, bac)
            : encode_bins_fast<core::MulFp64>(prow, uw, nb, L, st, bac);
        core::fast_finish(st, bac);
        bac.flush();
        nbits = bac.pos();
        e = nbits > cap_bits ? core::kErrCapacity : 0u;
        lean = worst > 31u;
    }
    if (lean) e = encode_bins_lean(prow, uw, nb, L, slot, cap_bits, &nbits);
    bac_bits[s] = nbits;
    err[s] = e;
}

__device__ uint32_t g_empty_word = 0;   // stands in for a stream of zero bits

// Fast prefix decoder of one stream: an unchecked single-basic-block loop while at least 32 bits of the stream
// remain, then the checked step for the end of the stream.
template <typename Mul, typename T>
__device__ __forceinline__ void decode_prefixes_fast(const T* __restrict__ mrow, uint32_t L, uint32_t size,
                                                     core::FastSource& bac, int16_t* __restrict__ dst)
{
    core::DecState st;
    core::fast_decode_start(st, bac);
    uint32_t i = 0, a = 0, k = 0;
    const Mul m0{__ldg(mrow)};
    Mul mul = m0;
    int16_t* dptr = dst;
    #pragma unroll 1
    while (i < size && bac.left >= 32u) {
        // the next multiplier is entry 0 (symbol finished) or entry k + 1: fetch the latter before the bit is known
        const Mul up{__ldg(mrow + (k + 1u < L ? k + 1u : k))};
        const uint32_t bit = core::fast_decode_bin<false>(st, bac, mul);
        a += bit;
        const bool done = !bit || k == L - 1u;
        if (done) *dptr = (int16_t)a;            // one predicated store; the pointer only advances
        dptr += done ? 1 : 0;
        i += done ? 1u : 0u;
        a = done ? 0u : a;
        k = done ? 0u : k + 1u;
        mul = done ? m0 : up;
    }
    k = a;      // (a and k never differ: both count the ones of the current symbol; the checked loop below uses k)
    while (i < size) {
        const uint32_t bit = core::fast_decode_bin<true>(st, bac, mul);
        const bool done = !bit || k == L - 1u;
        if (done) { dst[i] = (int16_t)(k + bit); i++; }
        k = done ? 0u : k + 1u;
        mul = Mul{__ldg(mrow + k)};
    }
}

// Decoder phase B for one symbol whose prefix decoded to a0 != 0: Exp-Golomb suffix and sign from the bypass stream.
__device__ __forceinline__ uint32_t bypass_symbol(uint32_t a0, uint32_t L, core::BitSource& byp, int16_t* __restrict__ dst)
{
    int v;
    const uint32_t eb = core::lean_decode_bypass(a0, L, byp, v);
    if (!eb) *dst = (int16_t)v;
    return eb;
}

__global__ void __launch_bounds__(256)
decode_streams3_kernel(int16_t* __restrict__ out, uint32_t n_streams, uint32_t size,
                       const double* __restrict__ table, const uint64_t* __restrict__ qtable,
                       const uint8_t* __restrict__ row_flags, uint32_t table_rows, uint32_t L,
                       const uint8_t* __restrict__ skip_mask, const uint8_t* __restrict__ bac_base,
                       const uint64_t* __restrict__ bac_off, const uint32_t* __restrict__ bac_bits,
                       const uint8_t* __restrict__ byp_base, const uint64_t* __restrict__ byp_off,
                       const uint32_t* __restrict__ byp_bits, uint32_t* __restrict__ err, uint32_t lanes,
                       uint32_t group)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t % lanes) return;
    const uint32_t j = t / lanes;
    if (j >= n_streams) return;
    const uint32_t s = slot_to_stream(j, group, table_rows);
    const uint32_t row = s % table_rows;
    if (skip_mask && skip_mask[row]) { err[s] = 0; return; }
    const double* prow = table + (size_t)row * L;
    int16_t* dst = out + (size_t)s * size;
    const uint32_t flags = row_flags ? row_flags[row] : (row_is_valid(prow, L) ? 1u : 0u);
    uint32_t e = 0, n_ok = size;
    // phase A: the truncated-unary prefix of every symbol
    if (!(flags & 1u)) {
        core::BitSource bac;
        bac.init(bac_base + bac_off[s], bac_bits[s]);
        core::DecState st;
        core::lean_decode_start(st, bac);
        uint32_t i = 0, a = 0, k = 0;
        while (i < size) {
            const double p = __ldg(prow + k);
            if (!(p > 0.0 && p < 1.0)) { e = core::kErrProbability; n_ok = i; break; }
            const uint32_t bit = core::lean_decode_bin(st, bac, p);
            a += bit;
            const bool done = !bit || k == L - 1u;
            if (done) { dst[i] = (int16_t)a; i++; a = 0; }
            k = done ? 0u : k + 1u;
        }
    } else {
        // The stream is read one word every few bins, each lane from its own cache lines: bring the lines into L1
        // now so that no refill of the window waits for L2 / HBM in the middle of the dependent chain.
        {
            const uint8_t* sp = bac_base + bac_off[s];
            const uint32_t nbytes = (bac_bits[s] + 7u) >> 3;
            for (uint32_t o = 0; o < nbytes && o < 64u * 128u; o += 128u)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + o));
        }
        core::FastSource bac;
        bac.init(bac_base + bac_off[s], bac_bits[s], &g_empty_word);
        if (flags & 2u) decode_prefixes_fast<core::MulFixed48>(qtable + (size_t)row * L, L, size, bac, dst);
        else decode_prefixes_fast<core::MulFp64>(prow, L, size, bac, dst);
    }
    // phase B: EG0 suffixes and signs from the bypass stream. The magnitudes come back from global memory, where phase A
    // left them (an L2 round trip of ~300 cycles per read when they are fetched one by one: a quarter of the kernel on
    // peaked maps): eight per 16-byte load, the next eight requested while these are worked on; all-zero groups are skipped.
    core::BitSource byp;
    byp.init(byp_base + byp_off[s], byp_bits[s]);
    uint32_t eb = 0;      // (an error of phase A stands unless the bypass stream fails first, as in the one-by-one loop)
    uint32_t i = 0;
    const uint32_t head = min(n_ok, (uint32_t)(((16u - ((uint32_t)reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) >> 1));
    for (; i < head && !eb; i++) {
        const uint32_t a0 = (uint32_t)(uint16_t)dst[i];
        if (a0) eb = bypass_symbol(a0, L, byp, dst + i);
    }
    uint4 cur = make_uint4(0u, 0u, 0u, 0u);
    if (i + 8u <= n_ok) cur = *reinterpret_cast<const uint4*>(dst + i);
    while (i + 8u <= n_ok && !eb) {
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (i + 16u <= n_ok) nxt = *reinterpret_cast<const uint4*>(dst + i + 8u);
        if (cur.x | cur.y | cur.z | cur.w) {
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            #pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t a0 = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                if (a0 && !eb) eb = bypass_symbol(a0, L, byp, dst + i + j);
            }
        }
        cur = nxt;
        i += 8u;
    }
    for (; i < n_ok && !eb; i++) {
        const uint32_t a0 = (uint32_t)(uint16_t)dst[i];
        if (a0) eb = bypass_symbol(a0, L, byp, dst + i);
    }
    if (eb) e = eb;
    err[s] = e;
}

// Threads per stream slot of the sequential kernels (only the first thread of a slot works): as few warps as
// keep about one warp per SM sub-partition busy, so that a small batch still spreads over the whole GPU.
inline uint32_t lanes_v2(uint32_t n_streams, uint32_t requested)
{
    static int forced = -1;
    if (forced < 0) {
        const char* env = getenv("EAE_CODER_LANES");
        forced = env ? atoi(env) : 0;
    }
    if (forced >= 1 && forced <= 32 && (forced & (forced - 1)) == 0) return (uint32_t)forced;
    if (requested >= 1 && requested <= 32 && (requested & (requested - 1)) == 0) return requested;
    uint32_t lanes = 32;
    while (lanes > 1 && (uint64_t)n_streams * lanes / 32 > 148ull * 4ull) lanes >>= 1;
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
// Histograms of int16 symbols per map (tools.py count_symbols :322-388 on indices), on PLANAR streams
// (stream s = image * C + map holds its `size` symbols contiguously: the layout the coder reads).
// Warp-level kernels: one warp per stream, 16-byte loads (8 symbols per lane and instruction), no global atomics on
// the way: extrema by warp shuffles, counts in a per-warp shared-memory histogram that is added to the result once.
// hist id of stream s: per_image ? s : s % C.

// The stream as 16-byte vectors where it is aligned, scalars before and after: f(symbol) for every symbol.
template <typename F>
__device__ __forceinline__ void for_each_symbol(const int16_t* __restrict__ src, uint32_t size, int lane, F f)
{
    const uint32_t head = min(size, (uint32_t)(((16u - ((uint32_t)reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 1));
    if ((uint32_t)lane < head) f((int)src[lane]);
    const uint4* v = reinterpret_cast<const uint4*>(src + head);
    const uint32_t nvec = (size - head) >> 3;
    for (uint32_t i = lane; i < nvec; i += 32) {
        const uint4 q = __ldg(v + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        #pragma unroll
        for (int j = 0; j < 4; j++) { f((int)(int16_t)(w[j] & 0xFFFFu)); f((int)(int16_t)(w[j] >> 16)); }
    }
    const uint32_t done = head + (nvec << 3);
    if (done + lane < size) f((int)src[done + lane]);      // fewer than 8 symbols remain
}

__global__ void __launch_bounds__(256)
stream_minmax_kernel(const int16_t* __restrict__ idx, uint32_t n_streams, uint32_t size, int32_t* __restrict__ mn,
                     int32_t* __restrict__ mx, unsigned long long* __restrict__ abs_sum)
{
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    int lo = 32767, hi = -32768;
    unsigned long long sum = 0;
    for_each_symbol(idx + (size_t)s * size, size, lane, [&](int v) {
        lo = v < lo ? v : lo; hi = v > hi ? v : hi; sum += (unsigned long long)(v < 0 ? -v : v);
    });
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    }
    if (lane == 0) { mn[s] = lo; mx[s] = hi; if (abs_sum) abs_sum[s] = sum; }
}

// Per-map extrema over all images (per_image == 0): thread = map.
__global__ void reduce_over_images_kernel(const int32_t* __restrict__ mn_s, const int32_t* __restrict__ mx_s,
                                          const unsigned long long* __restrict__ abs_s, uint32_t n_images, uint32_t C,
                                          int32_t* __restrict__ mn, int32_t* __restrict__ mx,
                                          unsigned long long* __restrict__ abs_sum)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= C) return;
    int lo = 32767, hi = -32768;
    unsigned long long sum = 0;
    for (uint32_t i = 0; i < n_images; i++) {
        lo = min(lo, mn_s[(size_t)i * C + m]); hi = max(hi, mx_s[(size_t)i * C + m]);
        if (abs_s) sum += abs_s[(size_t)i * C + m];
    }
    mn[m] = lo; mx[m] = hi;
    if (abs_sum) abs_sum[m] = sum;
}

// Dynamic shared memory: warps x cap counters. Symbols outside [mn, mn + cap) are not counted (the caller sizes cap).
__global__ void __launch_bounds__(256)
stream_hist_kernel(const int16_t* __restrict__ idx, uint32_t n_streams, uint32_t size, uint32_t C, int per_image,
                   const int32_t* __restrict__ mn, unsigned long long* __restrict__ hist, uint32_t cap)
{
    extern __shared__ uint32_t counts[];
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;      // (whole warps leave together; no block-wide barrier below)
    uint32_t* mine = counts + (threadIdx.x >> 5) * cap;
    for (uint32_t b = lane; b < cap; b += 32) mine[b] = 0;
    __syncwarp();
    const uint32_t j = per_image ? s : s % C;
    const int lo = mn[j];
    for_each_symbol(idx + (size_t)s * size, size, lane, [&](int v) {
        const uint32_t b = (uint32_t)(v - lo);
        if (b < cap) atomicAdd(&mine[b], 1u);
    });
    __syncwarp();
    unsigned long long* dst = hist + (size_t)j * cap;
    for (uint32_t b = lane; b < cap; b += 32) {
        const uint32_t c = mine[b];
        if (!c) continue;
        if (per_image) dst[b] = c;                             // this warp is the only writer of histogram j
        else atomicAdd(&dst[b], (unsigned long long)c);
    }
}

// Ranges too wide for a shared-memory histogram per warp: global atomics, thread = symbol.
__global__ void stream_hist_wide_kernel(const int16_t* __restrict__ idx, uint64_t n_elems, uint32_t size, uint32_t C,
                                        int per_image, const int32_t* __restrict__ mn,
                                        unsigned long long* __restrict__ hist, uint32_t cap)
{
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_elems;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = t / size;
        const uint64_t j = per_image ? s : s % C;
        const uint32_t b = (uint32_t)((int)idx[t] - mn[j]);
        if (b < cap) atomicAdd(&hist[j * cap + b], 1ull);
    }
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

size_t coder_encode_scratch_bytes(uint32_t n_streams, uint32_t size, uint32_t L)
{
    // per stream: the truncated-unary bit string (at most L bins per symbol) in whole words, + the bin count
    const size_t uwords = ((size_t)size * L + 31) / 32 + 1;
    return ((size_t)n_streams * (uwords + 1) * 4 + 255) & ~(size_t)255;
}

// Debug (scripts/pipeline_probe.py): EAE_PROBE_SKIP bit 0 skips the binarisation launch, bit 1 the arithmetic
// encoder, bit 2 the decoder - timing experiments only, the results are garbage.
static int probe_skip()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("EAE_PROBE_SKIP"); v = e ? atoi(e) : 0; }
    return v;
}

// Threads per CTA of the one-thread-per-stream kernels (debug: EAE_CODER_BLOCK overrides).
static uint32_t coder_block(uint64_t threads)
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("EAE_CODER_BLOCK"); v = e ? atoi(e) : 0; }
    if (v == 32 || v == 64 || v == 128 || v == 256) return (uint32_t)v;
    return threads <= 148ull * 4ull * 32ull ? 32u : 64u;
}

// (see prefer_max_shared in common.cuh)
static void coder_carveouts()
{
    static uint64_t seen = 0;
    if (!first_use_on_device(&seen)) return;
    prefer_max_shared(encode_streams3_kernel); prefer_max_shared(decode_streams3_kernel);
    prefer_max_shared(binarize_streams_kernel);
    prefer_max_shared(stream_minmax_kernel); prefer_max_shared(stream_hist_kernel);
    prefer_max_shared(reduce_over_images_kernel); prefer_max_shared(stream_hist_wide_kernel);
}

int launch_encode_streams(const int16_t* idx_planar, uint32_t n_streams, uint32_t size,
                          const double* table_dev, uint32_t table_rows, uint32_t L,
                          const uint8_t* skip_mask_dev, uint8_t* bac_slots, uint8_t* byp_slots,
                          uint32_t slot_bytes, uint32_t* bac_bits, uint32_t* byp_bits, uint32_t* err,
                          cudaStream_t st, uint32_t lanes_req, void* scratch, const uint64_t* qtable_dev,
                          const uint8_t* row_flags_dev)
{
    if (n_streams == 0) return 0;
    coder_carveouts();
    void* own = nullptr;
    if (!scratch) {     // callers without a persistent workspace (C-ABI stream entry points)
        EAE_CUDA_OK(cudaMallocAsync(&own, coder_encode_scratch_bytes(n_streams, size, L), st));
        scratch = own;
    }
    const uint32_t uwords = (uint32_t)(((size_t)size * L + 31) / 32 + 1);
    uint32_t* nbins = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* ubits = nbins + n_streams;
    const uint32_t warps = 4;
    const size_t smem = (size_t)warps * (L + 2u + 34u) * 4;
    ProfScope prof_binarize(kProfBinarize, st);
    if (!(probe_skip() & 1))
    binarize_streams_kernel<<<ceil_div_u32(n_streams, warps), warps * 32, smem, st>>>(
        idx_planar, n_streams, size, table_rows, L, skip_mask_dev, nbins, ubits, uwords, byp_slots, slot_bytes,
        byp_bits);
    EAE_LAUNCH_OK();
    prof_binarize.close();
    const uint32_t lanes = lanes_v2(n_streams, lanes_req);
    const uint32_t group = (n_streams % table_rows == 0) ? n_streams / table_rows : 0u;
    const uint64_t threads = (uint64_t)n_streams * lanes;
    const uint32_t block = coder_block(threads);   // few warps: one per CTA, spread over the SMs
    if (probe_skip() & 2) { if (own) EAE_CUDA_OK(cudaFreeAsync(own, st)); return 0; }
    encode_streams3_kernel<<<ceil_div_u32(threads, block), block, 0, st>>>(
            nbins, ubits, uwords, n_streams, table_dev, qtable_dev, qtable_dev ? row_flags_dev : nullptr, table_rows, L,
            skip_mask_dev, bac_slots, slot_bytes, coder_capacity_bits(size, L), bac_bits, err, lanes, group);
    EAE_LAUNCH_OK();
    if (own) EAE_CUDA_OK(cudaFreeAsync(own, st));
    return 0;
}

int launch_decode_streams(int16_t* idx_planar_out, uint32_t n_streams, uint32_t size,
                          const double* table_dev, uint32_t table_rows, uint32_t L,
                          const uint8_t* skip_mask_dev, const uint8_t* bac_base, const uint64_t* bac_off,
                          const uint32_t* bac_bits, const uint8_t* byp_base, const uint64_t* byp_off,
                          const uint32_t* byp_bits, uint32_t* err, cudaStream_t st, uint32_t lanes_req,
                          const uint64_t* qtable_dev, const uint8_t* row_flags_dev)
{
    if (n_streams == 0) return 0;
    coder_carveouts();
    const uint32_t lanes = lanes_v2(n_streams, lanes_req);
    const uint32_t group = (n_streams % table_rows == 0) ? n_streams / table_rows : 0u;
    const uint64_t threads = (uint64_t)n_streams * lanes;
    const uint32_t block = coder_block(threads);
    if (probe_skip() & 4) return 0;
    decode_streams3_kernel<<<ceil_div_u32(threads, block), block, 0, st>>>(
            idx_planar_out, n_streams, size, table_dev, qtable_dev, qtable_dev ? row_flags_dev : nullptr, table_rows, L,
            skip_mask_dev, bac_base, bac_off, bac_bits, byp_base, byp_off, byp_bits, err, lanes, group);
    EAE_LAUNCH_OK();
    return 0;
}

int launch_prepare_table(const double* table_dev, uint32_t rows, uint32_t L, uint64_t* qtable_dev,
                         uint8_t* row_flags_dev, cudaStream_t st)
{
    if (rows == 0) return 0;
    prepare_table_kernel<<<rows, 256, 0, st>>>(table_dev, L, qtable_dev, row_flags_dev);
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

int launch_histograms(const int16_t* idx_planar_dev, uint32_t n_images, uint32_t hw, uint32_t C,
                      int per_image, int32_t* mn, int32_t* mx, unsigned long long* abs_sum,
                      unsigned long long* hist, uint32_t cap, bool only_minmax, cudaStream_t st)
{
    const uint32_t n_streams = n_images * C;
    const uint64_t n_hist = per_image ? (uint64_t)n_streams : C;
    if (n_streams == 0 || hw == 0) return 0;
    coder_carveouts();
    ProfScope prof(kProfHist, st);
    if (only_minmax) {
        if (per_image) {
            stream_minmax_kernel<<<ceil_div_u32((uint64_t)n_streams * 32, 256), 256, 0, st>>>(idx_planar_dev, n_streams, hw, mn, mx, abs_sum);
            EAE_LAUNCH_OK();
            return 0;
        }
        // per stream first, then per map over the images
        int32_t* tmp = nullptr;
        EAE_CUDA_OK(cudaMallocAsync(&tmp, (size_t)n_streams * 16, st));
        int32_t* mn_s = tmp;
        int32_t* mx_s = tmp + n_streams;
        unsigned long long* abs_s = reinterpret_cast<unsigned long long*>(tmp + 2 * (size_t)n_streams);
        stream_minmax_kernel<<<ceil_div_u32((uint64_t)n_streams * 32, 256), 256, 0, st>>>(idx_planar_dev, n_streams, hw, mn_s, mx_s,
                                                                                         abs_sum ? abs_s : nullptr);
        EAE_LAUNCH_OK();
        reduce_over_images_kernel<<<ceil_div_u32(C, 128), 128, 0, st>>>(mn_s, mx_s, abs_sum ? abs_s : nullptr, n_images, C, mn, mx, abs_sum);
        EAE_LAUNCH_OK();
        EAE_CUDA_OK(cudaFreeAsync(tmp, st));
        return 0;
    }
    EAE_CUDA_OK(cudaMemsetAsync(hist, 0, n_hist * cap * sizeof(unsigned long long), st));
    // as many warps per CTA as fit 48 KB of counters (no opt-in needed); beyond 12 288 bins: global atomics
    uint32_t warps = cap ? (48u * 1024u) / (cap * 4u) : 8u;
    if (warps > 8u) warps = 8u;
    if (warps == 0u) {
        const uint64_t n_elems = (uint64_t)n_streams * hw;
        const uint32_t grid = (uint32_t)((n_elems + 255) / 256 < 148u * 16u ? (n_elems + 255) / 256 : 148u * 16u);
        stream_hist_wide_kernel<<<grid, 256, 0, st>>>(idx_planar_dev, n_elems, hw, C, per_image, mn, hist, cap);
        EAE_LAUNCH_OK();
        return 0;
    }
    stream_hist_kernel<<<ceil_div_u32(n_streams, warps), warps * 32, (size_t)warps * cap * 4, st>>>(
        idx_planar_dev, n_streams, hw, C, per_image, mn, hist, cap);
    EAE_LAUNCH_OK();
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

extern "C" int eae_histogram_streams_dev(const int16_t* idx_planar_dev, uint32_t n_images, uint32_t size, uint32_t nb_maps,
                                         int per_image, int32_t* min_dev, int32_t* max_dev, uint64_t* abs_sum_dev,
                                         uint64_t* hist_dev, uint32_t hist_cap, void* stream)
{
    if (!idx_planar_dev || !min_dev || (!hist_dev && !max_dev)) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (hist_dev && hist_cap == 0) { set_error("hist_cap is 0"); return EAE_ERR_ARGUMENT; }
    EAE_TRY(require_device());
    return launch_histograms(idx_planar_dev, n_images, size, nb_maps, per_image, min_dev, max_dev,
                             reinterpret_cast<unsigned long long*>(abs_sum_dev), reinterpret_cast<unsigned long long*>(hist_dev),
                             hist_cap, hist_dev == nullptr, (cudaStream_t)stream);
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
    DevBuf nhwc, idx, mn, mx, as, hist;
    EAE_TRY(nhwc.alloc(n_elems * 2));
    EAE_TRY(idx.alloc(n_elems * 2));
    EAE_TRY(mn.alloc(n_hist * 4));
    EAE_TRY(mx.alloc(n_hist * 4));
    EAE_TRY(as.alloc(n_hist * 8));
    EAE_CUDA_OK(cudaMemcpyAsync(nhwc.p, idx_nhwc, n_elems * 2, cudaMemcpyHostToDevice, st));
    // planar streams [image * nb_maps + map][h * w]: the layout the warp-level kernels (and the coder) read
    EAE_TRY(launch_transpose_i16(nhwc.as<int16_t>(), idx.as<int16_t>(), n_images, hw, nb_maps, st));
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
