// Per-stream core of the lossless coder, shared by the CUDA kernels (coder.cu) and, compiled as plain
// host C++ (-DEAE_HOST_ONLY), by the CPU test harness that checks its control flow against the oracle
// without a GPU (tests/host_harness). Bit-exact with the reference
// (kodak_tensorflow/lossless/c++/source/{LosslessCoder,BinaryArithmeticCoder,Bitstream}.cpp).
//
// Differences from a literal transcription, all output-preserving:
//  * E1/E2 rescalings are applied in one step: while the MSBs of low and high agree the reference
//    shifts one bit at a time (BinaryArithmeticCoder.cpp:213-236); the number of such shifts is the
//    number of leading equal bits, found with one count-leading-zeros. After them the MSBs differ and
//    only E3 steps (:238-246) can follow, and an E3 step never makes the MSBs equal again.
//  * the bypass bits of a symbol (EG0 suffix + sign, LosslessCoder.cpp:22-37, 58-111) are assembled in
//    a register and appended with one call.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__) && !defined(EAE_HOST_ONLY)
#define EAE_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define EAE_HD inline
#endif

namespace eae {
namespace core {

constexpr uint32_t kRangeMax = 0xFFFFu;       // BinaryArithmeticCoder.cpp:14
constexpr uint32_t kHalf = 0x7FFFu;           // :20
constexpr uint32_t kQuarter = 0x3FFFu;        // :26
constexpr uint32_t kThreeQuarters = 0xBFFDu;  // :27  (3 * 0x3FFF, not 0xBFFF)
constexpr uint32_t kMsb = 0x8000u;            // :33

constexpr uint32_t kErrCapacity = 1, kErrResource = 2, kErrPrecision = 3, kErrProbability = 4;

EAE_HD uint32_t clz32(uint32_t x)   // x != 0
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)x);
#else
    return (uint32_t)__builtin_clz(x);
#endif
}

EAE_HD uint32_t brev32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// The low `n` (1..32) bits of x in reversed order.
EAE_HD uint32_t reverse_bits(uint32_t x, uint32_t n) { return brev32(x) >> (32u - n); }

// (uint32_t)floor(p * range) with the reference's arithmetic: FP64 multiply rounded to nearest, never
// fused with anything, then floor (BinaryArithmeticCoder.cpp:154).
EAE_HD uint32_t mul_floor(double p, uint32_t range)
{
#ifdef __CUDA_ARCH__
    return __double2uint_rd(__dmul_rn(p, (double)range));
#else
    volatile double prod = p * (double)range;
    return (uint32_t)floor(prod);
#endif
}

template <typename T> EAE_HD T load_ro(const T* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// LSB-first bit writer: bit i of the stream is bit (i & 7) of byte (i >> 3) (Bitstream.cpp:36-58),
// i.e. bit (i & 31) of little-endian 32-bit word (i >> 5). The slot is 4-byte aligned and a whole
// number of words long.
struct BitSink {
    uint32_t* words;
    uint32_t cap_bits, nbits, widx, fill, acc;

    EAE_HD void init(uint8_t* slot, uint32_t cap)
    {
        words = reinterpret_cast<uint32_t*>(slot);
        cap_bits = cap; nbits = 0; widx = 0; fill = 0; acc = 0;
    }
    // Appends the `count` (1..32) low bits of `value`, first bit = bit 0. False on overflow
    // (Bitstream.cpp:38-41: capacity_error as soon as one bit does not fit).
    EAE_HD bool put(uint32_t value, uint32_t count)
    {
        if (nbits + count > cap_bits) return false;
        nbits += count;
        uint64_t wide = ((uint64_t)value << fill) | acc;
        fill += count;
        if (fill >= 32u) {
            words[widx++] = (uint32_t)wide;
            wide >>= 32;
            fill -= 32u;
        }
        acc = (uint32_t)wide;
        return true;
    }
    // `repeat` copies of `bit`.
    EAE_HD bool put_run(uint32_t bit, uint32_t repeat)
    {
        while (repeat) {
            const uint32_t c = repeat < 32u ? repeat : 32u;
            const uint32_t ones = c == 32u ? 0xFFFFFFFFu : ((1u << c) - 1u);
            if (!put(bit ? ones : 0u, c)) return false;
            repeat -= c;
        }
        return true;
    }
    EAE_HD void flush() { if (fill) words[widx] = acc; }
};

// LSB-first bit reader over an arbitrarily aligned byte range, fetching aligned 32-bit words. A word is
// only fetched when it holds at least one bit of the stream, so reads stay within the 4-byte-aligned
// span that contains the stream.
struct BitSource {
    const uint32_t* words;
    uint32_t nbits, rd, widx, fill;
    uint64_t acc;

    EAE_HD void init(const uint8_t* p, uint32_t bits)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        words = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        nbits = bits; rd = 0; widx = 0; fill = 0; acc = 0;
        if (bits) {
            const uint32_t skip = (uint32_t)(a & 3u) * 8u;
            acc = (uint64_t)(load_ro(words) >> skip);
            fill = 32u - skip;
            widx = 1;
        }
    }
    EAE_HD uint32_t remaining() const { return nbits - rd; }
    // Next n (1..16) bits, first bit at bit 0. Caller guarantees n <= remaining().
    EAE_HD uint32_t take(uint32_t n)
    {
        if (fill < n) {
            acc |= (uint64_t)load_ro(words + widx) << fill;
            widx++;
            fill += 32u;
        }
        const uint32_t v = (uint32_t)acc & ((1u << n) - 1u);
        acc >>= n;
        fill -= n;
        rd += n;
        return v;
    }
};

// Arithmetic-coder registers (BinaryArithmeticCoder.h:12-16).
struct BacState {
    uint32_t low, high, pending;
};

// encoding(bit, p) = encode_bit + rescale_encoding (BinaryArithmeticCoder.cpp:49-59, 158-252).
// Returns 0 or an error code.
EAE_HD uint32_t bac_encode(BacState& s, BitSink& out, uint32_t bit, double p)
{
    if (!(p > 0.0 && p < 1.0)) return kErrProbability;   // also catches NaN (:146-153)
    const uint32_t mid = s.low + mul_floor(p, s.high - s.low);
    uint32_t low = bit ? mid + 1u : s.low;
    uint32_t high = bit ? s.high : mid;
    if (high > kRangeMax || low > kRangeMax) return kErrPrecision;
    const uint32_t x = low ^ high;
    if ((x & kMsb) == 0u) {
        // n = number of leading equal bits of the 16-bit registers (1..16): n E1/E2 steps at once.
        const uint32_t n = clz32((x << 16) | 0x8000u);
        const uint32_t top = low >> (16u - n);            // the n bits to emit, first one is the MSB
        low = (low << n) & kRangeMax;
        high = ((high << n) & kRangeMax) | ((1u << n) - 1u);
        if (s.pending == 0u) {
            if (!out.put(reverse_bits(top, n), n)) return kErrCapacity;
        } else {
            // first bit, then the queued E3 bits (inverted, :317-337), then the other n - 1 bits
            const uint32_t first = (top >> (n - 1u)) & 1u;
            if (!out.put(first, 1u)) return kErrCapacity;
            if (!out.put_run(first ^ 1u, s.pending)) return kErrCapacity;
            s.pending = 0u;
            if (n > 1u && !out.put(reverse_bits(top & ((1u << (n - 1u)) - 1u), n - 1u), n - 1u)) return kErrCapacity;
        }
    }
    while (low > kQuarter && high <= kThreeQuarters) {    // E3 (:238-246); MSBs differ here
        low = (low - (kQuarter + 1u)) << 1;
        high = ((high - (kQuarter + 1u)) << 1) | 1u;
        s.pending++;
    }
    s.low = low;
    s.high = high;
    return 0u;
}

// stop_encoding (BinaryArithmeticCoder.cpp:61-102).
EAE_HD uint32_t bac_finish(BacState& s, BitSink& out)
{
    const uint32_t bit = s.low < kQuarter ? 0u : 1u;
    if (!out.put(bit, 1u)) return kErrCapacity;
    if (!out.put_run(bit ^ 1u, s.pending + 1u)) return kErrCapacity;
    return 0u;
}

// Bypass bits of one symbol: EG0(a - L) if a >= L, then the sign (0 = negative) if v != 0.
// At most 31 + 1 bits; returned in emission order from bit 0.
EAE_HD void bypass_code(int v, uint32_t a, uint32_t L, uint32_t& code, uint32_t& cnt)
{
    code = 0u;
    cnt = 0u;
    if (a >= L) {
        const uint32_t x1 = a - L + 1u;
        const uint32_t n = 31u - clz32(x1);
        code = (1u << n) - 1u;                              // n ones, then a zero at position n
        if (n) code |= reverse_bits(x1 - (1u << n), n) << (n + 1u);   // n suffix bits, MSB first
        cnt = 2u * n + 1u;
    }
    if (v != 0) {
        code |= (v > 0 ? 1u : 0u) << cnt;
        cnt++;
    }
}

// Whole-stream encoder: LosslessCoder::write_signed_ueg0 for every symbol (LosslessCoder.cpp:232-252),
// then stop_encoding. The loop advances ONE prefix bin per iteration so that GPU lanes working on
// different streams stay converged. Returns the error code; bit counts are in the sinks.
EAE_HD uint32_t encode_stream(const int16_t* src, uint32_t size, const double* prob, uint32_t L,
                              BitSink& bac, BitSink& byp)
{
    BacState st = {0u, kRangeMax, 0u};
    uint32_t i = 0, ones = 0, nb = 0, bin = 0;
    int next_v = size ? (int)load_ro(src) : 0;
    double p = load_ro(prob);
    for (;;) {
        if (bin == nb) {
            if (i == size) break;
            const int v = next_v;
            i++;
            if (i < size) next_v = (int)load_ro(src + i);
            const uint32_t a = (uint32_t)(v < 0 ? -v : v);
            ones = a < L ? a : L;
            nb = ones + (a < L ? 1u : 0u);
            bin = 0;
            uint32_t code, cnt;
            bypass_code(v, a, L, code, cnt);
            if (cnt && !byp.put(code, cnt)) return kErrCapacity;
        }
        const uint32_t e = bac_encode(st, bac, bin < ones ? 1u : 0u, p);
        if (e) return e;
        bin++;
        p = load_ro(prob + (bin < nb ? bin : 0u));
    }
    return bac_finish(st, bac);
}

// Whole-stream decoder: start_decoding (BinaryArithmeticCoder.cpp:104-122) then read_signed_ueg0 per
// symbol (LosslessCoder.cpp:254-276). `sticky`-bit behaviour of an exhausted stream is reproduced:
// start_decoding pads with the last bit read, rescale_decoding with the last bit read IN THAT CALL
// (or 0).
EAE_HD uint32_t decode_stream(int16_t* dst, uint32_t size, const double* prob, uint32_t L,
                              BitSource& bac, BitSource& byp)
{
    uint32_t low = 0, high = kRangeMax, code = 0;
    {
        const uint32_t have = bac.remaining() < 16u ? bac.remaining() : 16u;
        uint32_t keep = 0;
        if (have) {
            const uint32_t bits = bac.take(have);
            code = reverse_bits(bits, have);
            keep = (bits >> (have - 1u)) & 1u;
        }
        for (uint32_t k = have; k < 16u; k++) code = (code << 1) | keep;
    }
    for (uint32_t i = 0; i < size; i++) {
        uint32_t a = 0, bit = 0;
        for (uint32_t bin = 0;; bin++) {
            const double p = load_ro(prob + bin);
            if (!(p > 0.0 && p < 1.0)) return kErrProbability;
            const uint32_t mid = low + mul_floor(p, high - low);
            // decode_bit (:254-273): `bit` keeps its previous value when code is outside [low, high].
            if (code >= low && code <= mid) { high = mid; bit = 0; }
            else if (code > mid && code <= high) { low = mid + 1u; bit = 1; }
            // rescale_decoding (:275-315). E1/E2 in one step, then E3 steps.
            uint32_t in = 0;
            const uint32_t x = low ^ high;
            if (high <= kRangeMax && low <= kRangeMax && (x & kMsb) == 0u) {
                const uint32_t n = clz32((x << 16) | 0x8000u);
                low = (low << n) & kRangeMax;
                high = ((high << n) & kRangeMax) | ((1u << n) - 1u);
                const uint32_t have = bac.remaining() < n ? bac.remaining() : n;
                uint32_t fresh = 0;
                if (have) {
                    const uint32_t bits = bac.take(have);
                    fresh = reverse_bits(bits, have);
                    in = (bits >> (have - 1u)) & 1u;
                }
                for (uint32_t k = have; k < n; k++) fresh = (fresh << 1) | in;
                code = ((code << n) & kRangeMax) | fresh;
            }
            while (high > kHalf && low <= kHalf && high <= kThreeQuarters && low > kQuarter) {
                if (bac.remaining()) in = bac.take(1u);
                high = (((high - (kQuarter + 1u)) << 1) & kRangeMax) | 1u;
                low = ((low - (kQuarter + 1u)) << 1) & kRangeMax;
                code = (((code - (kQuarter + 1u)) << 1) & kRangeMax) | in;
            }
            if (!bit) break;
            a++;
            if (bin == L - 1u) break;
        }
        if (a == L) {   // read_eg0 (LosslessCoder.cpp:113-165), uint16_t arithmetic
            uint32_t n = 0, x = 0;
            for (;;) {
                if (!byp.remaining()) return kErrResource;
                if (!byp.take(1u)) break;
                n = (n + 1u) & 0xFFu;
            }
            for (uint32_t k = 0; k < n; k++) {
                if (!byp.remaining()) return kErrResource;
                x = ((x << 1) | byp.take(1u)) & 0xFFFFu;
            }
            x = (x + ((1u << (n & 31u)) - 1u)) & 0xFFFFu;
            a = (a + x) & 0xFFFFu;
        }
        int v = (int)(int16_t)(uint16_t)a;
        if (v != 0) {   // read_sign (LosslessCoder.cpp:39-56)
            if (!byp.remaining()) return kErrResource;
            if (!byp.take(1u)) v = -v;
        }
        dst[i] = (int16_t)v;
    }
    return 0u;
}

}  // namespace core
}  // namespace eae
