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
    // Next n (1..31) bits, first bit at bit 0. Caller guarantees n <= remaining().
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


// =================================================================================================
// Lean formulation (coder v2). Same bitstreams, restructured so that the strictly sequential part of a
// stream is ONE short, branch-light step per binary decision:
//  * the bypass stream (EG0 suffixes + signs) does not depend on the arithmetic coder at all, so the
//    encoder writes it in a separate, fully parallel pass and the decoder reads it after the arithmetic
//    decoding of the prefixes;
//  * the truncated-unary prefixes of a stream are just a bit string (a ones then a zero if a < L): the
//    parallel pass of the encoder writes that string, the sequential pass consumes one bit per step;
//  * E1/E2 (n steps) and E3 (k steps) are both applied in closed form, so a step has no loop.
//
// E3 in closed form. After E1/E2 the registers satisfy low < 0x8000 <= high. One E3 step
// (BinaryArithmeticCoder.cpp:238-246) needs low > 0x3FFF (bit 14 of low set) and high <= 0xBFFD (bit 14 of
// high clear, and high not in {0xBFFE, 0xBFFF}: the reference's 3 * 0x3FFF quirk) and maps
// low -> (low - 0x4000) << 1, high -> ((high - 0x4000) << 1) | 1. Hence step j (0-based) looks at bit 14 - j
// of the original registers: the bit tests allow kl = (number of ones of low from bit 14 down) and
// kh = (number of zeros of high from bit 14 down) steps. The quirk blocks step j iff the current high is
// 0xBFFE / 0xBFFF, i.e. iff high = 0x8000 | (2^m - 1) with m = 14 - j (then kh = j + 1: it blocks exactly the
// last step the bit test would allow) or high = 0xBFFE (j = 0). After k steps
// low = (low << k) & 0x7FFF, high = ((high << k) & 0x7FFF) | 0x8000 | (2^k - 1).
// tests/host_harness checks this against the literal loop for every (low, high) pair.
EAE_HD uint32_t e3_steps(uint32_t low, uint32_t high)
{
    const uint32_t kl = clz32(~(low << 17));                 // low < 0x8000: at most 15
    const uint32_t t = high & kHalf;
    uint32_t kh = t ? clz32(t << 17) : 15u;
    if (kh && ((t & (t + 1u)) == 0u || t == 0x3FFEu)) kh--;
    return kl < kh ? kl : kh;
}

// The low n (0..32) bits of x in reversed order.
EAE_HD uint32_t rev_n(uint32_t x, uint32_t n) { return (uint32_t)((uint64_t)brev32(x) >> (32u - n)); }

// One binary decision of the encoder: encode_bit + rescale_encoding (BinaryArithmeticCoder.cpp:49-59,
// 158-252). Returns 0 or an error code.
EAE_HD uint32_t lean_encode_bin(BacState& s, BitSink& out, uint32_t bit, double p)
{
    if (!(p > 0.0 && p < 1.0)) return kErrProbability;
    const uint32_t mid = s.low + mul_floor(p, s.high - s.low);
    uint32_t low = bit ? mid + 1u : s.low;
    uint32_t high = bit ? s.high : mid;
    if ((low | high) > kRangeMax) return kErrPrecision;
    const uint32_t n = clz32(((low ^ high) << 16) | 0x8000u);    // leading equal bits: E1/E2 steps
    const uint32_t msb_first = brev32(low << 16);                // bit 0 = MSB of low, in emission order
    low = (low << n) & kRangeMax;
    high = ((high << n) & kRangeMax) | ((1u << n) - 1u);
    const uint32_t k = e3_steps(low, high);
    s.low = (low << k) & kHalf;
    s.high = ((high << k) & kHalf) | kMsb | ((1u << k) - 1u);
    if (n) {
        const uint32_t bits = msb_first & ((1u << n) - 1u);
        const uint32_t q = s.pending;
        if (n + q <= 32u) {
            // first bit, q queued E3 bits (inverted, :317-337), the other n - 1 bits
            const uint32_t first = bits & 1u;
            const uint32_t run = first ? 0u : (uint32_t)((1ull << q) - 1ull);
            const uint32_t v = first | (run << 1) | (uint32_t)((uint64_t)(bits >> 1) << (1u + q));
            if (!out.put(v, n + q)) return kErrCapacity;
        } else {
            const uint32_t first = bits & 1u;
            if (!out.put(first, 1u)) return kErrCapacity;
            if (!out.put_run(first ^ 1u, q)) return kErrCapacity;
            if (n > 1u && !out.put(bits >> 1, n - 1u)) return kErrCapacity;
        }
        s.pending = 0u;
    }
    s.pending += k;
    return 0u;
}

// Decoder registers (BinaryArithmeticCoder.h:12-16 plus the code word).
struct DecState {
    uint32_t low, high, code;
};

// start_decoding (BinaryArithmeticCoder.cpp:104-122): 16 bits, padded with the last bit read.
EAE_HD void lean_decode_start(DecState& s, BitSource& bac)
{
    s.low = 0u; s.high = kRangeMax; s.code = 0u;
    const uint32_t have = bac.remaining() < 16u ? bac.remaining() : 16u;
    uint32_t keep = 0u;
    if (have) {
        const uint32_t bits = bac.take(have);
        s.code = rev_n(bits, have);
        keep = (bits >> (have - 1u)) & 1u;
    }
    for (uint32_t j = have; j < 16u; j++) s.code = (s.code << 1) | keep;
}

// One binary decision of the decoder: decode_bit + rescale_decoding (BinaryArithmeticCoder.cpp:254-315).
// low <= code <= high is an invariant of these updates whatever bits are shifted in (both rescalings map the
// three registers by the same affine map), so the reference's two range tests reduce to code > mid. The
// n + k new bits are read in one go; past the end of the stream the reference repeats the last bit read
// in the same rescaling call (or 0), reproduced in the slow branch. p must be valid (caller checks).
EAE_HD uint32_t lean_decode_bin(DecState& s, BitSource& bac, double p)
{
    const uint32_t mid = s.low + mul_floor(p, s.high - s.low);
    const uint32_t bit = s.code > mid ? 1u : 0u;
    uint32_t low = bit ? mid + 1u : s.low;
    uint32_t high = bit ? s.high : mid;
    const uint32_t n = clz32(((low ^ high) << 16) | 0x8000u);
    low = (low << n) & kRangeMax;
    high = ((high << n) & kRangeMax) | ((1u << n) - 1u);
    const uint32_t k = e3_steps(low, high);
    s.low = (low << k) & kHalf;
    s.high = ((high << k) & kHalf) | kMsb | ((1u << k) - 1u);
    const uint32_t total = n + k;                                // at most 30
    if (total) {
        uint32_t fresh;
        if (bac.remaining() >= total) {
            fresh = rev_n(bac.take(total), total);
        } else {
            const uint32_t have = bac.remaining();
            uint32_t in = 0u;
            fresh = 0u;
            if (have) {
                const uint32_t bits = bac.take(have);
                fresh = rev_n(bits, have);
                in = (bits >> (have - 1u)) & 1u;
            }
            for (uint32_t j = have; j < total; j++) fresh = (fresh << 1) | in;
        }
        s.code = (((s.code << total) | fresh) & kRangeMax) ^ (k ? kMsb : 0u);
    }
    return bit;
}

// Bypass part of read_signed_ueg0 for a symbol whose truncated-unary prefix decoded to `a` (0..L):
// read_eg0 (LosslessCoder.cpp:113-165, uint16_t arithmetic) when a == L, read_sign (:39-56) when the value is
// not 0. Returns 0 or kErrResource.
EAE_HD uint32_t lean_decode_bypass(uint32_t a, uint32_t L, BitSource& byp, int& v_out)
{
    if (a == L) {
        uint32_t n = 0, x = 0;
        for (;;) {
            if (!byp.remaining()) return kErrResource;
            if (!byp.take(1u)) break;
            n = (n + 1u) & 0xFFu;
        }
        for (uint32_t j = 0; j < n; j++) {
            if (!byp.remaining()) return kErrResource;
            x = ((x << 1) | byp.take(1u)) & 0xFFFFu;
        }
        x = (x + ((1u << (n & 31u)) - 1u)) & 0xFFFFu;
        a = (a + x) & 0xFFFFu;
    }
    int v = (int)(int16_t)(uint16_t)a;
    if (v != 0) {
        if (!byp.remaining()) return kErrResource;
        if (!byp.take(1u)) v = -v;
    }
    v_out = v;
    return 0u;
}


// =================================================================================================
// Fast formulation (coder v3): the lean step with no data-dependent branch on its common path.
//  * E1/E2 and E3 are one shift by s = n + k: with x = low ^ high and z = low & ~high (16-bit registers
//    after the interval update), n = leading zeros of x, and the E3 steps are the run of positions below
//    bit 15 - n where low has a one and high a zero, i.e. the leading ones of z << (n + 1). Then
//    low = (low << s) & 0x7FFF, high = ((high << s | 2^s - 1) & 0x7FFF) | 0x8000. The 0xBFFD quirk
//    (see e3_steps) blocked the last E3 step exactly when k > 0 and that formula gives high >= 0xFFFD;
//    one rarely taken branch redoes the shift with k - 1. Checked against the literal loops for every
//    register pair by tests/host_harness.
//  * Emission is MSB-first: the n renormalisation bits E (top bits of low) followed by the q queued
//    follow bits are the (n + q)-bit number E + ((2^q - 1) << (n - 1)) (a 0 followed by q ones plus the
//    carry of the decided bit). Bits are collected in the top of a 64-bit register and leave as whole,
//    bit-reversed 32-bit words (the stream is LSB-first per byte, Bitstream.cpp:36-58).
//  * Capacity is not tested per bit: stores beyond the slot are redirected to a spare word and the bit count
//    is compared with the capacity afterwards (same verdict: the reference fails as soon as one bit does not fit).
//  * Probabilities are validated per table row before the loop, not per use; rows with an invalid entry
//    take the lean path, which reports errors in the reference's order.

// x << s and x >> s for s in 0..63 with the PTX meaning (0 once s >= 32).
EAE_HD uint32_t shl_sat(uint32_t x, uint32_t s)
{
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
#else
    return s < 32u ? x << s : 0u;
#endif
}
EAE_HD uint32_t shr_sat(uint32_t x, uint32_t s)
{
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
#else
    return s < 32u ? x >> s : 0u;
#endif
}

// floor(p * range) policies: the reference's FP64 arithmetic, or a 48-bit fixed-point multiplier that the
// table validation kernel has proven to give the same result for every range (0..65535).
struct MulFp64 {
    double p;
    EAE_HD uint32_t operator()(uint32_t range) const { return mul_floor(p, range); }
};
struct MulFixed48 {
    uint64_t q;     // floor(p * 2^48)
    EAE_HD uint32_t operator()(uint32_t range) const { return (uint32_t)((q * (uint64_t)range) >> 48); }
};
EAE_HD uint64_t fixed48_of(double p) { return (p > 0.0 && p < 1.0) ? (uint64_t)(p * 281474976710656.0) : 0ull; }

// Interval update + combined rescaling. Returns n (E1/E2 steps); k (E3 steps) and the pre-shift low register
// are returned through the references.
EAE_HD uint32_t fast_rescale(uint32_t& low, uint32_t& high, uint32_t low1, uint32_t high1, uint32_t& k_out)
{
    const uint32_t x = low1 ^ high1;
    const uint32_t n = clz32((x << 16) | 0x8000u);
    const uint32_t z = low1 & ~high1;
    uint32_t k = clz32(~(shl_sat(z << n, 17u)));
    uint32_t s = n + k;
    uint32_t low2 = (low1 << s) & kHalf;
    uint32_t high2 = (((high1 << s) | ((1u << s) - 1u)) & kHalf) | kMsb;
    if (k && high2 >= 0xFFFDu) {
        k--; s--;
        low2 = (low1 << s) & kHalf;
        high2 = (((high1 << s) | ((1u << s) - 1u)) & kHalf) | kMsb;
    }
    low = low2; high = high2; k_out = k;
    return n;
}

struct FastSink {
    uint32_t cur;          // the word being filled, first bit at bit 31
    uint32_t fill;         // bits in cur, < 32 between calls
    uint32_t widx;         // whole words stored so far
    uint32_t spare;        // index of the word that absorbs stores past the capacity
    uint32_t* words;

    EAE_HD void init(uint8_t* slot, uint32_t cap_bits)
    {
        words = reinterpret_cast<uint32_t*>(slot);
        cur = 0; fill = 0; widx = 0;
        spare = (cap_bits + 31u) >> 5;      // the slot is at least 16 bytes longer than the capacity
    }
    EAE_HD uint32_t pos() const { return widx * 32u + fill; }
    // Appends the c (0..31) low bits of v (v < 2^c), most significant first. No branch: the store of a
    // completed word is the only conditional instruction.
    EAE_HD void put(uint32_t v, uint32_t c)
    {
        const uint32_t top = shl_sat(v, 32u - c);
        cur |= top >> fill;
        const uint32_t spill = shl_sat(top, 32u - fill);
        fill += c;
        const bool full = fill >= 32u;
        if (full) words[widx < spare ? widx : spare] = brev32(cur);
        widx += full ? 1u : 0u;
        cur = full ? spill : cur;
        fill -= full ? 32u : 0u;
    }
    EAE_HD void put_run(uint32_t bit, uint32_t repeat)
    {
        while (repeat) {
            const uint32_t c = repeat < 31u ? repeat : 31u;
            put(bit ? ((1u << c) - 1u) : 0u, c);
            repeat -= c;
        }
    }
    EAE_HD void flush() { if (fill) words[widx < spare ? widx : spare] = brev32(cur); }
};

// Arithmetic of one encoder bin. The bits it releases are returned as the `ec`-bit number `ev` (most
// significant bit first) for FastSink::put, which the kernels call one iteration later so that the two
// dependent chains overlap. ec can exceed 31 only after more than 15 consecutive E3 steps; callers then
// redo the stream with the lean formulation.
template <typename Mul>
EAE_HD void fast_encode_arith(BacState& s, uint32_t bit, const Mul& mul, uint32_t& ev, uint32_t& ec)
{
    const uint32_t mid = s.low + mul(s.high - s.low);
    const uint32_t low1 = bit ? mid + 1u : s.low;
    const uint32_t high1 = bit ? s.high : mid;
    uint32_t k;
    const uint32_t n = fast_rescale(s.low, s.high, low1, high1, k);
    const uint32_t q = n ? s.pending : 0u;         // follow bits leave with the first renormalisation bit
    s.pending = (n ? 0u : s.pending) + k;
    // n bits of low (MSB first) + q follow bits = (0 then q ones, from the bit after the first) + carry
    ec = n + q;
    ev = (low1 >> (16u - n)) + shr_sat(shl_sat(shl_sat(1u, q) - 1u, n), 1u);
}

// stop_encoding (BinaryArithmeticCoder.cpp:61-102) on the fast sink.
EAE_HD void fast_finish(BacState& s, FastSink& out)
{
    const uint32_t bit = s.low < kQuarter ? 0u : 1u;
    out.put(bit, 1u);
    out.put_run(bit ^ 1u, s.pending + 1u);
}

// MSB-first window over an LSB-first packed stream at any byte alignment. `left` counts the real bits not yet
// consumed; the window may also hold bytes that follow the stream, which are never used: take() needs
// s <= left, take_padded() implements the reference's behaviour at the end of the stream. Word fetches are
// clamped to the aligned span that contains the stream.
struct FastSource {
    uint64_t win;          // next bit at bit 63
    uint32_t have;         // bits in win, > 32 between calls
    uint32_t left;
    uint32_t next;         // prefetched word
    uint32_t widx, last;   // next word to fetch, index of the last word of the span
    const uint32_t* words;

    // `empty` is any readable word; it stands in for a stream of zero bits.
    EAE_HD void init(const uint8_t* p, uint32_t bits, const uint32_t* empty)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t skip = bits ? (uint32_t)(a & 3u) * 8u : 0u;
        words = bits ? reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3) : empty;
        last = bits ? ((skip + bits + 31u) >> 5) - 1u : 0u;
        left = bits;
        win = (uint64_t)brev32(load_ro(words) >> skip) << 32;
        have = 32u - skip;
        widx = 1;
        next = load_ro(words + (widx < last ? widx : last));
        widx++;
        refill();
    }
    EAE_HD void refill()
    {
        const bool need = have <= 32u;
        const uint64_t add = (uint64_t)brev32(next) << ((32u - have) & 63u);
        win |= need ? add : 0ull;
        have += need ? 32u : 0u;
        if (need) next = load_ro(words + (widx < last ? widx : last));
        widx += need ? 1u : 0u;
    }
    // Next s (0..32) bits as a number, first bit most significant. Needs s <= left.
    EAE_HD uint32_t take(uint32_t s)
    {
        const uint32_t v = shr_sat((uint32_t)(win >> 32), 32u - s);
        win <<= s;
        have -= s;
        left -= s;
        refill();
        return v;
    }
    // The reference's behaviour at the end of the stream (BinaryArithmeticCoder.cpp:104-122, 275-315): the
    // bits that are there, then the last of them (or 0) repeated.
    EAE_HD uint32_t take_padded(uint32_t s)
    {
        const uint32_t r = left < s ? left : s;
        uint32_t v = 0u, in = 0u;
        if (r) { v = take(r); in = v & 1u; }
        for (uint32_t j = r; j < s; j++) v = (v << 1) | in;
        return v;
    }
};

EAE_HD void fast_decode_start(DecState& s, FastSource& bac)
{
    s.low = 0u; s.high = kRangeMax;
    s.code = bac.take_padded(16u);
}

// One decoder bin. kChecked = false is the steady state (the caller guarantees left >= 30, so the n + k new
// bits are there); kChecked = true also handles the end of the stream.
template <bool kChecked, typename Mul>
EAE_HD uint32_t fast_decode_bin(DecState& s, FastSource& bac, const Mul& mul)
{
    const uint32_t mid = s.low + mul(s.high - s.low);
    const uint32_t bit = s.code > mid ? 1u : 0u;
    const uint32_t low1 = bit ? mid + 1u : s.low;
    const uint32_t high1 = bit ? s.high : mid;
    uint32_t k;
    const uint32_t n = fast_rescale(s.low, s.high, low1, high1, k);
    const uint32_t sh = n + k;
    const uint32_t fresh = kChecked ? bac.take_padded(sh) : bac.take(sh);
    s.code = (((s.code << sh) | fresh) & kRangeMax) ^ (k ? kMsb : 0u);
    return bit;
}


}  // namespace core
}  // namespace eae
