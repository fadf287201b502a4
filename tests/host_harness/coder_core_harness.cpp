// Host build of the product's per-stream coder core (csrc/coder_core.cuh compiled with -DEAE_HOST_ONLY):
// lets the CPU test-suite check the restructured control flow (batched E1/E2 rescaling, word-wise bit
// I/O) against the oracle without a GPU. Test infrastructure; not part of libeae_b200.so.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EAE_HOST_ONLY 1
#include "../../autoencoder_based_image_compression_b200/csrc/coder_core.cuh"

using namespace eae::core;

extern "C" int harness_encode(uint32_t size, const int16_t* in, uint32_t L, const double* probs,
                              uint32_t cap_bits, uint8_t* bac_out, uint32_t* bac_bits, uint8_t* byp_out,
                              uint32_t* byp_bits)
{
    const size_t slot = ((cap_bits / 8 + 15) / 16) * 16 + 16;
    uint8_t* a = (uint8_t*)aligned_alloc(16, slot);
    uint8_t* b = (uint8_t*)aligned_alloc(16, slot);
    memset(a, 0xAA, slot);   // the sink must not depend on a zeroed slot
    memset(b, 0xAA, slot);
    BitSink bac, byp;
    bac.init(a, cap_bits);
    byp.init(b, cap_bits);
    const uint32_t e = encode_stream(in, size, probs, L, bac, byp);
    bac.flush();
    byp.flush();
    *bac_bits = bac.nbits;
    *byp_bits = byp.nbits;
    memcpy(bac_out, a, (bac.nbits + 7) / 8);
    memcpy(byp_out, b, (byp.nbits + 7) / 8);
    free(a);
    free(b);
    return (int)e;
}

// `misalign` (0..3) places the streams at an unaligned address, as inside a packed container.
extern "C" int harness_decode(uint32_t size, int16_t* out, uint32_t L, const double* probs,
                              const uint8_t* bac_in, uint32_t bac_bits, const uint8_t* byp_in,
                              uint32_t byp_bits, uint32_t misalign)
{
    const size_t nb = (bac_bits + 7) / 8, nr = (byp_bits + 7) / 8;
    uint8_t* a = (uint8_t*)aligned_alloc(16, nb + 32);
    uint8_t* b = (uint8_t*)aligned_alloc(16, nr + 32);
    memset(a, 0x55, nb + 32);
    memset(b, 0x55, nr + 32);
    memcpy(a + misalign, bac_in, nb);
    memcpy(b + misalign, byp_in, nr);
    BitSource bac, byp;
    bac.init(a + misalign, bac_bits);
    byp.init(b + misalign, byp_bits);
    const uint32_t e = decode_stream(out, size, probs, L, bac, byp);
    free(a);
    free(b);
    return (int)e;
}

// ---- lean formulation (coder v2) ----------------------------------------------------------------
// Literal E3 loop (BinaryArithmeticCoder.cpp:238-246) against e3_steps() and the closed-form register
// update, for every pair low < 0x8000 <= high. Returns the number of mismatches.
extern "C" uint64_t harness_e3_exhaustive(void)
{
    uint64_t bad = 0;
    for (uint32_t low0 = 0; low0 < 0x8000u; low0++) {
        for (uint32_t high0 = 0x8000u; high0 <= 0xFFFFu; high0++) {
            uint32_t low = low0, high = high0, k = 0;
            while (low > kQuarter && high <= kThreeQuarters) {
                low = (low - (kQuarter + 1u)) << 1;
                high = ((high - (kQuarter + 1u)) << 1) | 1u;
                k++;
            }
            const uint32_t kk = e3_steps(low0, high0);
            const uint32_t l2 = (low0 << kk) & kHalf;
            const uint32_t h2 = ((high0 << kk) & kHalf) | kMsb | ((1u << kk) - 1u);
            if (kk != k || l2 != low || h2 != high) bad++;
        }
    }
    return bad;
}

// Host emulation of the two-pass encoder: pass 1 writes the truncated-unary bit string and the bypass
// stream (on the GPU: binarize_streams_kernel), pass 2 runs one lean step per bin.
extern "C" int harness_encode2(uint32_t size, const int16_t* in, uint32_t L, const double* probs,
                               uint32_t cap_bits, uint8_t* bac_out, uint32_t* bac_bits, uint8_t* byp_out,
                               uint32_t* byp_bits)
{
    const size_t slot = ((cap_bits / 8 + 15) / 16) * 16 + 16;
    uint8_t* a = (uint8_t*)aligned_alloc(16, slot);
    uint8_t* b = (uint8_t*)aligned_alloc(16, slot);
    memset(a, 0xAA, slot);
    memset(b, 0xAA, slot);
    BitSink bac, byp;
    bac.init(a, cap_bits);
    byp.init(b, cap_bits);
    const size_t uwords = ((size_t)size * (L + 1u) + 31) / 32 + 1;
    uint32_t* unary = (uint32_t*)calloc(uwords, 4);
    uint32_t nbins = 0;
    uint32_t e = 0;
    for (uint32_t i = 0; i < size && !e; i++) {
        const int v = in[i];
        const uint32_t mag = (uint32_t)(v < 0 ? -v : v);
        const uint32_t ones = mag < L ? mag : L;
        for (uint32_t j = 0; j < ones; j++, nbins++) unary[nbins >> 5] |= 1u << (nbins & 31);
        if (mag < L) nbins++;
        uint32_t code, cnt;
        bypass_code(v, mag, L, code, cnt);
        if (cnt && !byp.put(code, cnt)) e = kErrCapacity;
    }
    BacState st = {0u, kRangeMax, 0u};
    uint32_t k = 0;
    for (uint32_t g = 0; g < nbins && !e; g++) {
        const uint32_t bit = (unary[g >> 5] >> (g & 31)) & 1u;
        e = lean_encode_bin(st, bac, bit, probs[k]);
        k = (bit && k + 1u < L) ? k + 1u : 0u;
    }
    if (!e) e = bac_finish(st, bac);
    bac.flush();
    byp.flush();
    *bac_bits = bac.nbits;
    *byp_bits = byp.nbits;
    memcpy(bac_out, a, (bac.nbits + 7) / 8);
    memcpy(byp_out, b, (byp.nbits + 7) / 8);
    free(a);
    free(b);
    free(unary);
    return (int)e;
}

// Host emulation of the two-phase decoder (on the GPU: decode_streams2_kernel): arithmetic decoding of all
// prefixes, then the bypass pass.
extern "C" int harness_decode2(uint32_t size, int16_t* out, uint32_t L, const double* probs,
                               const uint8_t* bac_in, uint32_t bac_bits, const uint8_t* byp_in,
                               uint32_t byp_bits, uint32_t misalign)
{
    const size_t nb = (bac_bits + 7) / 8, nr = (byp_bits + 7) / 8;
    uint8_t* a = (uint8_t*)aligned_alloc(16, nb + 32);
    uint8_t* b = (uint8_t*)aligned_alloc(16, nr + 32);
    memset(a, 0x55, nb + 32);
    memset(b, 0x55, nr + 32);
    memcpy(a + misalign, bac_in, nb);
    memcpy(b + misalign, byp_in, nr);
    BitSource bac, byp;
    bac.init(a + misalign, bac_bits);
    byp.init(b + misalign, byp_bits);
    DecState st;
    lean_decode_start(st, bac);
    uint32_t e = 0, n_ok = size, i = 0, mag = 0, k = 0;
    while (i < size) {
        const double p = probs[k];
        if (!(p > 0.0 && p < 1.0)) { e = kErrProbability; n_ok = i; break; }
        const uint32_t bit = lean_decode_bin(st, bac, p);
        mag += bit;
        if (!bit || k == L - 1u) { out[i++] = (int16_t)mag; mag = 0; k = 0; } else k++;
    }
    for (i = 0; i < n_ok; i++) {
        int v;
        const uint32_t eb = lean_decode_bypass((uint32_t)(uint16_t)out[i], L, byp, v);
        if (eb) { e = eb; break; }
        out[i] = (int16_t)v;
    }
    free(a);
    free(b);
    return (int)e;
}

// ---- fast formulation (coder v3) ----------------------------------------------------------------
// fast_rescale() against the reference's literal rescaling loop (BinaryArithmeticCoder.cpp:182-252) for
// every register pair low <= high after an interval update. Returns the number of mismatches.
extern "C" uint64_t harness_rescale_exhaustive(void)
{
    uint64_t bad = 0;
    #pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad)
    for (int64_t lo = 0; lo <= 0xFFFF; lo++) {
        for (uint32_t hi = (uint32_t)lo; hi <= 0xFFFFu; hi++) {
            uint32_t low = (uint32_t)lo, high = hi, n = 0, k = 0;
            for (;;) {
                if (((low ^ high) & kMsb) == 0u) {
                    low = (low << 1) & kRangeMax;
                    high = ((high << 1) & kRangeMax) | 1u;
                    n++;
                } else if (low > kQuarter && high <= kThreeQuarters) {
                    low = (low - (kQuarter + 1u)) << 1;
                    high = ((high - (kQuarter + 1u)) << 1) | 1u;
                    k++;
                } else {
                    break;
                }
            }
            uint32_t l2 = 0, h2 = 0, k2 = 0;
            const uint32_t n2 = fast_rescale(l2, h2, (uint32_t)lo, hi, k2);
            if (n2 != n || k2 != k || l2 != low || h2 != high) bad++;
        }
    }
    return bad;
}

// 1 when the 48-bit fixed-point multiplier reproduces floor(p * range) for every range.
static int fixed48_matches(double p)
{
    const MulFp64 a{p};
    const MulFixed48 b{fixed48_of(p)};
    for (uint32_t r = 0; r <= 0xFFFFu; r++) if (a(r) != b(r)) return 0;
    return 1;
}

// Returns the largest number of bits one bin released (the fast sink takes at most 31 at a time).
template <typename Mul>
static uint32_t encode3_loop(const uint32_t* unary, uint32_t nbins, const Mul* muls, uint32_t L, BacState& st, FastSink& bac)
{
    uint32_t k = 0, ev = 0, ec = 0, worst = 0;
    for (uint32_t g = 0; g < nbins; g++) {
        bac.put(ev, ec);      // the previous bin's bits, as in the kernel's rotated loop
        const uint32_t bit = (unary[g >> 5] >> (g & 31)) & 1u;
        fast_encode_arith(st, bit, muls[k], ev, ec);
        if (ec > worst) worst = ec;
        if (ec > 31u) return worst;
        k = (bit && k + 1u < L) ? k + 1u : 0u;
    }
    bac.put(ev, ec);
    return worst;
}

// mode 0: FP64 multiplier, mode 1: fixed-point multiplier (when the table validates, as in the product).
extern "C" int harness_encode3(uint32_t size, const int16_t* in, uint32_t L, const double* probs,
                               uint32_t cap_bits, uint8_t* bac_out, uint32_t* bac_bits, uint8_t* byp_out,
                               uint32_t* byp_bits, int mode)
{
    bool row_ok = true;
    for (uint32_t j = 0; j < L; j++) row_ok = row_ok && probs[j] > 0.0 && probs[j] < 1.0;
    if (!row_ok) return harness_encode2(size, in, L, probs, cap_bits, bac_out, bac_bits, byp_out, byp_bits);
    const size_t slot = ((cap_bits / 8 + 15) / 16) * 16 + 16;
    uint8_t* a = (uint8_t*)aligned_alloc(16, slot);
    uint8_t* b = (uint8_t*)aligned_alloc(16, slot);
    memset(a, 0xAA, slot);
    memset(b, 0xAA, slot);
    FastSink bac;
    BitSink byp;
    bac.init(a, cap_bits);
    byp.init(b, cap_bits);
    const size_t uwords = ((size_t)size * L + 31) / 32 + 1;
    uint32_t* unary = (uint32_t*)calloc(uwords, 4);
    uint32_t nbins = 0;
    for (uint32_t i = 0; i < size; i++) {
        const int v = in[i];
        const uint32_t mag = (uint32_t)(v < 0 ? -v : v);
        const uint32_t ones = mag < L ? mag : L;
        for (uint32_t j = 0; j < ones; j++, nbins++) unary[nbins >> 5] |= 1u << (nbins & 31);
        if (mag < L) nbins++;
        uint32_t code, cnt;
        bypass_code(v, mag, L, code, cnt);
        if (cnt) byp.put(code, cnt);
    }
    bool fixed = mode == 1;
    for (uint32_t j = 0; j < L && fixed; j++) fixed = fixed48_matches(probs[j]);
    BacState st = {0u, kRangeMax, 0u};
    uint32_t worst;
    if (fixed) {
        MulFixed48* m = (MulFixed48*)malloc(sizeof(MulFixed48) * L);
        for (uint32_t j = 0; j < L; j++) m[j].q = fixed48_of(probs[j]);
        worst = encode3_loop(unary, nbins, m, L, st, bac);
        free(m);
    } else {
        MulFp64* m = (MulFp64*)malloc(sizeof(MulFp64) * L);
        for (uint32_t j = 0; j < L; j++) m[j].p = probs[j];
        worst = encode3_loop(unary, nbins, m, L, st, bac);
        free(m);
    }
    if (worst > 31u) {     // more than 31 bits from one bin: the kernels redo the stream with the lean loop
        free(a);
        free(b);
        free(unary);
        return harness_encode2(size, in, L, probs, cap_bits, bac_out, bac_bits, byp_out, byp_bits);
    }
    fast_finish(st, bac);
    bac.flush();
    byp.flush();
    const uint32_t e = bac.pos() > cap_bits ? kErrCapacity : 0u;
    *bac_bits = bac.pos();
    *byp_bits = byp.nbits;
    if (!e) {
        memcpy(bac_out, a, (bac.pos() + 7) / 8);
        memcpy(byp_out, b, (byp.nbits + 7) / 8);
    }
    free(a);
    free(b);
    free(unary);
    return (int)e;
}

extern "C" int harness_decode3(uint32_t size, int16_t* out, uint32_t L, const double* probs,
                               const uint8_t* bac_in, uint32_t bac_bits, const uint8_t* byp_in,
                               uint32_t byp_bits, uint32_t misalign, int mode)
{
    bool row_ok = true;
    for (uint32_t j = 0; j < L; j++) row_ok = row_ok && probs[j] > 0.0 && probs[j] < 1.0;
    if (!row_ok) return harness_decode2(size, out, L, probs, bac_in, bac_bits, byp_in, byp_bits, misalign);
    const size_t nb = (bac_bits + 7) / 8, nr = (byp_bits + 7) / 8;
    uint8_t* a = (uint8_t*)aligned_alloc(16, nb + 32);
    uint8_t* b = (uint8_t*)aligned_alloc(16, nr + 32);
    memset(a, 0x55, nb + 32);
    memset(b, 0x55, nr + 32);
    memcpy(a + misalign, bac_in, nb);
    memcpy(b + misalign, byp_in, nr);
    FastSource bac;
    BitSource byp;
    static const uint32_t empty_word = 0;
    bac.init(a + misalign, bac_bits, &empty_word);
    byp.init(b + misalign, byp_bits);
    bool fixed = mode == 1;
    for (uint32_t j = 0; j < L && fixed; j++) fixed = fixed48_matches(probs[j]);
    DecState st;
    fast_decode_start(st, bac);
    uint32_t e = 0, i = 0, mag = 0, k = 0;
    while (i < size) {
        uint32_t bit;
        const bool steady = bac.left >= 32u;     // the kernel's unchecked loop runs while this holds
        if (fixed) {
            const MulFixed48 m{fixed48_of(probs[k])};
            bit = steady ? fast_decode_bin<false>(st, bac, m) : fast_decode_bin<true>(st, bac, m);
        } else {
            const MulFp64 m{probs[k]};
            bit = steady ? fast_decode_bin<false>(st, bac, m) : fast_decode_bin<true>(st, bac, m);
        }
        mag += bit;
        if (!bit || k == L - 1u) { out[i++] = (int16_t)mag; mag = 0; k = 0; } else k++;
    }
    for (i = 0; i < size; i++) {
        int v;
        const uint32_t eb = lean_decode_bypass((uint32_t)(uint16_t)out[i], L, byp, v);
        if (eb) { e = eb; break; }
        out[i] = (int16_t)v;
    }
    free(a);
    free(b);
    return (int)e;
}
