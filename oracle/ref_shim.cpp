// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// C-ABI shim around the UNMODIFIED reference lossless coder, compiled from the sources where
// they lie under /root/reference/kodak_tensorflow/lossless/c++/source (see oracle/Makefile).
// Nothing from the reference is copied here: this file only #includes the reference headers
// and pokes at the private bit buffers (the Makefile passes -fno-access-control) so that tests
// can obtain golden BYTES, which the reference never exposes (compression.cpp:29-64 encodes and
// decodes inside one call).
//
// Output: oracle/_ref/libref_coder.so (git-ignored, travels to the GPU box).

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include "compression.h"  // reference: compress_lossless (compression.h:41-45)

namespace {

// "Error of type N ..." -> N; anything else -> fallback.
int parse_error_code(const char* what, int fallback)
{
    const char* key = "Error of type ";
    const char* at = std::strstr(what, key);
    if (!at) return fallback;
    return (int)std::strtol(at + std::strlen(key), nullptr, 10);
}

void copy_stream(const Bitstream& bs, uint8_t* out, uint32_t cap_bytes, uint32_t* nb_bits)
{
    *nb_bits = bs.m_write_index;
    const uint32_t nb_bytes = (bs.m_write_index + 7) >> 3;
    if (out) std::memcpy(out, bs.m_data.data(), nb_bytes < cap_bytes ? nb_bytes : cap_bytes);
}

void poke_stream(Bitstream& bs, const uint8_t* in, uint32_t nb_bits)
{
    const uint32_t nb_bytes = (nb_bits + 7) >> 3;
    if (bs.m_data.size() < nb_bytes) bs.m_data.resize(nb_bytes);
    std::memcpy(bs.m_data.data(), in, nb_bytes);
    bs.m_write_index = nb_bits;
    bs.m_read_index = 0;
}

}  // namespace

extern "C" {

// The reference entry point itself. Returns 0 on success, the reference error_code (1..4) when a
// std::runtime_error carries one, -1 for std::invalid_argument, -2 for std::out_of_range.
int ref_compress_lossless(uint32_t size, const int16_t* in, int16_t* out, uint8_t L,
                          const double* probs, uint32_t* nb_bits)
{
    try {
        *nb_bits = compress_lossless(size, in, out, L, probs);
        return 0;
    } catch (const std::invalid_argument&) {
        return -1;
    } catch (const std::out_of_range&) {
        return -2;
    } catch (const std::runtime_error& e) {
        return parse_error_code(e.what(), -3);
    }
}

// Encode side of compress_lossless (compression.cpp:24-49), additionally dumping both buffers.
int ref_encode_map(uint32_t size, const int16_t* in, uint8_t L, const double* probs,
                   uint8_t* bac_out, uint32_t bac_cap_bytes, uint32_t* bac_bits,
                   uint8_t* byp_out, uint32_t byp_cap_bytes, uint32_t* byp_bits)
{
    try {
        uint32_t required(size * std::max((uint32_t)32, (uint32_t)L));
        LosslessCoder coder(required, L, probs);
        error_code state(success);
        for (uint32_t i(0); i < size; i++) {
            state = coder.write_signed_ueg0(in[i]);
            if (state) return (int)state;
        }
        state = coder.stop_bac_encoding();
        if (state) return (int)state;
        copy_stream(coder.m_bac.m_bitstream, bac_out, bac_cap_bytes, bac_bits);
        copy_stream(coder.m_bitstream_bypass, byp_out, byp_cap_bytes, byp_bits);
        return 0;
    } catch (const std::out_of_range&) {
        return -2;
    }
}

// Decode side of compress_lossless (compression.cpp:51-63) fed with EXTERNAL bytes.
int ref_decode_map(uint32_t size, int16_t* out, uint8_t L, const double* probs,
                   const uint8_t* bac_in, uint32_t bac_bits,
                   const uint8_t* byp_in, uint32_t byp_bits)
{
    try {
        uint32_t required(size * std::max((uint32_t)32, (uint32_t)L));
        LosslessCoder coder(required, L, probs);
        poke_stream(coder.m_bac.m_bitstream, bac_in, bac_bits);
        poke_stream(coder.m_bitstream_bypass, byp_in, byp_bits);
        error_code state = coder.start_bac_decoding();
        if (state) return (int)state;
        for (uint32_t i(0); i < size; i++) {
            state = coder.read_signed_ueg0(out[i]);
            if (state) return (int)state;
        }
        return 0;
    } catch (const std::out_of_range&) {
        return -2;
    }
}

// Raw binary arithmetic coder: one probability per bit (tests.cpp:69-132).
int ref_bac_encode_bits(uint32_t n, const uint8_t* bits, const double* probs,
                        uint8_t* out, uint32_t cap_bytes, uint32_t* nb_bits)
{
    BinaryArithmeticCoder bac(cap_bytes * 8);
    error_code state(success);
    for (uint32_t i(0); i < n; i++) {
        state = bac.encoding(bits[i], probs[i]);
        if (state) return (int)state;
    }
    state = bac.stop_encoding();
    if (state) return (int)state;
    copy_stream(bac.m_bitstream, out, cap_bytes, nb_bits);
    return 0;
}

int ref_bac_decode_bits(uint32_t n, uint8_t* bits, const double* probs,
                        const uint8_t* in, uint32_t nb_bits)
{
    BinaryArithmeticCoder bac(nb_bits + 64);
    poke_stream(bac.m_bitstream, in, nb_bits);
    error_code state = bac.start_decoding();
    if (state) return (int)state;
    uint8_t storage(0);
    for (uint32_t i(0); i < n; i++) {
        state = bac.decoding(storage, probs[i]);
        if (state) return (int)state;
        bits[i] = storage;
    }
    return 0;
}

uint32_t ref_create_divisible(uint32_t input, uint32_t divisor) { return create_divisible(input, divisor); }
uint32_t ref_count_nb_bits(uint32_t input) { return count_nb_bits(input); }

}  // extern "C"
