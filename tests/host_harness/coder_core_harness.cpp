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
