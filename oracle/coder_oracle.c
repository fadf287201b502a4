/*
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's library; the product path never does.
 *
 * Plain-C restatement of the reference's lossless coder for one feature map:
 * signed UEG0 binarisation (truncated-unary prefix through a 16-bit binary arithmetic coder with
 * FIXED per-position probabilities, Exp-Golomb-0 suffix and sign as raw "bypass" bits).
 * Citations are into /root/reference/kodak_tensorflow/lossless/c++/source/.
 *
 * Pinned (tests/test_oracle_coder.py) against (i) the known answers in the reference's own tests
 * (tests.cpp:69-376, test_lossless.py:89-101) and (ii) byte-for-byte against the reference coder
 * compiled from its own sources (oracle/_ref/libref_coder.so) on randomized maps; the resulting
 * bytes are committed under tests/golden/.
 *
 * Build: make -C oracle   (-> oracle/liboracle_coder.so)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* utils.h:12-19 */
enum { EAE_OK = 0, EAE_CAPACITY = 1, EAE_RESOURCE = 2, EAE_PRECISION = 3, EAE_PROBABILITY = 4 };

/* BinaryArithmeticCoder.cpp:3-33. Note 3*0x3FFF = 0xBFFD (not 0xBFFF). */
#define R_MAX 0xFFFFu
#define R_HALF 0x7FFFu
#define R_QUARTER 0x3FFFu
#define R_3QUARTER (3u * R_QUARTER)
#define R_MSB 0x8000u

/* ---- bit buffer: bit i lives in byte i>>3 at position i&7 (Bitstream.cpp:36-79) ---- */
typedef struct {
    uint8_t* data;
    uint32_t cap_bits; /* multiple of 8 (Bitstream.cpp:3-7, utils.cpp:3-11) */
    uint32_t wr, rd;
} bits_t;

static uint32_t round_up(uint32_t x, uint32_t d) { uint32_t r = x % d; return r ? x + d - r : x; }

static int put(bits_t* b, unsigned bit)
{
    if (b->wr + 1 > b->cap_bits) return EAE_CAPACITY;
    uint8_t m = (uint8_t)(1u << (b->wr & 7));
    if (bit & 1) b->data[b->wr >> 3] |= m; else b->data[b->wr >> 3] &= (uint8_t)~m;
    b->wr++;
    return EAE_OK;
}

static int get(bits_t* b, unsigned* bit)
{
    if (b->rd >= b->wr) return EAE_RESOURCE;
    *bit = (b->data[b->rd >> 3] >> (b->rd & 7)) & 1u;
    b->rd++;
    return EAE_OK;
}

/* ---- arithmetic coder state (BinaryArithmeticCoder.h:11-16) ---- */
typedef struct { uint32_t low, high, pending, code; bits_t out; } bac_t;

static void bac_reset(bac_t* c) { c->low = 0; c->high = R_MAX; c->pending = 0; }

/* BinaryArithmeticCoder.cpp:144-156: split point, FP64 multiply then floor. */
static int split(const bac_t* c, double p, uint32_t* mid)
{
    if (isnan(p) || p <= 0.0 || p >= 1.0) return EAE_PROBABILITY;
    volatile double prod = p * (double)(c->high - c->low); /* volatile: no fused/extended evaluation */
    *mid = c->low + (uint32_t)floor(prod);
    return EAE_OK;
}

/* BinaryArithmeticCoder.cpp:317-337 */
static int flush_pending(bac_t* c, unsigned bit)
{
    for (; c->pending; c->pending--) { int e = put(&c->out, !(bit & 1)); if (e) return e; }
    return EAE_OK;
}

/* encoding() = encode_bit + rescale_encoding (BinaryArithmeticCoder.cpp:49-59, 158-252) */
static int bac_put(bac_t* c, unsigned bit, double p)
{
    uint32_t mid; int e = split(c, p, &mid);
    if (e) return e;
    if (bit & 1) c->low = mid + 1; else c->high = mid;
    if (c->high > R_MAX || c->low > R_MAX) return EAE_PRECISION;
    for (;;) {
        uint32_t top = c->high & R_MSB;
        if (top == (c->low & R_MSB)) {               /* E1 / E2 */
            if (top) { c->high -= R_MSB; c->low -= R_MSB; }
            c->high = (c->high << 1) | 1u; c->low <<= 1;
            if ((e = put(&c->out, top >> 15))) return e;
            /* pending==0 after the loop even on capacity error path in the reference? No: the
               reference returns mid-loop leaving m_nb_e3 partially drained only via the counter
               reset at the end; the error aborts the whole call so it is unobservable. */
            if ((e = flush_pending(c, top >> 15))) return e;
        } else if (c->low > R_QUARTER && c->high <= R_3QUARTER) { /* E3 */
            c->high -= R_QUARTER + 1; c->low -= R_QUARTER + 1;
            c->high = (c->high << 1) | 1u; c->low <<= 1;
            c->pending++;
        } else break;
    }
    return EAE_OK;
}

/* stop_encoding (BinaryArithmeticCoder.cpp:61-102) */
static int bac_finish(bac_t* c)
{
    unsigned bit = c->low < R_QUARTER ? 0u : 1u;
    int e;
    c->pending++;
    if ((e = put(&c->out, bit))) return e;
    if ((e = flush_pending(c, bit))) return e;
    bac_reset(c);
    return EAE_OK;
}

/* start_decoding (BinaryArithmeticCoder.cpp:104-122): 16 bits MSB-first, short streams are padded
   with the LAST bit read (the local `storage` keeps its value). */
static int bac_open(bac_t* c)
{
    unsigned keep = 0;
    for (int i = 0; i < 16; i++) {
        if (c->out.rd != c->out.wr) { int e = get(&c->out, &keep); if (e) return e; }
        c->code = (c->code << 1) | keep;
    }
    return EAE_OK;
}

/* decoding() = decode_bit + rescale_decoding (BinaryArithmeticCoder.cpp:124-135, 254-315).
   `*bit` is left untouched when code is outside [low, high] (malformed stream), as in :262-271. */
static int bac_get(bac_t* c, unsigned* bit, double p)
{
    uint32_t mid; int e = split(c, p, &mid);
    if (e) return e;
    if (c->code >= c->low && c->code <= mid) { c->high = mid; *bit = 0; }
    else if (c->code > mid && c->code <= c->high) { c->low = mid + 1; *bit = 1; }
    unsigned in = 0; /* re-zeroed per call, sticky inside the loop (:278, :300-307) */
    for (;;) {
        if (c->high <= R_HALF) { /* E1: nothing to subtract */ }
        else if (c->low > R_HALF) { c->high -= R_MSB; c->low -= R_MSB; c->code -= R_MSB; }
        else if (c->high <= R_3QUARTER && c->low > R_QUARTER) {
            c->high -= R_QUARTER + 1; c->low -= R_QUARTER + 1; c->code -= R_QUARTER + 1;
        } else break;
        if (c->out.rd != c->out.wr) { if ((e = get(&c->out, &in))) return e; }
        c->high = ((c->high << 1) & R_MAX) | 1u;
        c->low = (c->low << 1) & R_MAX;
        c->code = ((c->code << 1) & R_MAX) | in;
    }
    return EAE_OK;
}

/* utils.cpp:13-28 (floor(log2(double))+1; exact for every uint32 that matters here) */
static unsigned bit_length(uint32_t x) { unsigned n = 0; if (!x) return 1; while (x) { n++; x >>= 1; } return n; }

/* ---- one map's coder: BAC stream + bypass stream (LosslessCoder.h:14-19) ---- */
typedef struct { bac_t bac; bits_t raw; unsigned L; const double* p; } coder_t;

/* LosslessCoder.cpp:232-252 with :167-191 (prefix), :58-111 (EG0), :22-37 (sign). */
static int put_symbol(coder_t* k, int16_t v)
{
    uint32_t a = (uint32_t)abs((int)v);
    int e;
    /* Prefix: min(a, L) ones, then a zero iff a < L. Bin i uses p[i]. With L == 0 the reference
       indexes an empty vector and throws std::out_of_range (-> -2 at the API). */
    if (k->L == 0) return -2;
    uint32_t ones = a < k->L ? a : k->L;
    for (uint32_t i = 0; i < ones; i++) if ((e = bac_put(&k->bac, 1, k->p[i]))) return e;
    if (a < k->L) { if ((e = bac_put(&k->bac, 0, k->p[a]))) return e; }
    else {
        uint32_t x1 = a - k->L + 1;                 /* write_eg0(input): input_plus_1 */
        unsigned n = bit_length(x1) - 1;
        for (unsigned i = 0; i < n; i++) if ((e = put(&k->raw, 1))) return e;
        if ((e = put(&k->raw, 0))) return e;
        uint32_t suffix = x1 - (1u << n);
        for (unsigned i = 0; i < n; i++) if ((e = put(&k->raw, (suffix >> (n - 1 - i)) & 1u))) return e;
    }
    if (v) return put(&k->raw, v < 0 ? 0u : 1u);
    return EAE_OK;
}

/* LosslessCoder.cpp:254-276 with :193-230, :113-165, :39-56. */
static int get_symbol(coder_t* k, int16_t* v)
{
    if (k->L == 0) return -2;
    uint32_t a = 0; unsigned bit = 0; int e;
    for (unsigned i = 0;; i++) {
        if ((e = bac_get(&k->bac, &bit, k->p[i]))) return e;
        if (!bit) break;
        a++;
        if (i == k->L - 1) break;
    }
    if (a == k->L) {
        unsigned n = 0; uint32_t x = 0;
        for (;;) { if ((e = get(&k->raw, &bit))) return e; if (!bit) break; n = (n + 1) & 0xFF; }
        for (unsigned i = 0; i < n; i++) { if ((e = get(&k->raw, &bit))) return e; x = ((x << 1) | bit) & 0xFFFF; }
        x = (x + ((1u << (n & 31)) - 1)) & 0xFFFF;   /* uint16_t arithmetic in the reference */
        a = (a + x) & 0xFFFF;
    }
    int16_t out = (int16_t)(uint16_t)a;
    if (out) { if ((e = get(&k->raw, &bit))) return e; if (!bit) out = (int16_t)-out; }
    *v = out;
    return EAE_OK;
}

/* compression.cpp:24: size * max(32, L) bits for EACH of the two buffers. */
uint32_t oracle_capacity_bits(uint32_t size, uint32_t L) { return round_up(size * (L > 32 ? L : 32), 8); }

static int coder_init(coder_t* k, uint32_t size, uint32_t L, const double* p)
{
    memset(k, 0, sizeof *k);
    uint32_t cap = oracle_capacity_bits(size, L);
    k->bac.out.cap_bits = k->raw.cap_bits = cap;
    k->bac.out.data = (uint8_t*)calloc((cap >> 3) + 1, 1);
    k->raw.data = (uint8_t*)calloc((cap >> 3) + 1, 1);
    if (!k->bac.out.data || !k->raw.data) return -4;
    bac_reset(&k->bac);
    k->L = L; k->p = p;
    return 0;
}

static void coder_free(coder_t* k) { free(k->bac.out.data); free(k->raw.data); }

/* Encode one map; copy out both buffers. Returns the reference error_code (0..4), -1 for a NULL
   pointer (std::invalid_argument, compression.cpp:9-12), -2 for L == 0 (std::out_of_range). */
int oracle_encode_map(uint32_t size, const int16_t* in, uint32_t L, const double* probs,
                      uint8_t* bac_out, uint32_t bac_cap_bytes, uint32_t* bac_bits,
                      uint8_t* byp_out, uint32_t byp_cap_bytes, uint32_t* byp_bits)
{
    if (!in || !probs) return -1;
    coder_t k; int e = coder_init(&k, size, L, probs);
    if (e) return e;
    for (uint32_t i = 0; i < size && !e; i++) e = put_symbol(&k, in[i]);
    if (!e) e = bac_finish(&k.bac);
    if (!e) {
        uint32_t nb = (k.bac.out.wr + 7) >> 3, nr = (k.raw.wr + 7) >> 3;
        *bac_bits = k.bac.out.wr; *byp_bits = k.raw.wr;
        if (bac_out) memcpy(bac_out, k.bac.out.data, nb < bac_cap_bytes ? nb : bac_cap_bytes);
        if (byp_out) memcpy(byp_out, k.raw.data, nr < byp_cap_bytes ? nr : byp_cap_bytes);
    }
    coder_free(&k);
    return e;
}

/* Decode one map from external buffers (the standalone half the reference never exposes). */
int oracle_decode_map(uint32_t size, int16_t* out, uint32_t L, const double* probs,
                      const uint8_t* bac_in, uint32_t bac_bits, const uint8_t* byp_in, uint32_t byp_bits)
{
    if (!out || !probs || !bac_in || !byp_in) return -1;
    coder_t k; int e = coder_init(&k, size, L, probs);
    if (e) return e;
    uint32_t cap = k.bac.out.cap_bits;
    if (bac_bits > cap || byp_bits > cap) { coder_free(&k); return EAE_CAPACITY; }
    memcpy(k.bac.out.data, bac_in, (bac_bits + 7) >> 3); k.bac.out.wr = bac_bits;
    memcpy(k.raw.data, byp_in, (byp_bits + 7) >> 3); k.raw.wr = byp_bits;
    e = bac_open(&k.bac);
    for (uint32_t i = 0; i < size && !e; i++) e = get_symbol(&k, &out[i]);
    coder_free(&k);
    return e;
}

/* compress_lossless (compression.cpp:3-65): encode, count, decode in one call. */
int oracle_compress_lossless(uint32_t size, const int16_t* in, int16_t* out, uint32_t L,
                             const double* probs, uint32_t* nb_bits)
{
    if (!in || !out || !probs) return -1;
    coder_t k; int e = coder_init(&k, size, L, probs);
    if (e) return e;
    for (uint32_t i = 0; i < size && !e; i++) e = put_symbol(&k, in[i]);
    if (!e) e = bac_finish(&k.bac);
    if (!e) {
        *nb_bits = k.bac.out.wr + k.raw.wr;
        e = bac_open(&k.bac);
        for (uint32_t i = 0; i < size && !e; i++) e = get_symbol(&k, &out[i]);
    }
    coder_free(&k);
    return e;
}

/* Batched form used as the CPU baseline ("port"): n maps of `size` symbols, planar input. */
int oracle_compress_maps(uint32_t n_maps, uint32_t size, const int16_t* in_planar, int16_t* out_planar,
                         uint32_t L, const double* probs /* [n_maps * L] */, uint32_t* nb_bits /* [n_maps] */)
{
    for (uint32_t m = 0; m < n_maps; m++) {
        int e = oracle_compress_lossless(size, in_planar + (size_t)m * size, out_planar + (size_t)m * size,
                                         L, probs + (size_t)m * L, &nb_bits[m]);
        if (e) return e;
    }
    return 0;
}

/* Raw BAC over explicit bits (tests.cpp:69-132). */
int oracle_bac_encode_bits(uint32_t n, const uint8_t* bits, const double* probs,
                           uint8_t* out, uint32_t cap_bytes, uint32_t* nb_bits)
{
    bac_t c; memset(&c, 0, sizeof c);
    c.out.data = (uint8_t*)calloc(cap_bytes + 1, 1); c.out.cap_bits = cap_bytes * 8;
    bac_reset(&c);
    int e = 0;
    for (uint32_t i = 0; i < n && !e; i++) e = bac_put(&c, bits[i], probs[i]);
    if (!e) e = bac_finish(&c);
    if (!e) { *nb_bits = c.out.wr; memcpy(out, c.out.data, (c.out.wr + 7) >> 3); }
    free(c.out.data);
    return e;
}

uint32_t oracle_create_divisible(uint32_t x, uint32_t d) { return round_up(x, d); }
uint32_t oracle_count_nb_bits(uint32_t x) { return bit_length(x); }
