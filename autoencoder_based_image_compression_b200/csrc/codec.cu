// Codec handle: weights re-laid-out for the kernels, layer plans for the analysis / synthesis
// transforms (kodak_tensorflow/eae/graph/components.py:11-142) and the fused
// encode -> quantize -> lossless code -> container pipeline and its inverse.
#include <memory>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "conv_plan.cuh"
#include "internal.cuh"
#include "transforms.cuh"

using namespace eae;

// One step of the pipeline = ~25 dependent launches. From its second use with the same buffers, the part of a step that
// only touches the codec's workspace and the caller's output buffers is replayed as ONE CUDA graph (captured from the very
// launch sequence below): +4 % images/s with 16 pipeline slots (scripts/graph_probe.py), less host time per step.
struct StepGraphKey {
    int kind;                      // 0 compress (layers 2-3, quantizer, coder, container), 1 decompress (everything)
    uint32_t n, h, w, L;
    const void* p0; const void* p1; const void* p2; const void* p3;
    uint64_t cap;
    int math;
    uint32_t lanes;
    uint64_t generation;           // of the codec's device buffers
};
struct StepGraph {
    StepGraphKey key;
    cudaGraphExec_t exec = nullptr;
    uint32_t uses = 0;
    int launches = 0;              // kernel launches one replay stands for (eae_launch_count)
    uint64_t last_use = 0;
};

struct eae_codec {
    int device = 0;
    int learned = 0;   // are_bin_widths_learned: 4 GDN/IGDN instead of 6
    int math = EAE_MATH_FP32_SIMT;
    bool exact_now = true; // do the contractions of the transform being run use the 3-way split (set per chunk)
    int umma_mask = 0xF;   // which layer kinds run on tensor cores (debug: env EAE_UMMA_LAYERS)
    uint32_t coder_lanes = 0;   // threads per coded stream: 0 = auto (one warp per stream while they fit: lowest latency)
    int no_fuse = 0;       // debug: env EAE_NO_FUSE=1 keeps GDN / IGDN as separate launches
    int no_direct_conv1 = 0;   // debug: env EAE_NO_DIRECT_CONV1=1 keeps the im2col pass in front of layer 1
    int no_fuse_quant = 0;     // debug: env EAE_NO_FUSE_QUANT=1 keeps the quantizer / dequantizer as separate launches
    int gdn_precise = 2;       // env EAE_GDN_PRECISE: see run_layer
    // The four output phases of a transposed convolution as ONE grid: 40 us less per 24-image step when the transforms run
    // alone (the phases of a tile share their input box in L2, six dependent launches disappear). With 16 pipeline slots it
    // wins 1.6 % when the step is replayed as a graph (1.59 vs 1.615 ms per step) and loses 3 % when it is launched kernel
    // by kernel (1.77 vs 1.72 ms: a 2304-CTA grid keeps the other slots' small kernels waiting longer than four 576-CTA
    // grids do). phase_merge: -1 = where the step runs as a graph (default), 0 = never, 1 = always (env EAE_PHASE_MERGE).
    int phase_merge = -1;
    int no_phase_merge = 1;    // the decision for the call being served
    cudaStream_t own_stream = nullptr;
    // Experiment (env EAE_CODER_PRIORITY=1 / 2, off by default): the lossless-coding kernels on a side stream of the
    // greatest / least priority, forked from and joined to the caller's stream with events. Measured with 16 pipeline
    // slots: 1.87 ms per step with the side stream at either priority against 1.78 ms without it.
    cudaStream_t coder_stream = nullptr;
    cudaEvent_t fork_event = nullptr, join_event = nullptr;
    int coder_priority = 0;
    int use_graphs = 1;            // env EAE_GRAPHS=0 disables the step graphs
    int graphs_in_host_calls = 0;  // experiment: env EAE_GRAPHS_HOST=1
    bool host_call = false;        // set by the _host entry points: they launch directly (with one host thread per pipeline
                                   // slot, replaying graphs changes the end-to-end rate by less than its run-to-run spread,
                                   // while the device-resident entry points, driven by one thread, gain 6 %)
    uint64_t generation = 0;       // bumped whenever a device buffer of the codec is (re)allocated
    uint64_t graph_clock = 0;
    std::vector<StepGraph> graphs;

    // ---- weights (device) ----
    DevBuf w1m;            // [96][128]   im2col matrix of weights_1 (rows >= 81 are zero)
    DevBuf w2, w3;         // [25][128 in][128 out]  = TF [kh,kw,in,out] as is
    DevBuf w4, w5;         // [25][128 in][128 out]  = TF [kh,kw,out,in] transposed per tap
    DevBuf w6m;            // [128 in][128]  column ky*9+kx of weights_6 (columns >= 81 are zero)
    DevBuf gamma[6], beta[6], bias[5];
    // K-major copies (hi / lo tf32 split) for the tcgen05 path: [tap][128 out][Cin]
    DevBuf wk_hi[6], wk_lo[6], gk_hi[6], gk_lo[6];

    // ---- transform workspace for `ws_n` images of ws_h x ws_w ----
    uint32_t ws_n = 0, ws_h = 0, ws_w = 0;
    DevBuf bufA, buf1, buf2, buf3, img_u8, rec_u8;

    // ---- coder workspace ----
    uint32_t cw_streams = 0, cw_size = 0, cw_L = 0, cw_slot = 0;
    DevBuf idx_planar, bac_slots, byp_slots, bac_bits, byp_bits, err, bac_off, byp_off, total_bytes, enc_scratch;
    DevBuf table, mean, delta, flag, stats;
    DevBuf qtable, row_flags;          // coder v3: fixed-point multipliers + per-row validity (prepare_table_kernel)
    std::vector<double> table_host;    // the table the device copies were made from
    std::vector<float> delta_host, mean_host;
    uint64_t last_idx_elems = 0;
    // pinned, device-visible result block of the _host entry points: one kernel writes it, one synchronisation reads it
    struct HostMailbox* mailbox = nullptr;
    void* status_host = nullptr;          // pinned HostStatus of eae_codec_poll_status
    cudaStream_t last_stream = nullptr;   // stream of the last compress / decompress step
    uint32_t last_n_streams = 0;
    eae_batch_stats_t* stats_acc = nullptr;   // device accumulator of the per-step statistics (caller-owned), or NULL
};

struct HostMailbox {
    uint64_t total;
    uint32_t flag[2];       // [0] bit flags (int16 overflow), [1] first coder error
    uint32_t umma_err;      // time-out mask of the tensor path (cleared on read)
    uint32_t pad;
    eae_batch_stats_t stats;
};

namespace {

constexpr uint32_t kHeaderBytes = 32;
constexpr uint32_t kMagic = 0x42454145u;  // 'EAEB' little-endian
constexpr uint32_t kVersion = 1;

// Images per internal chunk: keeps the fp32 activation workspace (about 29 MB per 512x768 image)
// within a few GB regardless of the batch the caller passes.
uint32_t chunk_images(uint32_t h, uint32_t w)
{
    const uint64_t per_image = (uint64_t)(h / 4) * (w / 4) * 128 * 4 * 2 + (uint64_t)(h / 8) * (w / 8) * 128 * 4 +
                               (uint64_t)(h / 16) * (w / 16) * 128 * 4;
    uint64_t n = (6ull << 30) / per_image;
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    if (const char* env = getenv("EAE_CHUNK_IMAGES")) {      // tests: exercise the multi-chunk path on small batches
        const long v = atol(env);
        if (v >= 1 && (uint64_t)v < n) n = (uint64_t)v;
    }
    return (uint32_t)n;
}

int upload(DevBuf& d, const float* host, size_t n)
{
    EAE_TRY(d.alloc(n * sizeof(float)));
    EAE_CUDA_OK(cudaMemcpy(d.p, host, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// tf32 split used by the 3xTF32 mode: hi = round-to-nearest-even to 10 mantissa bits, lo = x - hi
// (exact in fp32), then lo rounded the same way.
float tf32_round(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u += 0x00000FFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int upload_split(DevBuf& hi, DevBuf& lo, const std::vector<float>& v)
{
    std::vector<float> h(v.size()), l(v.size());
    for (size_t i = 0; i < v.size(); i++) { h[i] = tf32_round(v[i]); l[i] = tf32_round(v[i] - h[i]); }
    EAE_TRY(upload(hi, h.data(), h.size()));
    EAE_TRY(upload(lo, l.data(), l.size()));
    return 0;
}

int ensure_workspace(eae_codec* c, uint32_t n, uint32_t h, uint32_t w)
{
    if (c->ws_n >= n && c->ws_h == h && c->ws_w == w) return 0;
    const size_t p1 = (size_t)(h / 4) * (w / 4), p2 = (size_t)(h / 8) * (w / 8), p3 = (size_t)(h / 16) * (w / 16);
    EAE_TRY(c->bufA.alloc(n * p1 * 128 * 4));
    EAE_TRY(c->buf1.alloc(n * p1 * 128 * 4));
    EAE_TRY(c->buf2.alloc(n * p2 * 128 * 4));
    EAE_TRY(c->buf3.alloc(n * p3 * 128 * 4));
    c->ws_n = n; c->ws_h = h; c->ws_w = w;
    c->generation++;
    return 0;
}

int ensure_coder(eae_codec* c, uint32_t n_streams, uint32_t size, uint32_t L)
{
    const uint32_t slot = eae_coder_slot_bytes(size, L);
    if (c->cw_streams >= n_streams && c->cw_size == size && c->cw_L == L) return 0;
    EAE_TRY(c->idx_planar.alloc((size_t)n_streams * size * 2));
    EAE_TRY(c->bac_slots.alloc((size_t)n_streams * slot));
    EAE_TRY(c->byp_slots.alloc((size_t)n_streams * slot));
    EAE_TRY(c->bac_bits.alloc((size_t)n_streams * 4));
    EAE_TRY(c->byp_bits.alloc((size_t)n_streams * 4));
    EAE_TRY(c->err.alloc((size_t)n_streams * 4));
    EAE_TRY(c->bac_off.alloc((size_t)n_streams * 8));
    EAE_TRY(c->byp_off.alloc((size_t)n_streams * 8));
    EAE_TRY(c->enc_scratch.alloc(coder_encode_scratch_bytes(n_streams, size, L)));
    if (!c->total_bytes.p) EAE_TRY(c->total_bytes.alloc(8));
    // [0] bit flags of the step (int16 overflow), [1] first coder error of the step, [2] [3] sticky status (kStatus*)
    if (!c->flag.p) { EAE_TRY(c->flag.alloc(16)); EAE_CUDA_OK(cudaMemset(c->flag.p, 0, 16)); }
    if (!c->stats.p) EAE_TRY(c->stats.alloc(sizeof(eae_batch_stats_t)));
    c->cw_streams = n_streams; c->cw_size = size; c->cw_L = L; c->cw_slot = slot;
    c->generation++;
    return 0;
}

int upload_params(eae_codec* c, const eae_coding_params_t* prm, cudaStream_t st)
{
    if (!prm || !prm->bin_widths || !prm->table) { set_error("coding params: NULL pointer"); return EAE_ERR_NULL; }
    const uint32_t L = prm->truncated_unary_length;
    if (L == 0) { set_error("truncated unary length is 0"); return EAE_ERR_UNARY_LENGTH; }
    if (L > 255) { set_error("truncated unary length %u exceeds 255", L); return EAE_ERR_ARGUMENT; }
    for (int i = 0; i < EAE_NB_MAPS; i++)
        if (!(prm->bin_widths[i] > 0.f)) { set_error("A quantization bin width is not strictly positive."); return EAE_ERR_ARGUMENT; }
    // The parameters rarely change between calls: upload (and re-derive the coder's fixed-point table) only
    // when they differ from what the device already holds.
    const size_t n_tab = (size_t)EAE_NB_MAPS * L;
    if (c->table_host.size() != n_tab || memcmp(c->table_host.data(), prm->table, n_tab * 8) != 0) {
        if (c->table.bytes < n_tab * 8) {
            EAE_TRY(c->table.alloc(n_tab * 8));
            EAE_TRY(c->qtable.alloc(n_tab * 8));
            c->generation++;
        }
        if (!c->row_flags.p) { EAE_TRY(c->row_flags.alloc(EAE_NB_MAPS)); c->generation++; }
        c->table_host.assign(prm->table, prm->table + n_tab);
        // pageable source: the copy is staged before the call returns, so the vector may change afterwards
        EAE_CUDA_OK(cudaMemcpyAsync(c->table.p, c->table_host.data(), n_tab * 8, cudaMemcpyHostToDevice, st));
        EAE_TRY(launch_prepare_table(c->table.as<double>(), EAE_NB_MAPS, L, c->qtable.as<uint64_t>(),
                                     c->row_flags.as<uint8_t>(), st));
    }
    if (!c->mean.p) { EAE_TRY(c->mean.alloc(EAE_NB_MAPS * 4)); c->generation++; }
    if (!c->delta.p) { EAE_TRY(c->delta.alloc(EAE_NB_MAPS * 4)); c->generation++; }
    if (c->delta_host.size() != EAE_NB_MAPS || memcmp(c->delta_host.data(), prm->bin_widths, EAE_NB_MAPS * 4) != 0) {
        c->delta_host.assign(prm->bin_widths, prm->bin_widths + EAE_NB_MAPS);
        EAE_CUDA_OK(cudaMemcpyAsync(c->delta.p, c->delta_host.data(), EAE_NB_MAPS * 4, cudaMemcpyHostToDevice, st));
    }
    std::vector<float> mean(EAE_NB_MAPS, 0.f);
    if (prm->map_mean) mean.assign(prm->map_mean, prm->map_mean + EAE_NB_MAPS);
    if (c->mean_host != mean) {
        c->mean_host = mean;
        EAE_CUDA_OK(cudaMemcpyAsync(c->mean.p, c->mean_host.data(), EAE_NB_MAPS * 4, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

int check_dims(uint32_t n, uint32_t h, uint32_t w)
{
    // EntropyAutoencoder.py:77-80, IsolatedDecoder.py:50-53
    if (h == 0 || w == 0 || h % EAE_STRIDE_PROD != 0 || w % EAE_STRIDE_PROD != 0) {
        set_error("The height/width of the images (%u x %u) is not divisible by the product of the three strides.", h, w);
        return EAE_ERR_ARGUMENT;
    }
    if ((uint64_t)n * (h / 4) * (w / 4) >= (1ull << 31)) { set_error("batch too large"); return EAE_ERR_ARGUMENT; }
    return 0;
}

// ---- layer launchers --------------------------------------------------------------------------
enum LayerKind : int { kLayerConv = 0, kLayerTconv = 1, kLayerGdn = 2, kLayerThin = 3 };

int run_gemm(eae_codec* c, const GemmPlan& plan, int kind, const UmmaWeights& uw, const UmmaWeights* gamma,
             cudaStream_t st)
{
    static const int prof_of_kind[4] = {kProfGemmConv, kProfGemmTconv, kProfGemmGdn, kProfGemmThin};
    ProfScope prof(prof_of_kind[kind], st);
    if (c->math == EAE_MATH_FP32_SIMT || !((c->umma_mask >> kind) & 1)) return launch_gemm_simt(plan, st);
    // GDN / IGDN always use the split contraction: a single-pass TF32 norm would put a 2^-12 relative
    // error on every activation, for 4 % of the FLOPs.
    const bool exact = c->exact_now || kind == kLayerGdn;
    return launch_gemm_umma(plan, uw, gamma, exact, st);
}

// Can the GDN / IGDN that follows a contraction of this kind run inside its epilogue?
bool can_fuse(const eae_codec* c, int kind)
{
    return c->math != EAE_MATH_FP32_SIMT && ((c->umma_mask >> kind) & 1) &&
           ((c->umma_mask >> kLayerGdn) & 1) && !c->no_fuse;
}

GemmPlan base_plan(const float* in, int Hin, int Win, int Cin, const float* w, const float* bias, float* out,
                   uint32_t n)
{
    GemmPlan p;
    memset(&p, 0, sizeof p);
    p.in = in; p.w = w; p.bias = bias; p.out = out;
    p.Hin = Hin; p.Win = Win; p.Cin = Cin;
    p.Hg = Hin; p.Wg = Win; p.in_mul = 1;
    p.Hout = Hin; p.Wout = Win; p.out_mul = 1; p.out_r = 0; p.out_s = 0;
    p.mode = kEpiBias; p.n_taps = 1;
    p.taps[0] = Tap{0, 0, 0};
    p.M = n * (uint32_t)Hin * (uint32_t)Win;
    return p;
}

UmmaWeights umma_weights(eae_codec* c, int layer, int n_taps)
{
    return UmmaWeights{c->wk_hi[layer].as<float>(), c->wk_lo[layer].as<float>(), n_taps};
}

// GDN / IGDN as a 1-tap contraction with the squared input (tfutils.py:393-397, 506-509). Pixel-wise,
// hence indifferent to the storage order of the pixels: the plan is a flat [n_pixels, 128] matrix.
// dq (IGDN on the tensor path only): the input is the dequantized planar indices instead of `in` (conv_plan.cuh).
struct DequantIn { const int16_t* idx; const float* mean; const float* delta; int hw; };

int run_gdn(eae_codec* c, const float* in, float* out, uint32_t n_pixels, int which, bool inverse, cudaStream_t st,
            const DequantIn* dq = nullptr)
{
    GemmPlan p = base_plan(in, 1, (int)n_pixels, 128, c->gamma[which].as<float>(), c->beta[which].as<float>(), out, 1);
    p.mode = inverse ? kEpiIgdn : kEpiGdn;
    if (dq) { p.dequant_idx = dq->idx; p.dequant_mean = dq->mean; p.dequant_delta = dq->delta; p.dequant_hw = dq->hw; }
    return run_gemm(c, p, kLayerGdn, UmmaWeights{c->gk_hi[which].as<float>(), c->gk_lo[which].as<float>(), 1}, nullptr, st);
}

// One contraction followed by GDN / IGDN number `gdn` (-1: none): inside the epilogue on the tensor path,
// as a second launch otherwise. `plans` are the launches that together produce `out_pixels` pixels at out.
int run_layer(eae_codec* c, GemmPlan* plans, int n_plans, int kind, const UmmaWeights& uw, int gdn, bool inverse,
              uint32_t out_pixels, cudaStream_t st)
{
    const bool fuse = gdn >= 0 && can_fuse(c, kind);
    const UmmaWeights gw{gdn >= 0 ? c->gk_hi[gdn].as<float>() : nullptr, gdn >= 0 ? c->gk_lo[gdn].as<float>() : nullptr, 1};
    for (int i = 0; i < n_plans; i++) {
        if (fuse) {
            plans[i].fuse = inverse ? 2 : 1; plans[i].fuse_beta = c->beta[gdn].as<float>();
            // mixed mode, synthesis side: the norm of a fused IGDN in one rounded-TF32 pass as well (its bar is the PSNR)
            plans[i].fuse_single_pass = (c->math == EAE_MATH_MIXED && !c->exact_now) ? 1 : 0;
            // IEEE sqrt / division where the norm itself is the 3xTF32 one: gdn_precise 2 = every such layer (default),
            // 1 = only GDN3, whose output is quantized, 0 = nowhere (MUFU forms; measurement knob)
            plans[i].fuse_precise = !plans[i].fuse_single_pass && (c->gdn_precise >= 2 || (c->gdn_precise == 1 && gdn == 2)) ? 1 : 0;
        }
    }
    const bool tensor = c->math != EAE_MATH_FP32_SIMT && ((c->umma_mask >> kind) & 1);
    if (n_plans > 1 && n_plans <= 4 && tensor && !c->no_phase_merge) {
        // the output phases of a transposed convolution as one grid (kernel version 4)
        static const int prof_of_kind[4] = {kProfGemmConv, kProfGemmTconv, kProfGemmGdn, kProfGemmThin};
        ProfScope prof(prof_of_kind[kind], st);
        EAE_TRY(launch_gemm_umma(plans[0], uw, fuse ? &gw : nullptr, c->exact_now, st, plans + 1, n_plans - 1));
    } else {
        for (int i = 0; i < n_plans; i++) EAE_TRY(run_gemm(c, plans[i], kind, uw, fuse ? &gw : nullptr, st));
    }
    if (gdn >= 0 && !fuse) EAE_TRY(run_gdn(c, plans[0].out, plans[0].out, out_pixels, gdn, inverse, st));
    return 0;
}

// conv k5 s2 SAME: out grid = in / 2, taps (ky - 1, kx - 1) (TF pads 1 before, 2 after). The input is
// stored parity-split (written so by its producer).
// quant: the quantizer of the latent fused into this layer's store (tensor path), or NULL.
struct QuantOut { int16_t* idx; const float* mean; const float* delta; uint32_t* flag; };

int run_conv5s2(eae_codec* c, const float* in, int Hin, int Win, const float* w, int layer, const float* bias,
                float* out, bool out_split, int gdn, uint32_t n, cudaStream_t st, const QuantOut* quant = nullptr)
{
    GemmPlan p = base_plan(in, Hin, Win, 128, w, bias, out, n);
    if (quant) { p.quant_idx = quant->idx; p.quant_mean = quant->mean; p.quant_delta = quant->delta; p.quant_flag = quant->flag; }
    p.Hg = Hin / 2; p.Wg = Win / 2; p.in_mul = 2;
    p.Hout = p.Hg; p.Wout = p.Wg;
    p.in_split = 1;
    p.out_split = out_split ? 1 : 0;
    p.n_taps = 25;
    for (int ky = 0; ky < 5; ky++)
        for (int kx = 0; kx < 5; kx++)
            p.taps[ky * 5 + kx] = Tap{ky - 1, kx - 1, (uint32_t)(ky * 5 + kx) * 128u * 128u};
    p.M = n * (uint32_t)p.Hg * (uint32_t)p.Wg;
    return run_layer(c, &p, 1, kLayerConv, umma_weights(c, layer, 25), gdn, false, p.M, st);
}

// conv2d_transpose k5 s2 SAME = 4 output phases; out[2a + r] gathers in[a + dy] through ky = r + 1 - 2 dy.
int run_tconv5s2(eae_codec* c, const float* in, int Hin, int Win, const float* w, int layer, const float* bias,
                 float* out, int gdn, uint32_t n, cudaStream_t st)
{
    GemmPlan plans[4];
    for (int r = 0; r < 2; r++) {
        for (int s = 0; s < 2; s++) {
            GemmPlan& p = plans[r * 2 + s];
            p = base_plan(in, Hin, Win, 128, w, bias, out, n);
            p.Hout = 2 * Hin; p.Wout = 2 * Win; p.out_mul = 2; p.out_r = r; p.out_s = s;
            int nt = 0;
            for (int ky = 0; ky < 5; ky++) {
                if (((r + 1 - ky) & 1) != 0) continue;
                const int dy = (r + 1 - ky) / 2;
                for (int kx = 0; kx < 5; kx++) {
                    if (((s + 1 - kx) & 1) != 0) continue;
                    const int dx = (s + 1 - kx) / 2;
                    p.taps[nt++] = Tap{dy, dx, (uint32_t)(ky * 5 + kx) * 128u * 128u};
                }
            }
            p.n_taps = nt;
        }
    }
    return run_layer(c, plans, 4, kLayerTconv, umma_weights(c, layer, 25), gdn, true, 4 * plans[0].M, st);
}

// parts: bit 0 = layer 1 (the only one that reads the caller's images), bit 1 = layers 2 and 3
// quant != NULL: the latent leaves layer 3 as planar int16 indices (y_dev is not written). Only where
// quantizer_fusable() says so.
bool quantizer_fusable(const eae_codec* c, uint32_t h)
{
    return c->math != EAE_MATH_FP32_SIMT && ((c->umma_mask >> kLayerConv) & 1) && umma_can_fuse_quantizer((int)(h / 16)) &&
           (c->learned || can_fuse(c, kLayerConv)) && !c->no_fuse_quant;
}

int encode_chunk(eae_codec* c, const uint8_t* img_dev, uint32_t n, uint32_t h, uint32_t w, float* y_dev,
                 cudaStream_t st, int parts = 3, const QuantOut* quant = nullptr)
{
    const int H1 = h / 4, W1 = w / 4, H2 = h / 8, W2 = w / 8;
    c->exact_now = c->math == EAE_MATH_TF32X3 || c->math == EAE_MATH_MIXED;      // the indices are decided here
    float* A = c->bufA.as<float>();
    float* x1 = c->buf1.as<float>();
    float* x2 = c->buf2.as<float>();
    // layer 1: conv k9 s4 (1 -> 128) as one 96-deep contraction (81 taps used) + GDN; output parity-split for the
    // stride-2 layer that follows. On the tensor path (kernel version 3) the patches are gathered from the uint8
    // image inside the kernel; otherwise an im2col pass writes them out first.
    if (parts & 1) {
        const bool direct = c->math != EAE_MATH_FP32_SIMT && ((c->umma_mask >> kLayerThin) & 1) &&
                            !c->no_direct_conv1;
        if (!direct) { ProfScope prof(kProfIm2col, st); EAE_TRY(launch_im2col_k9s4(img_dev, A, n, (int)h, (int)w, st)); }
        GemmPlan p = base_plan(A, H1, W1, kIm2colK, c->w1m.as<float>(), c->bias[0].as<float>(), x1, n);
        p.out_split = 1;
        if (direct) { p.img_u8 = img_dev; p.img_H = (int)h; p.img_W = (int)w; }
        EAE_TRY(run_layer(c, &p, 1, kLayerThin, umma_weights(c, 0, 1), 0, false, p.M, st));
    }
    if (!(parts & 2)) return 0;
    // layer 2 (output parity-split again), layer 3 (natural NHWC: it is the latent the API returns)
    EAE_TRY(run_conv5s2(c, x1, H1, W1, c->w2.as<float>(), 1, c->bias[1].as<float>(), x2, true, 1, n, st));
    EAE_TRY(run_conv5s2(c, x2, H2, W2, c->w3.as<float>(), 2, c->bias[2].as<float>(), y_dev, false,
                        c->learned ? -1 : 2, n, st, quant));
    return 0;
}

// Can the dequantizer run inside the operand load of the decoder's first IGDN (fixed bin widths, tensor path)?
bool dequantizer_fusable(const eae_codec* c)
{
    return !c->learned && c->math != EAE_MATH_FP32_SIMT && ((c->umma_mask >> kLayerGdn) & 1) && !c->no_fuse_quant;
}

// dq != NULL (only where dequantizer_fusable()): the latent is read as planar int16 indices, q_dev is not.
int decode_chunk(eae_codec* c, const float* q_dev, uint32_t n, uint32_t h, uint32_t w, uint8_t* out_u8_dev,
                 float* out_f32_dev, cudaStream_t st, const DequantIn* dq = nullptr)
{
    const int H1 = h / 4, W1 = w / 4, H2 = h / 8, W2 = w / 8, H3 = h / 16, W3 = w / 16;
    c->exact_now = c->math == EAE_MATH_TF32X3;
    float* P = c->bufA.as<float>();
    float* x1 = c->buf1.as<float>();
    float* x2 = c->buf2.as<float>();
    float* x3 = c->buf3.as<float>();
    const float* src = q_dev;
    if (!c->learned) {
        EAE_TRY(run_gdn(c, q_dev, x3, n * (uint32_t)(H3 * W3), 3, true, st, dq));
        src = x3;
    }
    EAE_TRY(run_tconv5s2(c, src, H3, W3, c->w4.as<float>(), 3, c->bias[3].as<float>(), x2, 4, n, st));
    EAE_TRY(run_tconv5s2(c, x2, H2, W2, c->w5.as<float>(), 4, c->bias[4].as<float>(), x1, 5, n, st));
    // layer 6: conv2d_transpose k9 s4 (128 -> 1), no bias. Kernel version 6 contracts, gathers and casts in one launch;
    // otherwise: per-pixel tap contributions, then col2im.
    if (c->math != EAE_MATH_FP32_SIMT && ((c->umma_mask >> kLayerThin) & 1)) {
        ProfScope prof(kProfGemmThin, st);
        return launch_tconv9s4_fused(x1, umma_weights(c, 5, 1), out_u8_dev, out_f32_dev, n, (int)h, (int)w, c->exact_now, st);
    }
    {
        GemmPlan p = base_plan(x1, H1, W1, 128, c->w6m.as<float>(), nullptr, P, n);
        EAE_TRY(run_layer(c, &p, 1, kLayerThin, umma_weights(c, 5, 1), -1, false, p.M, st));
    }
    { ProfScope prof(kProfCol2im, st); EAE_TRY(launch_col2im_k9s4(P, out_u8_dev, out_f32_dev, n, (int)h, (int)w, st)); }
    return 0;
}

// ---- container kernels -------------------------------------------------------------------------
// Exclusive scan of per-stream byte sizes -> payload offsets. Single CTA, 1024 threads. The same pass over the stream
// table also does what used to be three more launches and two strided copies per step:
//   compress   (header != NULL): container header + stream table, per-map bit totals / total bits / dead maps of the
//                                batch (eae_batch_stats_t, no memset needed), first coder error -> flag[1]
//   decompress (bac_out != NULL): de-interleaves the container's stream table into the bit-count arrays the decoder reads
static_assert(1024 % EAE_NB_MAPS == 0, "stream_offsets_kernel: a thread must always see the same map");
struct OffsetsExtra {
    uint32_t* header;            // container as uint32 words, or NULL
    uint32_t n, h, w, L;
    eae_batch_stats_t* stats;    // or NULL
    const uint32_t* err;         // per-stream coder error codes (with stats)
    uint32_t* flag;              // flag[1] = an error code if any stream failed; flag[2], flag[3]: sticky status (kStatus*)
    uint32_t* bac_out;           // or NULL
    uint32_t* byp_out;
    eae_batch_stats_t* acc;      // compress: running totals over steps (atomic adds), or NULL
    uint64_t limit;              // compress: capacity of the container; decompress: readable bytes of the container
    uint32_t cap_bits;           // decompress: largest bit count a stream buffer may have (compression.cpp:24)
};
// Sticky status word flag[2] of a codec: set by the steps, read and cleared by eae_codec_poll_status. flag[3] = the first
// coder error code (1..4) since the last poll.
constexpr uint32_t kStatusInt16 = 1u, kStatusNoRoom = 2u, kStatusBadTable = 4u, kStatusTruncated = 8u;

__global__ void __launch_bounds__(1024)
stream_offsets_kernel(const uint32_t* __restrict__ bac_bits, const uint32_t* __restrict__ byp_bits,
                      uint32_t bits_stride, uint32_t n, uint64_t base, uint64_t* __restrict__ bac_off,
                      uint64_t* __restrict__ byp_off, uint64_t* __restrict__ total_out, const OffsetsExtra x)
{
    __shared__ uint64_t warp_sum[32];
    __shared__ uint64_t running;
    __shared__ unsigned long long red[1024];
    __shared__ uint32_t red_dead[1024];
    __shared__ uint32_t first_err, bad_container;
    if (threadIdx.x == 0) { running = base; first_err = 0; bad_container = 0; }
    // stream s = start + thread with start a multiple of 1024 = 8 x 128: a thread always sees the same map (s % 128), so
    // the per-map totals accumulate in registers and meet once at the end (no atomics on the way)
    unsigned long long my_bits = 0;
    uint32_t my_dead = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t start = 0; start < n; start += 1024) {
        const uint32_t s = start + threadIdx.x;
        uint64_t nb = 0, nr = 0;
        if (s < n) {
            uint32_t bb = bac_bits[(size_t)s * bits_stride], rb = byp_bits[(size_t)s * bits_stride];
            if (x.bac_out && (bb > x.cap_bits || rb > x.cap_bits)) {
                // a stream table no encoder can have written: the stream is read as empty (the decoder then reports a
                // resource error for it) and the container is flagged
                bb = 0; rb = 0;
                atomicOr(&bad_container, kStatusBadTable);
            }
            nb = (bb + 7u) >> 3;
            nr = (rb + 7u) >> 3;
            if (x.header) { x.header[8 + 2 * (size_t)s] = bb; x.header[8 + 2 * (size_t)s + 1] = rb; }
            if (x.bac_out) { x.bac_out[s] = bb; x.byp_out[s] = rb; }
            if (x.stats) {
                my_bits += (unsigned long long)bb + rb;
                // A map is dead (tools.py:294-320) iff all its symbols are 0 iff no sign bit was written.
                my_dead += rb == 0 ? 1u : 0u;
                if (x.err[s]) atomicCAS(&first_err, 0u, x.err[s]);      // (rare)
            }
        }
        uint64_t v = nb + nr;
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_sum[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint64_t ws = warp_sum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const uint64_t incl = v + (wid ? warp_sum[wid - 1] : 0);
        const uint64_t at = running + incl - (nb + nr);
        if (s < n) {
            bac_off[s] = at; byp_off[s] = at + nb;
            if (x.bac_out && at + nb + nr > x.limit) {      // payload runs past the bytes the caller vouches for: never read
                bac_off[s] = base; byp_off[s] = base;
                x.bac_out[s] = 0; x.byp_out[s] = 0;
                atomicOr(&bad_container, kStatusTruncated);
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) running += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total_out = running;
        if (x.flag) {
            uint32_t sticky = bad_container;
            if (x.header && running > x.limit) sticky |= kStatusNoRoom;      // pack_payload_kernel drops what does not fit
            if (x.header && (x.flag[0] & 1u)) sticky |= kStatusInt16;        // (the quantizer ran before this kernel)
            if (sticky) atomicOr(x.flag + 2, sticky);
        }
        if (x.header) {
            x.header[0] = kMagic; x.header[1] = kVersion; x.header[2] = x.n; x.header[3] = x.h; x.header[4] = x.w;
            x.header[5] = EAE_NB_MAPS; x.header[6] = x.L; x.header[7] = 0;
        }
    }
    if (x.stats) {
        red[threadIdx.x] = my_bits;
        red_dead[threadIdx.x] = my_dead;
        __syncthreads();
        if (threadIdx.x < EAE_NB_MAPS) {
            unsigned long long bits = 0;
            uint32_t dead = 0;
            for (int k = 0; k < 1024 / EAE_NB_MAPS; k++) { bits += red[threadIdx.x + k * EAE_NB_MAPS]; dead += red_dead[threadIdx.x + k * EAE_NB_MAPS]; }
            x.stats->bits_per_map[threadIdx.x] = bits;
            if (x.acc) atomicAdd(reinterpret_cast<unsigned long long*>(&x.acc->bits_per_map[threadIdx.x]), bits);
            red[threadIdx.x] = bits;
            red_dead[threadIdx.x] = dead;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long total = 0, dead = 0;
            for (int m = 0; m < EAE_NB_MAPS; m++) { total += red[m]; dead += red_dead[m]; }
            x.stats->total_bits = total;
            x.stats->nb_dead_maps = dead;
            if (x.acc) {
                atomicAdd(reinterpret_cast<unsigned long long*>(&x.acc->total_bits), total);
                atomicAdd(reinterpret_cast<unsigned long long*>(&x.acc->nb_dead_maps), dead);
            }
            if (first_err) { atomicCAS(x.flag + 1, 0u, first_err); atomicCAS(x.flag + 3, 0u, first_err); }      // first stream error wins
        }
    }
}

// One warp per stream: slot bytes -> payload (destination byte-aligned only).
__global__ void __launch_bounds__(256)
pack_payload_kernel(uint8_t* __restrict__ container, uint64_t cap, const uint8_t* __restrict__ bac_slots,
                    const uint8_t* __restrict__ byp_slots, uint32_t slot_bytes,
                    const uint32_t* __restrict__ bac_bits, const uint32_t* __restrict__ byp_bits,
                    const uint64_t* __restrict__ bac_off, const uint64_t* __restrict__ byp_off, uint32_t n_streams)
{
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    const uint32_t nb = (bac_bits[s] + 7u) >> 3, nr = (byp_bits[s] + 7u) >> 3;
    const uint64_t ob = bac_off[s], orr = byp_off[s];
    if (orr + nr > cap) return;   // host checks total_bytes against the capacity
    const uint8_t* sb = bac_slots + (size_t)s * slot_bytes;
    const uint8_t* sr = byp_slots + (size_t)s * slot_bytes;
    for (uint32_t i = lane; i < nb; i += 32) container[ob + i] = sb[i];
    for (uint32_t i = lane; i < nr; i += 32) container[orr + i] = sr[i];
}

// Everything a _host entry point needs to know about the batch, written straight into pinned host memory.
__global__ void mailbox_kernel(HostMailbox* __restrict__ mb, const uint64_t* __restrict__ total,
                               const uint32_t* __restrict__ flag, const eae_batch_stats_t* __restrict__ stats,
                               uint32_t* __restrict__ umma_flag)
{
    if (threadIdx.x == 0) {
        mb->total = total ? *total : 0ull;
        mb->flag[0] = flag[0];
        mb->flag[1] = flag[1];
        mb->umma_err = umma_flag ? atomicExch(umma_flag, 0u) : 0u;
    }
    if (stats) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(stats);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&mb->stats);
        for (uint32_t i = threadIdx.x; i < sizeof(eae_batch_stats_t) / 4; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
}

// dst[0 .. *nbytes) = src[0 .. *nbytes), 16 bytes per thread: the container goes to pinned host memory in wide,
// coalesced writes without the host having to learn its size first (both buffers are 16-byte aligned).
__global__ void copy_prefix_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                   const uint64_t* __restrict__ nbytes, uint64_t cap)
{
    const uint64_t n = *nbytes < cap ? *nbytes : cap;
    const uint64_t n16 = (n + 15) / 16;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// flag[1] = error code of the lowest-numbered stream with an error (0 if none), as the host loop used to find it.
__global__ void first_error_kernel(const uint32_t* __restrict__ err, uint32_t n_streams, uint32_t* __restrict__ key)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_streams && err[s]) atomicMin(key, (s << 8) | (err[s] & 0xFFu));
}
__global__ void first_error_finish_kernel(uint32_t* key, uint32_t* flag1)
{
    *flag1 = *key == 0xFFFFFFFFu ? 0u : (*key & 0xFFu);
}

// (see prefer_max_shared in common.cuh)
void codec_carveouts()
{
    static uint64_t seen = 0;
    if (!first_use_on_device(&seen)) return;
    prefer_max_shared(stream_offsets_kernel); prefer_max_shared(pack_payload_kernel);
    prefer_max_shared(mailbox_kernel); prefer_max_shared(copy_prefix_kernel);
    prefer_max_shared(first_error_kernel); prefer_max_shared(first_error_finish_kernel);
}

int ensure_mailbox(eae_codec* c)
{
    if (c->mailbox) return 0;
    void* p = nullptr;
    EAE_CUDA_OK(cudaHostAlloc(&p, sizeof(HostMailbox), cudaHostAllocPortable | cudaHostAllocMapped));
    c->mailbox = reinterpret_cast<HostMailbox*>(p);
    return 0;
}

int post_mailbox(eae_codec* c, const uint64_t* total_dev, const eae_batch_stats_t* stats_dev, cudaStream_t st)
{
    EAE_TRY(ensure_mailbox(c));
    void* mb_dev = nullptr;
    EAE_CUDA_OK(cudaHostGetDevicePointer(&mb_dev, c->mailbox, 0));
    mailbox_kernel<<<1, 64, 0, st>>>(reinterpret_cast<HostMailbox*>(mb_dev), total_dev, c->flag.as<uint32_t>(), stats_dev,
                                     c->math != EAE_MATH_FP32_SIMT ? umma_error_flag_dev() : nullptr);
    EAE_LAUNCH_OK();
    return 0;
}

int check_mailbox(const eae_codec* c, const char* what)
{
    const HostMailbox& mb = *c->mailbox;
    if (mb.umma_err) { set_error("tcgen05 GEMM pipeline timed out (role mask 0x%x)", mb.umma_err); return EAE_ERR_CUDA; }
    if (mb.flag[0] & 1u) { set_error("The rounded array elements cannot be represented as 16-bit signed integers."); return EAE_ERR_INT16_RANGE; }
    if (mb.flag[1]) { set_error("Error of type %u during the %s.", mb.flag[1], what); return (int)mb.flag[1]; }
    return 0;
}

// Device-visible alias of a host buffer if it is pinned / registered, else NULL.
uint8_t* device_alias_of_pinned(const void* host)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return nullptr;
    return reinterpret_cast<uint8_t*>(attr.devicePointer);
}

// Work enqueued on the returned stream starts after everything already enqueued on `st`.
int fork_coder_stream(eae_codec* c, cudaStream_t st, cudaStream_t* out)
{
    *out = st;
    if (!c->coder_priority) return 0;
    if (!c->coder_stream) {
        int least = 0, greatest = 0;
        EAE_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        EAE_CUDA_OK(cudaStreamCreateWithPriority(&c->coder_stream, cudaStreamNonBlocking, c->coder_priority == 2 ? least : greatest));
        EAE_CUDA_OK(cudaEventCreateWithFlags(&c->fork_event, cudaEventDisableTiming));
        EAE_CUDA_OK(cudaEventCreateWithFlags(&c->join_event, cudaEventDisableTiming));
    }
    EAE_CUDA_OK(cudaEventRecord(c->fork_event, st));
    EAE_CUDA_OK(cudaStreamWaitEvent(c->coder_stream, c->fork_event, 0));
    *out = c->coder_stream;
    return 0;
}

// Work enqueued on `st` from now on starts after everything enqueued on the coder stream.
int join_coder_stream(eae_codec* c, cudaStream_t st, cudaStream_t cs)
{
    if (cs == st) return 0;
    EAE_CUDA_OK(cudaEventRecord(c->join_event, cs));
    EAE_CUDA_OK(cudaStreamWaitEvent(st, c->join_event, 0));
    return 0;
}

// ---- step graphs ------------------------------------------------------------------------------
bool graphs_usable(const eae_codec* c, cudaStream_t st)
{
    static int timing = -1;
    if (timing < 0) timing = getenv("EAE_UMMA_TIMING") ? 1 : 0;      // the timing experiments synchronise inside the launchers
    return c->use_graphs && !c->host_call && st != nullptr && !profiling_enabled() && !timing && !c->coder_priority;
}

StepGraph* find_step_graph(eae_codec* c, const StepGraphKey& key)
{
    for (StepGraph& g : c->graphs)
        if (memcmp(&g.key, &key, sizeof key) == 0) { g.uses++; g.last_use = ++c->graph_clock; return &g; }
    if (c->graphs.size() >= 16) {      // evict the least recently used entry
        size_t victim = 0;
        for (size_t i = 1; i < c->graphs.size(); i++) if (c->graphs[i].last_use < c->graphs[victim].last_use) victim = i;
        if (c->graphs[victim].exec) cudaGraphExecDestroy(c->graphs[victim].exec);
        c->graphs.erase(c->graphs.begin() + (long)victim);
    }
    StepGraph g;
    memcpy(&g.key, &key, sizeof key);
    g.uses = 1; g.last_use = ++c->graph_clock;
    c->graphs.push_back(g);
    return &c->graphs.back();
}

StepGraphKey make_step_key(const eae_codec* c, int kind, uint32_t n, uint32_t h, uint32_t w, uint32_t L, const void* p0,
                           const void* p1, const void* p2, const void* p3, uint64_t cap)
{
    StepGraphKey key;
    memset(&key, 0, sizeof key);      // (padding bytes too: keys are compared with memcmp)
    key.kind = kind; key.n = n; key.h = h; key.w = w; key.L = L;
    key.p0 = p0; key.p1 = p1; key.p2 = p2; key.p3 = p3; key.cap = cap;
    key.math = c->math; key.lanes = c->coder_lanes; key.generation = c->generation;
    return key;
}

// Runs body(st) either directly or - from the second use of `key` on - captured once and replayed as a graph.
template <typename Body>
int run_as_step_graph(eae_codec* c, const StepGraphKey& key, cudaStream_t st, Body body)
{
    if (!graphs_usable(c, st)) return body(st);
    StepGraph* g = find_step_graph(c, key);
    if (g->exec) {
        EAE_CUDA_OK(cudaGraphLaunch(g->exec, st));
        count_launch(g->launches);
        return 0;
    }
    if (g->uses < 2) return body(st);      // one-off calls are not worth a capture
    const uint64_t before = eae_launch_count();
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->use_graphs = 0;
        return body(st);
    }
    const int rc = body(st);
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
    cudaGraphExec_t exec = nullptr;
    cudaError_t e_inst = cudaErrorUnknown;
    if (rc == 0 && e_end == cudaSuccess && graph) e_inst = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != 0) { cudaGetLastError(); return rc; }
    if (e_inst != cudaSuccess) {           // nothing has run: fall back to plain launches, for good
        cudaGetLastError();
        c->use_graphs = 0;
        return body(st);
    }
    g = find_step_graph(c, key);           // (the vector did not change; keeps the pointer honest)
    g->uses--;
    g->exec = exec;
    g->launches = (int)(eae_launch_count() - before);
    EAE_CUDA_OK(cudaGraphLaunch(exec, st));
    return 0;
}

int compress_dev_impl(eae_codec* c, const eae_coding_params_t* prm, const uint8_t* img_dev, uint32_t n,
                      uint32_t h, uint32_t w, uint8_t* container_dev, uint64_t cap, uint64_t* total_dev,
                      eae_batch_stats_t* stats_dev, cudaStream_t st)
{
    EAE_TRY(check_dims(n, h, w));
    codec_carveouts();
    EAE_TRY(upload_params(c, prm, st));
    const uint32_t L = prm->truncated_unary_length;
    const uint32_t hw3 = (h / 16) * (w / 16);
    const uint32_t n_streams = n * EAE_NB_MAPS;
    const uint32_t chunk = chunk_images(h, w);
    EAE_TRY(ensure_workspace(c, n < chunk ? n : chunk, h, w));
    EAE_TRY(ensure_coder(c, n_streams, hw3, L));
    if (cap < kHeaderBytes + 8ull * n_streams) { set_error("container capacity too small"); return EAE_ERR_ARGUMENT; }
    c->no_phase_merge = c->phase_merge < 0 ? !(n <= chunk && graphs_usable(c, st)) : !c->phase_merge;
    c->last_idx_elems = (uint64_t)n_streams * hw3;      // (host state: outside the body, which may be replayed as a graph)
    c->last_n_streams = n_streams;
    // parts: as encode_chunk's (1 = layer 1 has already been launched)
    auto body = [&](cudaStream_t st, int parts) -> int {
    EAE_CUDA_OK(cudaMemsetAsync(c->flag.p, 0, 8, st));
    for (uint32_t i0 = 0; i0 < n; i0 += chunk) {
        const uint32_t nc = n - i0 < chunk ? n - i0 : chunk;
        float* y = c->buf3.as<float>();
        int16_t* idx = c->idx_planar.as<int16_t>() + (size_t)i0 * EAE_NB_MAPS * hw3;
        if (quantizer_fusable(c, h)) {
            // layer 3 (+ GDN) writes the indices itself: no fp32 latent, no quantizer launch
            const QuantOut quant{idx, c->mean.as<float>(), c->delta.as<float>(), c->flag.as<uint32_t>()};
            EAE_TRY(encode_chunk(c, img_dev + (size_t)i0 * h * w, nc, h, w, y, st, parts, &quant));
            continue;
        }
        EAE_TRY(encode_chunk(c, img_dev + (size_t)i0 * h * w, nc, h, w, y, st, parts));
        ProfScope prof(kProfQuantize, st);
        EAE_TRY(launch_quantize_to_planar(y, c->mean.as<float>(), c->delta.as<float>(), idx, nullptr,
                                          nc, hw3, c->flag.as<uint32_t>(), st));
    }
    cudaStream_t cs = st;
    EAE_TRY(fork_coder_stream(c, st, &cs));
    {
        ProfScope prof(kProfCoderEncode, cs);
        EAE_TRY(launch_encode_streams(c->idx_planar.as<int16_t>(), n_streams, hw3, c->table.as<double>(), EAE_NB_MAPS, L,
                                      nullptr, c->bac_slots.as<uint8_t>(), c->byp_slots.as<uint8_t>(), c->cw_slot,
                                      c->bac_bits.as<uint32_t>(), c->byp_bits.as<uint32_t>(), c->err.as<uint32_t>(), cs,
                                      c->coder_lanes, c->enc_scratch.p, c->qtable.as<uint64_t>(),
                                      c->row_flags.as<uint8_t>()));
    }
    {
    ProfScope prof_pack(kProfPack, cs);
    eae_batch_stats_t* sd = stats_dev ? stats_dev : c->stats.as<eae_batch_stats_t>();
    const OffsetsExtra extra{reinterpret_cast<uint32_t*>(container_dev), n, h, w, L, sd, c->err.as<uint32_t>(),
                             c->flag.as<uint32_t>(), nullptr, nullptr, c->stats_acc, cap, 0};
    stream_offsets_kernel<<<1, 1024, 0, cs>>>(c->bac_bits.as<uint32_t>(), c->byp_bits.as<uint32_t>(), 1, n_streams,
                                              kHeaderBytes + 8ull * n_streams, c->bac_off.as<uint64_t>(),
                                              c->byp_off.as<uint64_t>(), total_dev, extra);
    EAE_LAUNCH_OK();
    pack_payload_kernel<<<ceil_div_u32((uint64_t)n_streams * 32, 256), 256, 0, cs>>>(
        container_dev, cap, c->bac_slots.as<uint8_t>(), c->byp_slots.as<uint8_t>(), c->cw_slot,
        c->bac_bits.as<uint32_t>(), c->byp_bits.as<uint32_t>(), c->bac_off.as<uint64_t>(),
        c->byp_off.as<uint64_t>(), n_streams);
    EAE_LAUNCH_OK();
    }
    return join_coder_stream(c, st, cs);
    };
    if (n <= chunk && graphs_usable(c, st)) {
        // layer 1 is the only reader of the caller's images: launched as it is; the rest of the step as a graph
        EAE_TRY(encode_chunk(c, img_dev, n, h, w, c->buf3.as<float>(), st, 1));
        return run_as_step_graph(c, make_step_key(c, 0, n, h, w, L, container_dev, total_dev, stats_dev, nullptr, cap), st,
                                 [&](cudaStream_t s) { return body(s, 2); });
    }
    return body(st, 3);
}

// nbytes: readable bytes at container_dev (the exact size of the container or the capacity of its buffer)
int decompress_dev_impl(eae_codec* c, const eae_coding_params_t* prm, const uint8_t* container_dev, uint64_t nbytes,
                        uint32_t n, uint32_t h, uint32_t w, uint8_t* rec_dev, cudaStream_t st)
{
    EAE_TRY(check_dims(n, h, w));
    codec_carveouts();
    EAE_TRY(upload_params(c, prm, st));
    const uint32_t L = prm->truncated_unary_length;
    const uint32_t hw3 = (h / 16) * (w / 16);
    const uint32_t n_streams = n * EAE_NB_MAPS;
    const uint32_t chunk = chunk_images(h, w);
    EAE_TRY(ensure_workspace(c, n < chunk ? n : chunk, h, w));
    EAE_TRY(ensure_coder(c, n_streams, hw3, L));
    c->no_phase_merge = c->phase_merge < 0 ? !(n <= chunk && graphs_usable(c, st)) : !c->phase_merge;
    c->last_idx_elems = (uint64_t)n_streams * hw3;
    c->last_n_streams = n_streams;
    auto body = [&](cudaStream_t st) -> int {
    EAE_CUDA_OK(cudaMemsetAsync(c->flag.p, 0, 8, st));
    cudaStream_t cs = st;
    EAE_TRY(fork_coder_stream(c, st, &cs));
    const uint32_t* tbl = reinterpret_cast<const uint32_t*>(container_dev + kHeaderBytes);
    // payload offsets, and the stream table de-interleaved into the bit-count arrays the decoder reads
    const OffsetsExtra extra{nullptr, 0, 0, 0, 0, nullptr, nullptr, c->flag.as<uint32_t>(), c->bac_bits.as<uint32_t>(),
                             c->byp_bits.as<uint32_t>(), nullptr, nbytes, eae_coder_capacity_bytes(hw3, L) * 8u};
    stream_offsets_kernel<<<1, 1024, 0, cs>>>(tbl, tbl + 1, 2, n_streams, kHeaderBytes + 8ull * n_streams,
                                              c->bac_off.as<uint64_t>(), c->byp_off.as<uint64_t>(),
                                              c->total_bytes.as<uint64_t>(), extra);
    EAE_LAUNCH_OK();
    {
        ProfScope prof_dec(kProfCoderDecode, cs);
        EAE_TRY(launch_decode_streams(c->idx_planar.as<int16_t>(), n_streams, hw3, c->table.as<double>(), EAE_NB_MAPS, L,
                                      nullptr, container_dev, c->bac_off.as<uint64_t>(), c->bac_bits.as<uint32_t>(),
                                      container_dev, c->byp_off.as<uint64_t>(), c->byp_bits.as<uint32_t>(),
                                      c->err.as<uint32_t>(), cs, c->coder_lanes, c->qtable.as<uint64_t>(),
                                      c->row_flags.as<uint8_t>()));
    }
    EAE_TRY(join_coder_stream(c, st, cs));
    for (uint32_t i0 = 0; i0 < n; i0 += chunk) {
        const uint32_t nc = n - i0 < chunk ? n - i0 : chunk;
        // the dequantized latent lives in bufA's tail-free region: use buf3 when IGDN4 is absent,
        // otherwise a separate buffer is needed because IGDN4 writes buf3.
        float* q = c->learned ? c->buf3.as<float>() : c->bufA.as<float>();
        if (dequantizer_fusable(c)) {
            // IGDN4 reads the indices itself: no fp32 latent, no dequantizer launch
            const DequantIn dq{c->idx_planar.as<int16_t>() + (size_t)i0 * EAE_NB_MAPS * hw3, c->mean.as<float>(),
                               c->delta.as<float>(), (int)hw3};
            EAE_TRY(decode_chunk(c, nullptr, nc, h, w, rec_dev + (size_t)i0 * h * w, nullptr, st, &dq));
            continue;
        }
        {
            ProfScope prof(kProfDequantize, st);
            EAE_TRY(launch_dequantize_from_planar(c->idx_planar.as<int16_t>() + (size_t)i0 * EAE_NB_MAPS * hw3,
                                                  c->mean.as<float>(), c->delta.as<float>(), q, nc, hw3, st));
        }
        EAE_TRY(decode_chunk(c, q, nc, h, w, rec_dev + (size_t)i0 * h * w, nullptr, st));
    }
    return 0;
    };
    if (n <= chunk) return run_as_step_graph(c, make_step_key(c, 1, n, h, w, L, container_dev, rec_dev, nullptr, nullptr, nbytes), st, body);
    return body(st);
}

}  // namespace

// =================================================================================================
extern "C" int eae_codec_create(eae_codec_t** out, const eae_weights_t* wt, int learned, int device)
{
    if (!out || !wt) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    const float* need[] = {wt->weights_1, wt->biases_1, wt->gamma_1, wt->beta_1, wt->weights_2, wt->biases_2,
                           wt->gamma_2, wt->beta_2, wt->weights_3, wt->biases_3, wt->weights_4, wt->biases_4,
                           wt->gamma_5, wt->beta_5, wt->weights_5, wt->biases_5, wt->gamma_6, wt->beta_6,
                           wt->weights_6};
    for (const float* p : need) if (!p) { set_error("a required weight pointer is NULL"); return EAE_ERR_NULL; }
    if (!learned && (!wt->gamma_3 || !wt->beta_3 || !wt->gamma_4 || !wt->beta_4)) {
        set_error("gamma_3/beta_3/gamma_4/beta_4 are required when the bin widths are not learned");
        return EAE_ERR_NULL;
    }
    EAE_TRY(require_device());
    EAE_CUDA_OK(cudaSetDevice(device));
    std::unique_ptr<eae_codec> c(new eae_codec);
    c->device = device;
    c->learned = learned ? 1 : 0;

    // weights_1 [9,9,1,128] -> [96][128]
    {
        std::vector<float> m((size_t)kIm2colK * 128, 0.f);
        memcpy(m.data(), wt->weights_1, (size_t)81 * 128 * 4);
        EAE_TRY(upload(c->w1m, m.data(), m.size()));
    }
    EAE_TRY(upload(c->w2, wt->weights_2, (size_t)25 * 128 * 128));
    EAE_TRY(upload(c->w3, wt->weights_3, (size_t)25 * 128 * 128));
    // conv2d_transpose filters [kh,kw,out,in] -> per tap [in][out]
    auto transpose_taps = [](const float* src, std::vector<float>& dst) {
        dst.resize((size_t)25 * 128 * 128);
        for (int t = 0; t < 25; t++)
            for (int o = 0; o < 128; o++)
                for (int i = 0; i < 128; i++)
                    dst[((size_t)t * 128 + i) * 128 + o] = src[((size_t)t * 128 + o) * 128 + i];
    };
    {
        std::vector<float> t4, t5;
        transpose_taps(wt->weights_4, t4);
        transpose_taps(wt->weights_5, t5);
        EAE_TRY(upload(c->w4, t4.data(), t4.size()));
        EAE_TRY(upload(c->w5, t5.data(), t5.size()));
    }
    // weights_6 [9,9,1(out),128(in)] -> [128 in][128 cols], column = ky*9+kx
    {
        std::vector<float> m((size_t)128 * 128, 0.f);
        for (int t = 0; t < 81; t++)
            for (int i = 0; i < 128; i++) m[(size_t)i * 128 + t] = wt->weights_6[(size_t)t * 128 + i];
        EAE_TRY(upload(c->w6m, m.data(), m.size()));
    }
    const float* gammas[6] = {wt->gamma_1, wt->gamma_2, wt->gamma_3, wt->gamma_4, wt->gamma_5, wt->gamma_6};
    const float* betas[6] = {wt->beta_1, wt->beta_2, wt->beta_3, wt->beta_4, wt->beta_5, wt->beta_6};
    for (int i = 0; i < 6; i++) {
        if (!gammas[i]) continue;
        EAE_TRY(upload(c->gamma[i], gammas[i], (size_t)128 * 128));
        EAE_TRY(upload(c->beta[i], betas[i], 128));
    }
    const float* biases[5] = {wt->biases_1, wt->biases_2, wt->biases_3, wt->biases_4, wt->biases_5};
    for (int i = 0; i < 5; i++) EAE_TRY(upload(c->bias[i], biases[i], 128));

    // K-major, tf32-split copies for the tensor path: [taps][128 out][Cin]
    {
        std::vector<float> k;
        // conv1: [128 out][96]  (TF [81][1][128 out] transposed, zero-padded)
        k.assign((size_t)128 * kIm2colK, 0.f);
        for (int t = 0; t < 81; t++)
            for (int o = 0; o < 128; o++) k[(size_t)o * kIm2colK + t] = wt->weights_1[(size_t)t * 128 + o];
        EAE_TRY(upload_split(c->wk_hi[0], c->wk_lo[0], k));
        // conv2 / conv3: TF [tap][in][out] -> [tap][out][in]
        const float* convs[2] = {wt->weights_2, wt->weights_3};
        for (int l = 0; l < 2; l++) {
            k.assign((size_t)25 * 128 * 128, 0.f);
            for (int t = 0; t < 25; t++)
                for (int i = 0; i < 128; i++)
                    for (int o = 0; o < 128; o++)
                        k[((size_t)t * 128 + o) * 128 + i] = convs[l][((size_t)t * 128 + i) * 128 + o];
            EAE_TRY(upload_split(c->wk_hi[1 + l], c->wk_lo[1 + l], k));
        }
        // transposed convs: TF [tap][out][in] is already K-major
        k.assign(wt->weights_4, wt->weights_4 + (size_t)25 * 128 * 128);
        EAE_TRY(upload_split(c->wk_hi[3], c->wk_lo[3], k));
        k.assign(wt->weights_5, wt->weights_5 + (size_t)25 * 128 * 128);
        EAE_TRY(upload_split(c->wk_hi[4], c->wk_lo[4], k));
        // tconv3: rows = the 81 filter taps (zero-padded to 128), K = 128 input channels: TF [81][1][128 in]
        k.assign((size_t)128 * 128, 0.f);
        memcpy(k.data(), wt->weights_6, (size_t)81 * 128 * 4);
        EAE_TRY(upload_split(c->wk_hi[5], c->wk_lo[5], k));
        // gamma[j in][i out] -> [i][j]
        for (int g = 0; g < 6; g++) {
            if (!gammas[g]) continue;
            k.assign((size_t)128 * 128, 0.f);
            for (int j = 0; j < 128; j++)
                for (int i = 0; i < 128; i++) k[(size_t)i * 128 + j] = gammas[g][(size_t)j * 128 + i];
            EAE_TRY(upload_split(c->gk_hi[g], c->gk_lo[g], k));
        }
    }
    if (const char* env = getenv("EAE_UMMA_LAYERS")) c->umma_mask = atoi(env);
    if (const char* env = getenv("EAE_NO_FUSE")) c->no_fuse = atoi(env);
    if (const char* env = getenv("EAE_NO_DIRECT_CONV1")) c->no_direct_conv1 = atoi(env);
    if (const char* env = getenv("EAE_NO_FUSE_QUANT")) c->no_fuse_quant = atoi(env);
    if (const char* env = getenv("EAE_GDN_PRECISE")) c->gdn_precise = atoi(env);
    if (const char* env = getenv("EAE_PHASE_MERGE")) c->phase_merge = atoi(env);
    if (const char* env = getenv("EAE_CODER_PRIORITY")) c->coder_priority = atoi(env);
    if (const char* env = getenv("EAE_GRAPHS")) c->use_graphs = atoi(env);
    if (const char* env = getenv("EAE_GRAPHS_HOST")) c->graphs_in_host_calls = atoi(env);
    *out = c.release();
    return 0;
}

extern "C" int eae_codec_destroy(eae_codec_t* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->mailbox) cudaFreeHost(c->mailbox);
    if (c->status_host) cudaFreeHost(c->status_host);
    for (StepGraph& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (c->coder_stream) { cudaStreamSynchronize(c->coder_stream); cudaStreamDestroy(c->coder_stream); }
    if (c->fork_event) cudaEventDestroy(c->fork_event);
    if (c->join_event) cudaEventDestroy(c->join_event);
    delete c;
    return 0;
}

extern "C" int eae_codec_set_math(eae_codec_t* c, int mode)
{
    if (!c) { set_error("NULL codec"); return EAE_ERR_NULL; }
    if (mode != EAE_MATH_FP32_SIMT && mode != EAE_MATH_TF32X3 && mode != EAE_MATH_TF32 && mode != EAE_MATH_MIXED) {
        set_error("unknown math mode %d", mode); return EAE_ERR_ARGUMENT;
    }
    if (mode != EAE_MATH_FP32_SIMT) EAE_TRY(umma_available());
    c->math = mode;
    return 0;
}

extern "C" int eae_codec_set_coder_lanes(eae_codec_t* c, uint32_t lanes)
{
    if (!c) { set_error("NULL codec"); return EAE_ERR_NULL; }
    if (lanes > 32 || (lanes & (lanes - 1)) != 0) { set_error("coder lanes must be 0 or a power of two <= 32"); return EAE_ERR_ARGUMENT; }
    c->coder_lanes = lanes;
    return 0;
}

extern "C" int eae_codec_get_math(const eae_codec_t* c) { return c ? c->math : EAE_ERR_NULL; }

extern "C" int eae_encode_dev(eae_codec_t* c, const uint8_t* img_dev, uint32_t n, uint32_t h, uint32_t w,
                              float* y_dev, void* stream)
{
    if (c) c->no_phase_merge = c->phase_merge == 1 ? 0 : 1;      // launched kernel by kernel
    if (!c || !img_dev || !y_dev) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_dims(n, h, w));
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t chunk = chunk_images(h, w);
    EAE_TRY(ensure_workspace(c, n < chunk ? n : chunk, h, w));
    const size_t ypi = (size_t)(h / 16) * (w / 16) * 128;
    for (uint32_t i0 = 0; i0 < n; i0 += chunk) {
        const uint32_t nc = n - i0 < chunk ? n - i0 : chunk;
        EAE_TRY(encode_chunk(c, img_dev + (size_t)i0 * h * w, nc, h, w, y_dev + i0 * ypi, st));
    }
    return 0;
}

extern "C" int eae_decode_dev(eae_codec_t* c, const float* q_dev, uint32_t n, uint32_t h, uint32_t w,
                              uint8_t* rec_dev, void* stream)
{
    if (c) c->no_phase_merge = c->phase_merge == 1 ? 0 : 1;      // launched kernel by kernel
    if (!c || !q_dev || !rec_dev) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_dims(n, h, w));
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t chunk = chunk_images(h, w);
    EAE_TRY(ensure_workspace(c, n < chunk ? n : chunk, h, w));
    const size_t ypi = (size_t)(h / 16) * (w / 16) * 128;
    for (uint32_t i0 = 0; i0 < n; i0 += chunk) {
        const uint32_t nc = n - i0 < chunk ? n - i0 : chunk;
        EAE_TRY(decode_chunk(c, q_dev + i0 * ypi, nc, h, w, rec_dev + (size_t)i0 * h * w, nullptr, st));
    }
    return 0;
}

extern "C" int eae_encode_host(eae_codec_t* c, const uint8_t* img, uint32_t n, uint32_t h, uint32_t w,
                               float* y_out, void* stream)
{
    if (!c || !img || !y_out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_dims(n, h, w));
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf di, dy;
    const size_t nin = (size_t)n * h * w, ny = (size_t)n * (h / 16) * (w / 16) * 128;
    EAE_TRY(di.alloc(nin));
    EAE_TRY(dy.alloc(ny * 4));
    EAE_CUDA_OK(cudaMemcpyAsync(di.p, img, nin, cudaMemcpyHostToDevice, st));
    EAE_TRY(eae_encode_dev(c, di.as<uint8_t>(), n, h, w, dy.as<float>(), stream));
    EAE_CUDA_OK(cudaMemcpyAsync(y_out, dy.p, ny * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (c->math != EAE_MATH_FP32_SIMT) EAE_TRY(umma_check_error(st));
    return 0;
}

extern "C" int eae_decode_host(eae_codec_t* c, const float* q, uint32_t n, uint32_t h, uint32_t w,
                               uint8_t* rec_out, void* stream)
{
    if (!c || !q || !rec_out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_dims(n, h, w));
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf dq, dr;
    const size_t nout = (size_t)n * h * w, nq = (size_t)n * (h / 16) * (w / 16) * 128;
    EAE_TRY(dq.alloc(nq * 4));
    EAE_TRY(dr.alloc(nout));
    EAE_CUDA_OK(cudaMemcpyAsync(dq.p, q, nq * 4, cudaMemcpyHostToDevice, st));
    EAE_TRY(eae_decode_dev(c, dq.as<float>(), n, h, w, dr.as<uint8_t>(), stream));
    EAE_CUDA_OK(cudaMemcpyAsync(rec_out, dr.p, nout, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (c->math != EAE_MATH_FP32_SIMT) EAE_TRY(umma_check_error(st));
    return 0;
}

extern "C" int eae_decode_float_host(eae_codec_t* c, const float* q, uint32_t n, uint32_t h, uint32_t w,
                                     float* rec_out, void* stream)
{
    if (c) c->no_phase_merge = c->phase_merge == 1 ? 0 : 1;      // launched kernel by kernel
    if (!c || !q || !rec_out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_dims(n, h, w));
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf dq, dr;
    const size_t nout = (size_t)n * h * w, per_q = (size_t)(h / 16) * (w / 16) * 128;
    EAE_TRY(dq.alloc((size_t)n * per_q * 4));
    EAE_TRY(dr.alloc(nout * 4));
    EAE_CUDA_OK(cudaMemcpyAsync(dq.p, q, (size_t)n * per_q * 4, cudaMemcpyHostToDevice, st));
    const uint32_t chunk = chunk_images(h, w);
    EAE_TRY(ensure_workspace(c, n < chunk ? n : chunk, h, w));
    for (uint32_t i0 = 0; i0 < n; i0 += chunk) {
        const uint32_t nc = n - i0 < chunk ? n - i0 : chunk;
        EAE_TRY(decode_chunk(c, dq.as<float>() + i0 * per_q, nc, h, w, nullptr, dr.as<float>() + (size_t)i0 * h * w, st));
    }
    EAE_CUDA_OK(cudaMemcpyAsync(rec_out, dr.p, nout * 4, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    if (c->math != EAE_MATH_FP32_SIMT) EAE_TRY(umma_check_error(st));
    return 0;
}

extern "C" uint64_t eae_container_bound(uint32_t n, uint32_t h, uint32_t w, uint32_t L)
{
    const uint64_t size = (uint64_t)(h / 16) * (w / 16);
    const uint64_t n_streams = (uint64_t)n * EAE_NB_MAPS;
    return kHeaderBytes + 8 * n_streams + 2 * n_streams * (uint64_t)eae_coder_capacity_bytes((uint32_t)size, L);
}

extern "C" int eae_compress_dev(eae_codec_t* c, const eae_coding_params_t* prm, const uint8_t* img_dev,
                                uint32_t n, uint32_t h, uint32_t w, uint8_t* container_dev, uint64_t cap,
                                uint64_t* total_dev, eae_batch_stats_t* stats_dev, void* stream)
{
    if (!c || !img_dev || !container_dev || !total_dev) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_CUDA_OK(cudaSetDevice(c->device));
    c->last_stream = (cudaStream_t)stream;
    return compress_dev_impl(c, prm, img_dev, n, h, w, container_dev, cap, total_dev, stats_dev,
                             (cudaStream_t)stream);
}

extern "C" int eae_decompress_dev(eae_codec_t* c, const eae_coding_params_t* prm, const uint8_t* container_dev,
                                  uint64_t container_bytes, uint32_t n, uint32_t h, uint32_t w, uint8_t* rec_dev,
                                  void* stream)
{
    if (!c || !container_dev || !rec_dev) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (container_bytes < kHeaderBytes + 8ull * n * EAE_NB_MAPS) { set_error("container shorter than its stream table"); return EAE_ERR_ARGUMENT; }
    EAE_CUDA_OK(cudaSetDevice(c->device));
    c->last_stream = (cudaStream_t)stream;
    return decompress_dev_impl(c, prm, container_dev, container_bytes, n, h, w, rec_dev, (cudaStream_t)stream);
}

extern "C" int eae_codec_set_stats_accumulator(eae_codec_t* c, eae_batch_stats_t* acc_dev)
{
    if (!c) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    c->stats_acc = acc_dev;
    c->generation++;        // (the pointer is baked into captured steps)
    return 0;
}

// Status of the device-resident entry points since the last poll (they cannot report device-side failures themselves).
namespace {
struct HostStatus { uint32_t sticky, first_err, decode_err, umma_err; uint64_t total; };

__global__ void __launch_bounds__(1024)
status_kernel(HostStatus* __restrict__ out, uint32_t* __restrict__ flag, const uint32_t* __restrict__ err, uint32_t n_streams,
              const uint64_t* __restrict__ total, uint32_t* __restrict__ umma_flag)
{
    __shared__ uint32_t key;
    if (threadIdx.x == 0) key = 0xFFFFFFFFu;
    __syncthreads();
    // error of the lowest-numbered stream of the last step (the decoder leaves its codes only there)
    uint32_t mine = 0xFFFFFFFFu;
    for (uint32_t s = threadIdx.x; s < n_streams; s += blockDim.x)
        if (err[s] && mine == 0xFFFFFFFFu) mine = (s << 8) | (err[s] & 0xFFu);
    if (mine != 0xFFFFFFFFu) atomicMin(&key, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        out->sticky = atomicExch(flag + 2, 0u);
        out->first_err = atomicExch(flag + 3, 0u);
        out->decode_err = key == 0xFFFFFFFFu ? 0u : (key & 0xFFu);
        out->umma_err = umma_flag ? atomicExch(umma_flag, 0u) : 0u;
        out->total = total ? *total : 0ull;
        __threadfence_system();
    }
}
}  // namespace

extern "C" int eae_codec_poll_status(eae_codec_t* c, void* stream, eae_codec_status_t* out)
{
    if (!c) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (out) memset(out, 0, sizeof *out);
    if (!c->flag.p || !c->err.p) return 0;      // nothing has run on this codec yet
    if (!c->status_host) {
        void* p = nullptr;
        EAE_CUDA_OK(cudaHostAlloc(&p, sizeof(HostStatus), cudaHostAllocPortable | cudaHostAllocMapped));
        c->status_host = p;
    }
    void* dev = nullptr;
    EAE_CUDA_OK(cudaHostGetDevicePointer(&dev, c->status_host, 0));
    status_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<HostStatus*>(dev), c->flag.as<uint32_t>(), c->err.as<uint32_t>(),
                                      c->last_n_streams, nullptr,
                                      c->math != EAE_MATH_FP32_SIMT ? umma_error_flag_dev() : nullptr);
    EAE_LAUNCH_OK();
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    const HostStatus hs = *reinterpret_cast<const HostStatus*>(c->status_host);
    const uint32_t coder = hs.first_err ? hs.first_err : hs.decode_err;
    if (out) {
        out->int16_overflow = (hs.sticky & kStatusInt16) ? 1u : 0u;
        out->container_overflow = (hs.sticky & kStatusNoRoom) ? 1u : 0u;
        out->container_invalid = (hs.sticky & (kStatusBadTable | kStatusTruncated)) ? 1u : 0u;
        out->coder_error = coder;
        out->tensor_timeout_mask = hs.umma_err;
    }
    if (hs.umma_err) { set_error("tcgen05 GEMM pipeline timed out (role mask 0x%x)", hs.umma_err); return EAE_ERR_CUDA; }
    if (hs.sticky & kStatusInt16) { set_error("The rounded array elements cannot be represented as 16-bit signed integers."); return EAE_ERR_INT16_RANGE; }
    if (hs.sticky & kStatusNoRoom) { set_error("a container did not fit the capacity given to eae_compress_dev"); return EAE_ERR_ARGUMENT; }
    if (hs.sticky & kStatusBadTable) { set_error("a container's stream table exceeds the coder capacity"); return EAE_ERR_CAPACITY; }
    if (hs.sticky & kStatusTruncated) { set_error("a container is shorter than its stream table says"); return EAE_ERR_RESOURCE; }
    if (coder) { set_error("Error of type %u during the coding.", coder); return (int)coder; }
    return 0;
}

extern "C" int eae_compress_host(eae_codec_t* c, const eae_coding_params_t* prm, const uint8_t* img, uint32_t n,
                                 uint32_t h, uint32_t w, uint8_t* container, uint64_t cap, uint64_t* out_bytes,
                                 eae_batch_stats_t* stats, void* stream)
{
    if (!c || !img || !container || !out_bytes) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    EAE_TRY(check_dims(n, h, w));
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nin = (size_t)n * h * w;
    if (c->img_u8.bytes < nin) EAE_TRY(c->img_u8.alloc(nin));
    // The size of the container is only known after coding. It is assembled in device memory (sized by the caller's
    // capacity, bounded by the worst case); when the caller's buffer is pinned host memory a kernel that reads the
    // size on the device copies it there in wide writes (one synchronisation per call), otherwise the host copies
    // it once it knows the size.
    uint64_t bound = eae_container_bound(n, h, w, prm ? prm->truncated_unary_length : 1);
    uint64_t dcap = cap < bound ? cap : bound;
    uint8_t* direct = device_alias_of_pinned(container);
    if (direct && ((reinterpret_cast<uintptr_t>(direct) & 15u) || (cap & 15u))) direct = nullptr;     // needs 16-byte granularity
    if (c->rec_u8.bytes < dcap + 16) EAE_TRY(c->rec_u8.alloc(dcap + 16));
    EAE_CUDA_OK(cudaMemcpyAsync(c->img_u8.p, img, nin, cudaMemcpyHostToDevice, st));
    if (!c->total_bytes.p) EAE_TRY(c->total_bytes.alloc(8));
    c->host_call = !c->graphs_in_host_calls;
    struct Reset { eae_codec* c; ~Reset() { c->host_call = false; } } reset{c};
    c->last_stream = st;
    EAE_TRY(compress_dev_impl(c, prm, c->img_u8.as<uint8_t>(), n, h, w, c->rec_u8.as<uint8_t>(), dcap,
                              c->total_bytes.as<uint64_t>(), nullptr, st));
    if (direct) {
        copy_prefix_kernel<<<148, 256, 0, st>>>(direct, c->rec_u8.as<uint8_t>(), c->total_bytes.as<uint64_t>(), dcap);
        EAE_LAUNCH_OK();
    }
    EAE_TRY(post_mailbox(c, c->total_bytes.as<uint64_t>(), c->stats.as<eae_batch_stats_t>(), st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    EAE_TRY(check_mailbox(c, "encoding"));
    const uint64_t total = c->mailbox->total;
    if (total > dcap) { set_error("container needs %llu bytes, capacity is %llu", (unsigned long long)total, (unsigned long long)cap); return EAE_ERR_ARGUMENT; }
    if (!direct) {
        EAE_CUDA_OK(cudaMemcpyAsync(container, c->rec_u8.p, total, cudaMemcpyDeviceToHost, st));
        EAE_CUDA_OK(cudaStreamSynchronize(st));
    }
    *out_bytes = total;
    if (stats) *stats = c->mailbox->stats;
    return 0;
}

extern "C" int eae_decompress_host(eae_codec_t* c, const eae_coding_params_t* prm, const uint8_t* container,
                                   uint64_t nbytes, uint8_t* rec, uint64_t rec_cap, void* stream)
{
    if (!c || !container || !rec) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (nbytes < kHeaderBytes) { set_error("container shorter than its header"); return EAE_ERR_ARGUMENT; }
    uint32_t hdr[8];
    memcpy(hdr, container, sizeof hdr);
    if (hdr[0] != kMagic || hdr[1] != kVersion || hdr[5] != EAE_NB_MAPS) { set_error("not an EAEB v1 container"); return EAE_ERR_ARGUMENT; }
    const uint32_t n = hdr[2], h = hdr[3], w = hdr[4];
    EAE_TRY(check_dims(n, h, w));
    if (prm && prm->truncated_unary_length != hdr[6]) { set_error("container was coded with L = %u", hdr[6]); return EAE_ERR_ARGUMENT; }
    const uint64_t n_streams = (uint64_t)n * EAE_NB_MAPS;
    if (nbytes < kHeaderBytes + 8 * n_streams) { set_error("container shorter than its stream table"); return EAE_ERR_ARGUMENT; }
    // Validate the payload size on the host before trusting offsets on the device.
    uint64_t need = kHeaderBytes + 8 * n_streams;
    const uint32_t cap_bits = eae_coder_capacity_bytes((h / 16) * (w / 16), hdr[6]) * 8;
    for (uint64_t s = 0; s < n_streams; s++) {
        uint32_t bb, rb;
        memcpy(&bb, container + kHeaderBytes + 8 * s, 4);
        memcpy(&rb, container + kHeaderBytes + 8 * s + 4, 4);
        if (bb > cap_bits || rb > cap_bits) { set_error("stream %llu exceeds the coder capacity", (unsigned long long)s); return EAE_ERR_CAPACITY; }
        need += ((uint64_t)bb + 7) / 8 + ((uint64_t)rb + 7) / 8;
    }
    if (need > nbytes) { set_error("container truncated: needs %llu bytes", (unsigned long long)need); return EAE_ERR_RESOURCE; }
    if ((uint64_t)n * h * w > rec_cap) { set_error("reconstruction buffer too small"); return EAE_ERR_ARGUMENT; }
    EAE_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (c->rec_u8.bytes < nbytes + 16) EAE_TRY(c->rec_u8.alloc(nbytes + 16));
    const size_t nout = (size_t)n * h * w;
    if (c->img_u8.bytes < nout) EAE_TRY(c->img_u8.alloc(nout));
    EAE_CUDA_OK(cudaMemcpyAsync(c->rec_u8.p, container, nbytes, cudaMemcpyHostToDevice, st));
    c->host_call = !c->graphs_in_host_calls;
    struct Reset { eae_codec* c; ~Reset() { c->host_call = false; } } reset{c};
    c->last_stream = st;
    EAE_TRY(decompress_dev_impl(c, prm, c->rec_u8.as<uint8_t>(), nbytes, n, h, w, c->img_u8.as<uint8_t>(), st));
    // error of the lowest-numbered failing stream -> flag[1] -> mailbox; the reconstruction rides the same stream
    EAE_CUDA_OK(cudaMemsetAsync(c->flag.as<uint32_t>() + 1, 0xFF, 4, st));
    first_error_kernel<<<ceil_div_u32(n_streams, 256), 256, 0, st>>>(c->err.as<uint32_t>(), (uint32_t)n_streams,
                                                                    c->flag.as<uint32_t>() + 1);
    EAE_LAUNCH_OK();
    first_error_finish_kernel<<<1, 1, 0, st>>>(c->flag.as<uint32_t>() + 1, c->flag.as<uint32_t>() + 1);
    EAE_LAUNCH_OK();
    EAE_TRY(post_mailbox(c, nullptr, nullptr, st));
    EAE_CUDA_OK(cudaMemcpyAsync(rec, c->img_u8.p, nout, cudaMemcpyDeviceToHost, st));
    EAE_CUDA_OK(cudaStreamSynchronize(st));
    EAE_TRY(check_mailbox(c, "decoding"));
    return 0;
}

extern "C" int eae_last_indices_host(eae_codec_t* c, int16_t* out, uint64_t n_elems)
{
    if (!c || !out) { set_error("NULL pointer"); return EAE_ERR_NULL; }
    if (n_elems > c->last_idx_elems) { set_error("only %llu indices available", (unsigned long long)c->last_idx_elems); return EAE_ERR_ARGUMENT; }
    // on the stream of the step that wrote them (a non-blocking stream does not synchronise with the legacy one)
    EAE_CUDA_OK(cudaSetDevice(c->device));
    EAE_CUDA_OK(cudaMemcpyAsync(out, c->idx_planar.p, n_elems * 2, cudaMemcpyDeviceToHost, c->last_stream));
    EAE_CUDA_OK(cudaStreamSynchronize(c->last_stream));
    return 0;
}
