// Host-side orchestration of one teacher-forced pass (forward + loss, backward) and the C ABI.
//
// Restates Tacotron2NV.forward (models/tacotron2nv.py:81-127) and Tacotron2Loss
// (modules_tacotron2nv/tacotron2nv_loss.py:17-52) and derives their backward by hand
//  as: batched GEMMs for everything that does not sit on a sequential
// dependency (cuBLAS, fp32 or TF32) + three persistent cooperative kernels for the parts that do
// (encoder BiLSTM, attention chain, decoder-RNN chain; lstm_rec.cu / attn_chain.cu) + small fused
// element-wise kernels (model_kernels.cu).
#include <cublas_v2.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace msa {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define MSA_BLAS(expr)                                                                        \
    do {                                                                                      \
        cublasStatus_t _s = (expr);                                                           \
        if (_s != CUBLAS_STATUS_SUCCESS) {                                                    \
            msa::set_error("%s failed: cublas status %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); \
            return 1000 + (int)_s;                                                            \
        }                                                                                     \
    } while (0)

struct Dims {
    int B, T, L;
    int C, Kc, nEnc, Hh, Ds, Dsin, E, Pd, Ha, Hd, A, F, Kl, M, Cp, Kp, nPost, Cmax;
    int64_t BL, TB, BT, TBL;
    int conv_tc;      // convolutions as implicit GEMMs (gemm_tc.cu): no im2col buffers, tap-major weight copies instead
};

}  // namespace msa

using namespace msa;

struct msa_handle {
    msa_config cfg;
    int device = 0, sm_count = 0;
    size_t smem_limit = 0;
    cublasHandle_t blas = nullptr;
    std::vector<std::string> names;
    std::vector<int64_t> offs, numels;
    std::map<std::string, int64_t> off_by_name;
    int64_t total = 0;
    std::vector<int64_t> bn_offs;
    std::vector<int> bn_ch;
    int64_t bn_total = 0;
    // state of the last forward (for backward / get_buffer)
    bool fwd_valid = false;
    Dims d{};
    // optional per-kernel timing with CUDA events on the launch stream (bench.py's roofline leg)
    bool in_bwd = false;   // GEMM precision policy 1: fp32 GEMMs in the forward pass, TF32 in the backward pass
    bool prof = false;       // CUDA-event timing of the persistent kernels (bench.py)
    bool prof_inkernel = false;   // + per-phase cycle counters / per-warp traces inside them (profiles/ scripts only: slower kernels)
    int trace_t0 = 0;
    float* gemm_scratch = nullptr;   // K-split partial tiles of the tcgen05 GEMM (in the pass workspace)
    size_t gemm_scratch_floats = 0;
    cudaStream_t cur_stream = nullptr;
    // Launch-wide abort word of the persistent kernels (common.cuh, SpinGuard): device memory owned by the handle, zeroed at
    // msa_create and STICKY -- no pass clears it, so a polling time-out anywhere in a meta-step is still visible when the host (or
    // msa_abort_guard on the device) looks at it; msa_abort_clear resets it.
    unsigned int* abort_dev = nullptr;
    // Hand-written tcgen05 / TMA GEMM (gemm_tc.cu) for the x.W^T contractions.  tc_mode 2 (default): the fp32-accurate 3xTF32
    // products of at least tc_min MACs (3 x tc_min when K < 1024) -- at the default dimensions the LSTM input projections, the
    // k=5 convolutions of encoder and postnet, MW and the mel / gate projection, 1.4-4x faster there than cuBLAS's SIMT sgemm;
    // 1: every eligible contraction incl. the single-TF32 ones (cuBLAS's tensor-op kernels are a little faster: DESIGN.md 4.5);
    // 0: cuBLAS only.  Env MSA_GEMM_TC / MSA_GEMM_TC_MIN.
    int tc_mode = 2;
    long long tc_min = 100000000LL;
    long long tc_min_bwd = 30000000LL;     // mode 3: smallest backward contraction (MACs) routed to the tcgen05 kernel
    bool tc_enabled = true;
    // Encoder / postnet convolutions as implicit GEMMs on the tcgen05 kernel (activation operand = shifted 3-D TMA map, BatchNorm
    // statistics from the epilogue) under the tensor-core GEMM policies; the strict fp32 policy keeps im2col + cuBLAS.  Env MSA_CONV_TC=0: off.
    bool conv_tc = false;
    // K split of under-filled product grids (convolutions and plain GEMMs): set per pass -- on for a single pass (its GEMMs run one
    // after the other and a 32-tile grid leaves most SMs idle), off for the single-TF32 products while the stages of several tasks run
    // side by side on their own streams (the other tasks' kernels fill the SMs; the partial tiles and reduce launches only cost).
    // The fp32-accurate 3xTF32 products keep their split: the tensor core's fp32 accumulation truncates, an unsplit K = 2560 sum is
    // accurate to 1.3e-5 instead of 3.6e-6 (tests/test_gpu_conv_tc.py), and the TF32 backward amplifies that difference
    bool conv_split = true;
    int conv_split_mode = -1;    // env MSA_CONV_SPLIT: -1 per pass as above, 0 never, 1 always
    int rec_flags = -1;    // hand-off variant of the persistent kernels; -1 = per-kernel default (env MSA_REC_FLAGS overrides, development only)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    std::vector<int> prof_id;
    size_t prof_used = 0;
    // per-task inputs of the last forward (a grouped forward has G of them; a plain one G = 1)
    int G = 1;
    size_t ws_stride = 0;          // bytes between the workspaces of consecutive tasks of a group
    const int64_t *tokens[kGroupMax] = {}, *tok_len[kGroupMax] = {}, *mel_len[kGroupMax] = {}, *spk_ids[kGroupMax] = {};
    const float* spk_in[kGroupMax] = {};
    const uint8_t* masks[kGroupMax] = {};
    // Grouped chain kernels (chain_mma.cu): bf16x3 tensor-core gate products, G*B rows per hand-off.  mma_mode 0: never,
    // 1 (default): for grouped launches under the tensor-core GEMM policies, 2: also for single passes.  Env MSA_CHAIN_MMA.
    int mma_mode = 1;
    // side streams of a grouped pass: between two recurrences the tasks' GEMMs / element-wise kernels are independent (each task
    // has its own workspace slice, cuBLAS workspace and reduction scratch), so task g's stage runs on stream g and fills the SMs the
    // small per-task kernels of one task leave idle; fork / join with events around every stage.  Env MSA_GROUP_STREAMS=0: off.
    cudaStream_t aux[kGroupMax] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kGroupMax] = {};
    bool group_streams = true;
    // Per-task-weight groups (the test-split passes and later inner steps of a meta-batch: every task has its own adapted weights):
    // up to pt_group tasks per attention-chain launch, their recurrent weight slices streamed from L2 as ready-made fragments
    // (chain_mma.cu, PT variants).  pt_frag: the fragment buffer (pt_group x attn_chain_pt_frag_bytes, allocated on first use).
    // Env MSA_PT_GROUP (0 / 1: one launch per task).  Measured on B200 (default dims, B = 4, T = 200; ms per task): forward 2.13 single,
    // 1.49 / 1.43 in groups of 2 / 3 (four tasks' fragments, 78 MB, no longer stay in L2 next to MW); backward 2.42 single, 2.43 / 2.12
    // in groups of 2 / 3 (both operands of its recurrent tile on a cp.async ring) -- so both chains are grouped by three
    // (MSA_PT_BWD=0: backward chains as single-task launches).
    int pt_group = 3;
    bool pt_bwd = true;
    // one-shot event of the next backward pass (msa_backward_mark_event): recorded when every gradient outside the encoder is final
    cudaEvent_t bwd_ev = nullptr;
    void* pt_frag = nullptr;
    size_t pt_frag_task_bytes = 0;
    int64_t off(const std::string& n) const { return off_by_name.at(n); }
};

namespace msa {

static void add_param(msa_handle* h, const std::string& name, int64_t numel) {
    h->names.push_back(name);
    h->offs.push_back(h->total);
    h->numels.push_back(numel);
    h->off_by_name[name] = h->total;
    h->total += align_up(numel);
}

static void build_layout(msa_handle* h) {
    const msa_config& c = h->cfg;
    const int C = c.enc_dim, Hh = C / 2, E = C + c.spk_dim;
    add_param(h, "embedding.weight", (int64_t)c.n_symbols * C);
    for (int i = 0; i < c.enc_n_convs; ++i) {
        const std::string p = "encoder.convolutions." + std::to_string(i);
        add_param(h, p + ".0.conv.weight", (int64_t)C * C * c.enc_kernel);
        add_param(h, p + ".0.conv.bias", C);
        add_param(h, p + ".1.weight", C);
        add_param(h, p + ".1.bias", C);
    }
    for (const char* sfx : {"", "_reverse"}) {
        add_param(h, std::string("encoder.lstm.weight_ih_l0") + sfx, (int64_t)4 * Hh * C);
        add_param(h, std::string("encoder.lstm.weight_hh_l0") + sfx, (int64_t)4 * Hh * Hh);
        add_param(h, std::string("encoder.lstm.bias_ih_l0") + sfx, 4 * Hh);
        add_param(h, std::string("encoder.lstm.bias_hh_l0") + sfx, 4 * Hh);
    }
    if (c.spk_mode == 2) add_param(h, "speaker_embedder.weight", (int64_t)c.num_speakers * c.spk_in_dim);
    if (c.spk_mode == 1) {
        add_param(h, "speaker_lin.weight", (int64_t)c.spk_dim * c.spk_in_dim);
        add_param(h, "speaker_lin.bias", c.spk_dim);
    }
    add_param(h, "decoder.prenet.layers.0.linear_layer.weight", (int64_t)c.prenet_dim * c.n_mel);
    add_param(h, "decoder.prenet.layers.1.linear_layer.weight", (int64_t)c.prenet_dim * c.prenet_dim);
    add_param(h, "decoder.attention_rnn.weight_ih", (int64_t)4 * c.attn_rnn_dim * (c.prenet_dim + E));
    add_param(h, "decoder.attention_rnn.weight_hh", (int64_t)4 * c.attn_rnn_dim * c.attn_rnn_dim);
    add_param(h, "decoder.attention_rnn.bias_ih", 4 * c.attn_rnn_dim);
    add_param(h, "decoder.attention_rnn.bias_hh", 4 * c.attn_rnn_dim);
    const std::string a = "decoder.attention_layer.";
    add_param(h, a + "query_layer.linear_layer.weight", (int64_t)c.attn_dim * c.attn_rnn_dim);
    add_param(h, a + "inputs_layer.linear_layer.weight", (int64_t)c.attn_dim * E);
    add_param(h, a + "v.linear_layer.weight", c.attn_dim);
    add_param(h, a + "v.linear_layer.bias", 1);
    if (c.trans_agent) {
        add_param(h, a + "ta.weight", c.attn_rnn_dim + E);
        add_param(h, a + "ta.bias", 1);
    }
    add_param(h, a + "location_layer.location_conv1d.weight", (int64_t)c.loc_filters * 2 * c.loc_kernel);
    add_param(h, a + "location_layer.location_dense.linear_layer.weight", (int64_t)c.attn_dim * c.loc_filters);
    add_param(h, "decoder.decoder_rnn.weight_ih", (int64_t)4 * c.dec_rnn_dim * (c.attn_rnn_dim + E));
    add_param(h, "decoder.decoder_rnn.weight_hh", (int64_t)4 * c.dec_rnn_dim * c.dec_rnn_dim);
    add_param(h, "decoder.decoder_rnn.bias_ih", 4 * c.dec_rnn_dim);
    add_param(h, "decoder.decoder_rnn.bias_hh", 4 * c.dec_rnn_dim);
    add_param(h, "decoder.linear_projection.linear_layer.weight", (int64_t)c.n_mel * (c.dec_rnn_dim + E));
    add_param(h, "decoder.linear_projection.linear_layer.bias", c.n_mel);
    add_param(h, "decoder.gate_layer.linear_layer.weight", c.dec_rnn_dim + E);
    add_param(h, "decoder.gate_layer.linear_layer.bias", 1);
    for (int i = 0; i < c.post_n_convs; ++i) {
        const std::string p = "postnet.convolutions." + std::to_string(i);
        const int ci = i == 0 ? c.n_mel : c.post_dim, co = i == c.post_n_convs - 1 ? c.n_mel : c.post_dim;
        add_param(h, p + ".0.conv.weight", (int64_t)co * ci * c.post_kernel);
        add_param(h, p + ".0.conv.bias", co);
        add_param(h, p + ".1.weight", co);
        add_param(h, p + ".1.bias", co);
    }
    for (int i = 0; i < c.enc_n_convs + c.post_n_convs; ++i) {
        const int ch = i < c.enc_n_convs ? C : (i == c.enc_n_convs + c.post_n_convs - 1 ? c.n_mel : c.post_dim);
        h->bn_ch.push_back(ch);
        h->bn_offs.push_back(h->bn_total);
        h->bn_total += 2 * align_up(ch);
    }
}

static Dims make_dims(const msa_config& c, int B, int T, int L) {
    Dims d;
    d.B = B; d.T = T; d.L = L;
    d.C = c.enc_dim; d.Kc = c.enc_kernel; d.nEnc = c.enc_n_convs; d.Hh = d.C / 2;
    d.Ds = c.spk_dim; d.Dsin = c.spk_in_dim; d.E = d.C + d.Ds; d.Pd = c.prenet_dim;
    d.Ha = c.attn_rnn_dim; d.Hd = c.dec_rnn_dim; d.A = c.attn_dim; d.F = c.loc_filters; d.Kl = c.loc_kernel;
    d.M = c.n_mel; d.Cp = c.post_dim; d.Kp = c.post_kernel; d.nPost = c.post_n_convs;
    d.Cmax = std::max(d.M, d.Cp);
    d.BL = (int64_t)B * L; d.TB = (int64_t)T * B; d.BT = d.TB; d.TBL = d.TB * L;
    d.conv_tc = 0;
    return d;
}
// slab statistics of the largest convolution output: [slabs][3][channels]
static int64_t conv_stats_floats(const Dims& d) {
    const int64_t rows = std::max(d.BL, d.BT), tmax = std::max(d.T, d.L);
    const int64_t slabs = (rows + 31) / 32 + 4 * (int64_t)d.B * ((tmax + 127) / 128);
    return slabs * 3 * std::max(d.C, d.Cmax);
}

// ---- workspace -------------------------------------------------------------------------------------
#define WS_LIST(X)                                                                                     \
    X(spk_vec, d.B * d.Ds)                                                                             \
    X(enc_x, (d.nEnc + 1) * d.BL * d.C) X(enc_y, d.nEnc * d.BL * d.C) X(enc_bn, d.nEnc * 2 * d.C)      \
    X(enc_col, d.conv_tc ? 0 : d.nEnc * d.BL * d.Kc * d.C) X(x3_tm, d.BL * d.C)            \
    X(enc_wp, d.conv_tc ? (int64_t)d.nEnc * d.Kc * d.C * d.C : 0) X(post_wp, d.conv_tc ? (int64_t)d.nPost * d.Kp * d.Cmax * d.Cmax : 0) \
    X(conv_stats, d.conv_tc ? conv_stats_floats(d) : 0) \
    X(enc_zx, 2 * d.BL * 4 * d.Hh) X(enc_g, 2 * d.BL * 4 * d.Hh) X(enc_c, 2 * d.BL * d.Hh)             \
    X(enc_h, 2 * d.BL * d.Hh) X(memory, d.BL * d.E)                                                    \
    X(frames, (d.TB + d.B) * d.M) X(target, d.BT * d.M) X(p1, (d.TB + d.B) * d.Pd)                     \
    X(xpre, (d.TB + d.B) * d.Pd) X(xw, d.TB * 4 * d.Ha) X(pm, d.BL * d.A)                              \
    X(mw_rm, d.BL * 4 * d.Ha) X(mw_pm, d.BL * 4 * d.Ha)                                                \
    X(ha, d.TB * d.Ha) X(ca, d.TB * d.Ha) X(ga, d.TB * 4 * d.Ha) X(q, d.TB * d.A)                      \
    X(align_tm, d.TBL) X(cum, d.TBL) X(s, d.TBL * d.A) X(convf, d.TBL * d.F) X(znorm, d.TB)            \
    X(ebuf, d.TBL) X(ctx, d.TB * d.E)                                                                   \
    X(zd, d.TB * 4 * d.Hd) X(hd, d.TB * d.Hd) X(cd, d.TB * d.Hd) X(gd, d.TB * 4 * d.Hd)                \
    X(mel_tm, d.TB * d.M) X(gate_tm, d.TB)                                                             \
    X(post_x, (d.nPost + 1) * d.BT * d.Cmax) X(post_y, d.nPost * d.BT * d.Cmax)                        \
    X(post_bn, d.nPost * 2 * d.Cmax) X(post_col, d.conv_tc ? 0 : d.nPost * d.BT * d.Kp * d.Cmax)                       \
    X(post_bt, d.BT * d.M) X(gate_bt, d.BT) X(red_scr, kRedScrFloats) \
    X(loss_part, 1024) X(loss, 32) X(dpre, d.BT * d.M) X(dpost, d.BT * d.M) X(dgate, d.BT)             \
    X(bdx0, d.BT * d.Cmax) X(bdx1, d.BT * d.Cmax) X(bdy, d.BT * d.Cmax)                                \
    X(bdcol, d.conv_tc ? 0 : d.BT * d.Kp * d.Cmax)                            \
    X(bn_scr, 2 * std::max(d.Cmax, d.C)) X(dmel_bt, d.BT * d.M) X(dmel_tm, d.TB * d.M)                 \
    X(dgate_tm, d.TB) X(dhd, d.TB * d.Hd) X(dzd, d.TB * 4 * d.Hd) X(dha, d.TB * d.Ha)                  \
    X(dctx, d.TB * d.E) X(da_ext, d.TBL) X(dza, d.TB * 4 * d.Ha) X(dq, d.TB * d.A) X(de, d.TBL)        \
    X(ds, d.TBL * d.A) X(dconvf, d.TBL * d.F) X(dat, 2 * d.TBL) X(dpm, d.BL * d.A)                      \
    X(aplain, d.TBL) X(fsum, d.TB) X(ustash, d.TB) X(dzu, d.TB) X(mta, d.BL)                            \
    X(dmw, d.BL * 4 * d.Ha) X(dmem, d.BL * d.E) X(dxp, (d.TB + d.B) * d.Pd)                            \
    X(dp1, (d.TB + d.B) * d.Pd) X(denc_h, 2 * d.BL * d.Hh) X(dzx, 2 * d.BL * 4 * d.Hh)                 \
    X(dx3_tm, d.BL * d.C) X(edx0, d.BL * d.C) X(edx1, d.BL * d.C) X(edy, d.BL * d.C)                   \
    X(edcol, d.conv_tc ? 0 : d.BL * d.Kc * d.C) X(dspk, d.B * d.Ds)                             \
    X(wloc_part, (int64_t)wloc_grad_partials(d.T, d.B) * d.F * 2 * d.Kl) X(gemm_lo, tc_scratch_floats(d)) X(prof, kProfFloats) X(trace, kTraceFloats)

constexpr int64_t kProfFloats = 2 * 6 * 256 * 8;
constexpr int64_t kTraceWordsPerKernel = (int64_t)256 * 16 * 4 * 12 * 2;   // int64 words: [ctas][warps][steps][tags][2]
constexpr int64_t kTraceFloats = 2 * 6 * kTraceWordsPerKernel;
// scratch of the 3xTF32 operand split: (rows of A + rows of B) x padded K of the largest x.W^T contraction of the pass
static int64_t tc_scratch_floats(const Dims& d) {
    const int64_t rows_a = std::max<int64_t>(std::max<int64_t>(d.TB + d.B, d.BL), d.BT);
    const int64_t rows_b = std::max<int64_t>(std::max<int64_t>(4 * d.Ha, 4 * d.Hd), std::max<int64_t>(std::max<int64_t>(4 * d.Hh, d.C), d.Cmax));
    const int64_t k = std::max<int64_t>(std::max<int64_t>((int64_t)d.Kc * d.C, (int64_t)d.Kp * d.Cmax), std::max<int64_t>(d.Ha, d.Hd) + d.E + d.Pd);
    return (std::max(rows_a, rows_b) + rows_b) * ((k + 3) / 4 * 4) + 128;
}
struct Ws {
#define X(name, n) float* name; int64_t n_##name;
    WS_LIST(X)
#undef X
    unsigned int* abort_word;
    void* blas_ws;
    size_t blas_ws_bytes;
    size_t total_bytes;
};
constexpr size_t kBlasWs = 64u << 20;
constexpr int kProfCtas = 256, kProfSlots = 8;   // phase-cycle counters: [PROF_N kernels][kProfCtas][kProfSlots] int64

static Ws ws_layout(const Dims& d, void* base) {
    Ws w;
    size_t o = 0;
    char* b = static_cast<char*>(base);
#define X(name, n)                                        \
    w.n_##name = (int64_t)(n);                            \
    w.name = reinterpret_cast<float*>(b + o);             \
    o += (size_t)align_up((int64_t)(n), 64) * sizeof(float);
    WS_LIST(X)
#undef X
    w.abort_word = reinterpret_cast<unsigned int*>(b + o);
    o += 256;
    w.blas_ws = b + o;
    w.blas_ws_bytes = kBlasWs;
    o += kBlasWs;
    w.total_bytes = o;
    return w;
}

static Dims make_dims_h(const msa_handle* h, int B, int T, int L) {
    Dims d = make_dims(h->cfg, B, T, L);
    d.conv_tc = h->conv_tc ? 1 : 0;
    return d;
}

// ---- mask sections ---------------------------------------------------------------------------------
struct MaskSec {
    std::string name;
    int64_t off, numel;
    float p;
};
static std::vector<MaskSec> mask_sections(const msa_config& c, int B, int T, int L) {
    std::vector<MaskSec> v;
    int64_t o = 0;
    auto add = [&](const std::string& n, int64_t numel, float p) {
        v.push_back({n, o, numel, p});
        o += align_up(numel, 16);
    };
    for (int i = 0; i < c.enc_n_convs; ++i) add("enc" + std::to_string(i) + " [B][L][C]", (int64_t)B * L * c.enc_dim, 0.5f);
    add("prenet0 [T+1][B][P]", (int64_t)(T + 1) * B * c.prenet_dim, 0.5f);
    add("prenet1 [T+1][B][P]", (int64_t)(T + 1) * B * c.prenet_dim, 0.5f);
    add("attn_h [T][B][Ha]", (int64_t)T * B * c.attn_rnn_dim, c.p_attn_dropout);
    add("dec_h [T][B][Hd]", (int64_t)T * B * c.dec_rnn_dim, c.p_dec_dropout);
    for (int i = 0; i < c.post_n_convs; ++i)
        add("post" + std::to_string(i) + " [B][T][C]", (int64_t)B * T * (i == c.post_n_convs - 1 ? c.n_mel : c.post_dim), 0.5f);
    v.push_back({"", o, 0, 0.f});  // sentinel: total
    return v;
}

// ---- cuBLAS, row-major convention: C[MxN] = alpha * op(A) op(B) + beta * C -------------------------
// bias1 / bias2 (x . W^T only): C = op(A) op(B) + column bias.  On the tcgen05 route the bias is part of the epilogue (C is
// write-only); on the cuBLAS route C is filled with the bias rows first and accumulated onto (beta = 1).
static int gemm(msa_handle* h, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                const float* Bm, int64_t ldb, float beta, float* Cm, int64_t ldc, const float* bias1 = nullptr,
                const float* bias2 = nullptr) {
    if (M == 0 || N == 0) return 0;
    if (K == 0) {
        MSA_CHECK(beta == 1.f, MSA_E_ARG, "gemm: K == 0 with beta != 1");
        return 0;
    }
    const bool tf32 = h->cfg.gemm_tf32 >= 2 || (h->cfg.gemm_tf32 == 1 && h->in_bwd);
    // x . W^T contractions (both operands K-contiguous) go to the hand-written tcgen05 / TMA kernel (gemm_tc.cu) under the
    // tensor-core policies: 3xTF32 split where fp32 accuracy is required (forward of policy 1), single TF32 otherwise
    // (measured on B200, profiles/r01_gemm_tc_v14.txt: the kernel has ~10 us of fixed cost, so short-K / small products stay on cuBLAS)
    const long long macs = (long long)M * N * K;
    // fp32-accurate products (3xTF32 against cuBLAS's SIMT sgemm): measured per shape on B200 (profiles/r02_gemm_shapes.txt) the
    // kernel also wins or ties the small products (>= 1e7 MACs; one 256-deep prenet product is 2.7 us slower and rides along)
    const bool tc_small = !tf32 && macs >= 10000000LL;
    const bool tc_size = (macs >= h->tc_min && (macs >= 3 * h->tc_min || K >= 1024)) || tc_small;
    const bool nt = !ta && tb;
    const bool own_ok = h->tc_enabled && h->cfg.gemm_tf32 >= 1 && N >= 8 && M * N >= 4096 && gemm_tc_supported(M, N, K, A, lda, Bm, ldb, Cm, ldc) &&
                        (tf32 || (h->gemm_scratch && gemm_tc_scratch_floats(M, N, K) <= h->gemm_scratch_floats));
    // tc_mode 2: the fp32-accurate x . W^T products; 1: every eligible x . W^T product; 3: every eligible product incl. the backward
    // contractions with MN-major operands (dX = dY . W, dW = dY^T . X)
    const bool route = own_ok && ((h->tc_mode == 2 && !tf32 && nt && tc_size) || (h->tc_mode == 1 && nt && tc_size) ||
                                  (h->tc_mode == 3 && (nt ? (tc_size || tf32) : macs >= h->tc_min_bwd)));
    static const bool log_env = getenv("MSA_GEMM_LOG") != nullptr;      // development: one line per product (shape, pass half, route)
    if (log_env) fprintf(stderr, "gemm ta=%d tb=%d M=%lld N=%lld K=%lld bwd=%d tf32=%d own=%d\n", (int)ta, (int)tb, (long long)M, (long long)N,
                         (long long)K, (int)h->in_bwd, (int)tf32, (int)route);
    if (route)
        return gemm_tc(ta, tb, M, N, K, alpha, A, lda, Bm, ldb, beta, Cm, ldc, tf32 ? 2 : 0,
                       (h->gemm_scratch && gemm_tc_scratch_floats(M, N, K) <= h->gemm_scratch_floats && (h->conv_split || !tf32)) ? h->gemm_scratch : nullptr,
                       h->cur_stream, bias1, bias2);
    if (bias1 != nullptr) {
        MSA_CHECK(beta == 0.f && ldc == N, MSA_E_ARG, "gemm: a column bias needs beta == 0 and a dense C");
        MSA_TRY(k_fill_rows(Cm, bias1, bias2, M, (int)N, h->cur_stream));
        beta = 1.f;
    }
    const cublasComputeType_t ct = tf32 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F;
    MSA_BLAS(cublasGemmEx(h->blas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, (int)N, (int)M, (int)K, &alpha,
                          Bm, CUDA_R_32F, (int)ldb, A, CUDA_R_32F, (int)lda, &beta, Cm, CUDA_R_32F, (int)ldc, ct,
                          CUBLAS_GEMM_DEFAULT));
    return 0;
}
static int gemm_batched(msa_handle* h, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                        int64_t sa, const float* Bm, int64_t ldb, int64_t sb, float beta, float* Cm, int64_t ldc, int64_t sc,
                        int batch) {
    if (M == 0 || N == 0 || batch == 0) return 0;
    if (K == 0) return 0;
    const bool tf32 = h->cfg.gemm_tf32 >= 2 || (h->cfg.gemm_tf32 == 1 && h->in_bwd);
    const cublasComputeType_t ct = tf32 ? CUBLAS_COMPUTE_32F_FAST_TF32 : CUBLAS_COMPUTE_32F;
    MSA_BLAS(cublasGemmStridedBatchedEx(h->blas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, (int)N, (int)M,
                                        (int)K, &alpha, Bm, CUDA_R_32F, (int)ldb, sb, A, CUDA_R_32F, (int)lda, sa, &beta, Cm,
                                        CUDA_R_32F, (int)ldc, sc, batch, ct, CUBLAS_GEMM_DEFAULT));
    return 0;
}

// conv1d ("same") + BatchNorm(train) + activation + dropout, channels-last rows = B*Tn
// wp != nullptr: implicit GEMM on the tcgen05 kernel (wp = tap-major copy of the weight, stats = slab-statistics scratch): two
// launches per layer (three when the product is split over K) and no im2col matrix
static int conv_bn_fwd(msa_handle* h, cudaStream_t st, const float* params, const std::string& pfx, const float* x, float* y,
                       float* xout, float* col, const float* wp, float* stats, float* bn_mean, float* bn_invstd, float* running, int B, int Tn,
                       int Ci, int Co, int K, int act, const uint8_t* mask, float* red_scr) {
    // col (im2col of the layer input) is a per-layer buffer: the backward pass reuses it
    const int64_t rows = (int64_t)B * Tn;
    if (wp != nullptr) {
        const bool tf32 = h->cfg.gemm_tf32 >= 2 || (h->cfg.gemm_tf32 == 1 && h->in_bwd);
        float* scr = (h->gemm_scratch && conv_tc_scratch_floats(B, Tn, Ci, Co, K) <= h->gemm_scratch_floats) ? h->gemm_scratch : nullptr;
        if (!h->conv_split && tf32) scr = nullptr;      // the fp32-accurate product keeps its K split (accumulation accuracy, see msa_handle)
        MSA_TRY(conv_tc_fwd(x, B, Tn, Ci, wp, Co, K, params + h->off(pfx + ".0.conv.bias"), y, tf32 ? 2 : 0, scr, stats, st));
        MSA_TRY(k_bn_slab_act_drop_fwd(y, stats, conv_tc_stat_slabs(B, Tn, Ci, Co, K, scr != nullptr), rows, Co, bn_mean, bn_invstd, running,
                                       (int)align_up(Co), params + h->off(pfx + ".1.weight"), params + h->off(pfx + ".1.bias"), mask, 2.0f, act,
                                       xout, st));
        return 0;
    }
    // im2col columns are in (ci, k) order: the native weight [Co][Ci*K] is the GEMM operand
    MSA_TRY(k_im2col(x, col, B, Tn, Ci, K, st));
    MSA_TRY(gemm(h, false, true, rows, Co, (int64_t)K * Ci, 1.f, col, (int64_t)K * Ci, params + h->off(pfx + ".0.conv.weight"),
                 (int64_t)K * Ci, 0.f, y, Co, params + h->off(pfx + ".0.conv.bias")));
    MSA_TRY(k_bn_stats(y, rows, Co, bn_mean, bn_invstd, running, (int)align_up(Co), red_scr, st));
    MSA_TRY(k_bn_act_drop_fwd(y, bn_mean, bn_invstd, params + h->off(pfx + ".1.weight"), params + h->off(pfx + ".1.bias"), mask,
                              2.0f, act, xout, rows, Co, st));
    return 0;
}
// backward of the above: dout -> dx (through dropout, act, BN, conv); parameter grads into `grads`
static int conv_bn_bwd(msa_handle* h, cudaStream_t st, const float* params, float* grads, float gs, int acc,
                       const std::string& pfx, const float* x, const float* y, const float* dout, float* dx, float* dy, float* col,
                       const float* wp, float* dcol, float* scr, const float* bn_mean, const float* bn_invstd, int B, int Tn,
                       int Ci, int Co, int K, int act, const uint8_t* mask, bool need_dx, float* red_scr) {
    // im2col route: the im2col of x is still in `col` (written by conv_bn_fwd of the same pass); implicit route (wp): x itself
    const float* wt = params + h->off(pfx + ".0.conv.weight");
    const int64_t rows = (int64_t)B * Tn, KC = (int64_t)K * Ci;
    const bool tf32_tc = wp != nullptr && (h->cfg.gemm_tf32 >= 2 || (h->cfg.gemm_tf32 == 1 && h->in_bwd));
    MSA_TRY(k_bn_act_drop_bwd(dout, y, bn_mean, bn_invstd, params + h->off(pfx + ".1.weight"), params + h->off(pfx + ".1.bias"),
                              mask, 2.0f, act, dy, grads + h->off(pfx + ".1.weight"), grads + h->off(pfx + ".1.bias"), scr, rows,
                              Co, gs, acc, red_scr, st));
    // implicit route under the TF32 policy: the bias-gradient reduction (over the unrounded dy) also stores dy back rounded to the
    // nearest TF32 value -- its two consumers below are single-TF32 tensor-core products, which would otherwise TRUNCATE it
    MSA_TRY(k_colsum(dy, rows, Co, Co, grads + h->off(pfx + ".0.conv.bias"), gs, acc, nullptr, red_scr, st, tf32_tc ? dy : nullptr));
    if (wp != nullptr) {
        const bool tf32 = h->cfg.gemm_tf32 >= 2 || (h->cfg.gemm_tf32 == 1 && h->in_bwd);
        MSA_TRY(conv_tc_dw(dy, x, B, Tn, Co, Ci, K, gs, acc, grads + h->off(pfx + ".0.conv.weight"), tf32 ? 3 : 0, st));
        if (need_dx) {
            float* sc = (h->gemm_scratch && conv_tc_scratch_floats(B, Tn, Co, Ci, K) <= h->gemm_scratch_floats) ? h->gemm_scratch : nullptr;
            if (!h->conv_split) sc = nullptr;
            MSA_TRY(conv_tc_dx(dy, B, Tn, Co, wp, Ci, K, dx, tf32 ? 3 : 0, sc, st));
        }
        return 0;
    }
    // dW = dy^T . col lands in the parameter layout [Co][Ci][K] directly (scaled / accumulated by the GEMM itself)
    MSA_TRY(gemm(h, true, false, Co, KC, rows, gs, dy, Co, col, KC, acc ? 1.f : 0.f, grads + h->off(pfx + ".0.conv.weight"), KC));
    if (need_dx) {
        MSA_TRY(gemm(h, false, false, rows, KC, Co, 1.f, dy, Co, wt, KC, 0.f, dcol, KC));
        MSA_TRY(k_col2im(dcol, dx, B, Tn, Ci, K, st));
    }
    return 0;
}

// Wp[k][co][ci] copies of every encoder / postnet conv weight of one task, eight layers per launch
static int conv_repack_all(msa_handle* h, const Dims& d, const Ws& w, const float* params, cudaStream_t st) {
    std::vector<const float*> src;
    std::vector<float*> dst;
    std::vector<int> co, ci, k;
    for (int i = 0; i < d.nEnc; ++i) {
        src.push_back(params + h->off("encoder.convolutions." + std::to_string(i) + ".0.conv.weight"));
        dst.push_back(w.enc_wp + (int64_t)i * d.Kc * d.C * d.C);
        co.push_back(d.C); ci.push_back(d.C); k.push_back(d.Kc);
    }
    for (int i = 0; i < d.nPost; ++i) {
        src.push_back(params + h->off("postnet.convolutions." + std::to_string(i) + ".0.conv.weight"));
        dst.push_back(w.post_wp + (int64_t)i * d.Kp * d.Cmax * d.Cmax);
        co.push_back(i == d.nPost - 1 ? d.M : d.Cp); ci.push_back(i == 0 ? d.M : d.Cp); k.push_back(d.Kp);
    }
    for (size_t i0 = 0; i0 < src.size(); i0 += 8) {
        const int n = (int)std::min<size_t>(8, src.size() - i0);
        MSA_TRY(conv_tc_repack(n, src.data() + i0, dst.data() + i0, co.data() + i0, ci.data() + i0, k.data() + i0, st));
    }
    return 0;
}

// event-timing slots: the six single-task recurrences, then the same six as grouped launches (chain_mma.cu with G > 1); the
// in-kernel phase counters / traces exist once per recurrence (PROF_BASE slots)
// "_pt": attention-chain launches over several tasks with per-task weights (streamed weight fragments, chain_mma.cu)
enum ProfId { PROF_ENC_LSTM_FWD = 0, PROF_ATTN_FWD, PROF_DEC_LSTM_FWD, PROF_DEC_LSTM_BWD, PROF_ATTN_BWD, PROF_ENC_LSTM_BWD, PROF_BASE,
              PROF_ATTN_FWD_PT = 2 * PROF_BASE, PROF_ATTN_BWD_PT, PROF_N };
static const char* kProfNames[PROF_N] = {"enc_lstm_fwd", "attn_chain_fwd", "dec_lstm_fwd", "dec_lstm_bwd", "attn_chain_bwd", "enc_lstm_bwd",
                                         "enc_lstm_fwd_grp", "attn_chain_fwd_grp", "dec_lstm_fwd_grp", "dec_lstm_bwd_grp", "attn_chain_bwd_grp", "enc_lstm_bwd_grp",
                                         "attn_chain_fwd_pt", "attn_chain_bwd_pt"};
static int prof_base_id(int id) { return id == PROF_ATTN_FWD_PT ? PROF_ATTN_FWD : (id == PROF_ATTN_BWD_PT ? PROF_ATTN_BWD : id % PROF_BASE); }
struct ProfScope {
    msa_handle* h; cudaStream_t st; int slot = -1;
    ProfScope(msa_handle* h_, int id, cudaStream_t st_) : h(h_), st(st_) {
        if (!h->prof) return;
        if (h->prof_used == h->prof_ev.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            h->prof_ev.push_back({a, b});
            h->prof_id.push_back(id);
        }
        slot = (int)h->prof_used++;
        h->prof_id[slot] = id;
        cudaEventRecord(h->prof_ev[slot].first, st);
    }
    ~ProfScope() { if (slot >= 0) cudaEventRecord(h->prof_ev[slot].second, st); }
};

// per-phase cycle counters of the persistent kernels live in the workspace ("prof"); written only while profiling is on
static long long* prof_ptr(const msa_handle* h, const Ws& w, int id) {
    return h->prof_inkernel ? reinterpret_cast<long long*>(w.prof) + (size_t)id * kProfCtas * kProfSlots : nullptr;
}

static long long* trace_ptr(const msa_handle* h, const Ws& w, int id) {
    return h->prof_inkernel ? reinterpret_cast<long long*>(w.trace) + (size_t)id * kTraceWordsPerKernel : nullptr;
}

static int check_cfg(const msa_config& c) {
    MSA_CHECK(c.enc_dim > 0 && c.enc_dim % 8 == 0, MSA_E_UNSUPPORTED, "encoder_embedding_dim must be a positive multiple of 8");
    MSA_CHECK(c.enc_kernel % 2 == 1 && c.post_kernel % 2 == 1 && c.loc_kernel % 2 == 1, MSA_E_UNSUPPORTED, "kernel sizes must be odd");
    MSA_CHECK(c.attn_rnn_dim % 4 == 0 && c.dec_rnn_dim % 4 == 0, MSA_E_UNSUPPORTED, "rnn dims must be multiples of 4");
    MSA_CHECK(c.spk_mode >= 0 && c.spk_mode <= 2, MSA_E_ARG, "spk_mode");
    MSA_CHECK(c.loc_filters >= 1 && c.loc_filters <= 32, MSA_E_UNSUPPORTED, "attention_location_n_filters must be in [1,32]");
    MSA_CHECK(c.attn_norm == 0 || c.attn_norm == 1, MSA_E_ARG, "attn_norm");
    MSA_CHECK(c.enc_n_convs >= 1 && c.post_n_convs >= 2, MSA_E_UNSUPPORTED, "need >= 1 encoder conv and >= 2 postnet convs");
    return 0;
}

static int train_check(const msa_handle* h) {
    const msa_config& c = h->cfg;
    MSA_CHECK(!c.trans_agent || c.forward_attn, MSA_E_ARG, "trans_agent needs forward_attn (forward_attn.py:222)");
    return 0;
}

}  // namespace msa

// =====================================================================================================
extern "C" {

const char* msa_last_error_string(void) { return g_err; }
int msa_version(void) { return 1; }

int msa_create(const msa_config* cfg, int device, msa_handle** out) {
    MSA_CHECK(cfg && out, MSA_E_ARG, "msa_create: null argument");
    MSA_TRY(check_cfg(*cfg));
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    MSA_CHECK(e == cudaSuccess && ndev > 0, MSA_E_NODEVICE, "msa_create: no CUDA device (%s); there is no CPU fallback",
              cudaGetErrorString(e));
    MSA_CHECK(device >= 0 && device < ndev, MSA_E_ARG, "msa_create: device %d out of range", device);
    MSA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MSA_CUDA(cudaGetDeviceProperties(&prop, device));
    MSA_CHECK(prop.major >= 10, MSA_E_NODEVICE, "msa_create: device %s is sm_%d%d, this library is built for sm_100a only", prop.name,
              prop.major, prop.minor);
    int coop = 0;
    MSA_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    MSA_CHECK(coop, MSA_E_NODEVICE, "msa_create: device does not support cooperative launches");
    msa_handle* h = new msa_handle();
    h->cfg = *cfg;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_limit = prop.sharedMemPerBlockOptin;
    if (const char* e = getenv("MSA_REC_FLAGS")) h->rec_flags = atoi(e);
    if (const char* e = getenv("MSA_CHAIN_MMA")) h->mma_mode = atoi(e);
    if (const char* e = getenv("MSA_GROUP_STREAMS")) h->group_streams = atoi(e) != 0;
    if (const char* e = getenv("MSA_PT_GROUP")) h->pt_group = std::min(atoi(e), (int)kPtGroupMax);
    if (const char* e = getenv("MSA_PT_BWD")) h->pt_bwd = atoi(e) != 0;
    if (const char* e = getenv("MSA_GEMM_TC")) {
        h->tc_mode = atoi(e);
        h->tc_enabled = h->tc_mode != 0;
        if (h->tc_mode == 1) h->tc_min = 0;
    }
    {
        const msa_config& c = h->cfg;
        const bool env_on = !(getenv("MSA_CONV_TC") && atoi(getenv("MSA_CONV_TC")) == 0);
        h->conv_tc = env_on && h->tc_enabled && c.gemm_tf32 >= 1 && c.enc_dim % 4 == 0 && c.post_dim % 4 == 0 && c.n_mel % 4 == 0 &&
                     (c.enc_kernel & 1) && (c.post_kernel & 1) && c.enc_kernel <= 15 && c.post_kernel <= 15;
    }
    if (const char* e = getenv("MSA_CONV_SPLIT")) h->conv_split_mode = atoi(e);
    if (const char* e = getenv("MSA_GEMM_TC_MIN")) h->tc_min = atoll(e);
    if (const char* e = getenv("MSA_GEMM_TC_MIN_BWD")) h->tc_min_bwd = atoll(e);
    build_layout(h);
    cublasStatus_t s = cublasCreate(&h->blas);
    if (s != CUBLAS_STATUS_SUCCESS) {
        delete h;
        set_error("cublasCreate failed: %d", (int)s);
        return 1000 + (int)s;
    }
    cublasSetPointerMode(h->blas, CUBLAS_POINTER_MODE_HOST);
    if (cudaMalloc(&h->abort_dev, 256) != cudaSuccess || cudaMemset(h->abort_dev, 0, 256) != cudaSuccess) {
        cublasDestroy(h->blas);
        delete h;
        set_error("msa_create: cannot allocate the abort word");
        return MSA_E_NODEVICE;
    }
    *out = h;
    return 0;
}

int msa_destroy(msa_handle* h) {
    if (!h) return 0;
    if (h->blas) cublasDestroy(h->blas);
    if (h->abort_dev) cudaFree(h->abort_dev);
    if (h->pt_frag) cudaFree(h->pt_frag);
    for (int g = 0; g < kGroupMax; ++g) {
        if (h->aux[g]) cudaStreamDestroy(h->aux[g]);
        if (h->ev_join[g]) cudaEventDestroy(h->ev_join[g]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto& e : h->prof_ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete h;
    return 0;
}

int msa_sm_count(const msa_handle* h) { return h ? h->sm_count : 0; }
long long msa_launch_count(void) { return g_launches.load(); }
int msa_profile_enable(msa_handle* h, int enable) {
    MSA_CHECK(h, MSA_E_ARG, "msa_profile_enable: null handle");
    h->prof = enable != 0;
    h->prof_inkernel = enable >= 2;
    h->prof_used = 0;
    return 0;
}
int msa_profile_kernels(void) { return PROF_N; }
const char* msa_profile_name(int id) { return id >= 0 && id < PROF_N ? kProfNames[id] : ""; }
int msa_profile_read(msa_handle* h, double* total_ms, int64_t* counts) {
    MSA_CHECK(h && total_ms && counts, MSA_E_ARG, "msa_profile_read: null argument");
    for (int i = 0; i < PROF_N; ++i) { total_ms[i] = 0.0; counts[i] = 0; }
    for (size_t i = 0; i < h->prof_used; ++i) {
        MSA_CUDA(cudaEventSynchronize(h->prof_ev[i].second));
        float ms = 0.f;
        MSA_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[i].first, h->prof_ev[i].second));
        total_ms[h->prof_id[i]] += ms;
        counts[h->prof_id[i]] += 1;
    }
    h->prof_used = 0;
    return 0;
}
int msa_param_count(const msa_handle* h) { return h ? (int)h->names.size() : 0; }
int64_t msa_param_total(const msa_handle* h) { return h ? h->total : 0; }
int msa_param_info(const msa_handle* h, int i, const char** name, int64_t* offset, int64_t* numel) {
    MSA_CHECK(h && i >= 0 && i < (int)h->names.size(), MSA_E_ARG, "msa_param_info: index");
    if (name) *name = h->names[i].c_str();
    if (offset) *offset = h->offs[i];
    if (numel) *numel = h->numels[i];
    return 0;
}
int msa_bn_count(const msa_handle* h) { return h ? (int)h->bn_ch.size() : 0; }
int64_t msa_bn_total(const msa_handle* h) { return h ? h->bn_total : 0; }
int msa_bn_info(const msa_handle* h, int i, int64_t* offset, int32_t* channels) {
    MSA_CHECK(h && i >= 0 && i < (int)h->bn_ch.size(), MSA_E_ARG, "msa_bn_info: index");
    if (offset) *offset = h->bn_offs[i];
    if (channels) *channels = h->bn_ch[i];
    return 0;
}
int msa_mask_count(const msa_handle* h) { return h ? h->cfg.enc_n_convs + 4 + h->cfg.post_n_convs : 0; }
int64_t msa_mask_total(const msa_handle* h, int B, int T, int L) {
    if (!h) return 0;
    return mask_sections(h->cfg, B, T, L).back().off;
}
int msa_mask_info(const msa_handle* h, int i, int B, int T, int L, const char** name, int64_t* offset, int64_t* numel, float* p) {
    MSA_CHECK(h, MSA_E_ARG, "msa_mask_info: null handle");
    static thread_local std::vector<MaskSec> secs;
    secs = mask_sections(h->cfg, B, T, L);
    MSA_CHECK(i >= 0 && i + 1 < (int)secs.size(), MSA_E_ARG, "msa_mask_info: index");
    if (name) *name = secs[i].name.c_str();
    if (offset) *offset = secs[i].off;
    if (numel) *numel = secs[i].numel;
    if (p) *p = secs[i].p;
    return 0;
}
int msa_masks_generate(msa_handle* h, uint8_t* masks, int B, int T, int L, uint64_t seed, void* stream) {
    MSA_CHECK(h && masks, MSA_E_ARG, "msa_masks_generate: null argument");
    auto secs = mask_sections(h->cfg, B, T, L);
    std::vector<int64_t> offs, ns;
    std::vector<float> ps;
    for (size_t i = 0; i + 1 < secs.size(); ++i) { offs.push_back(secs[i].off); ns.push_back(secs[i].numel); ps.push_back(secs[i].p); }
    return k_masks_generate(masks, offs.data(), ns.data(), ps.data(), (int)offs.size(), seed, (cudaStream_t)stream);
}

size_t msa_workspace_bytes(const msa_handle* h, int B, int T, int L) {
    if (!h || B <= 0 || T <= 0 || L <= 0) return 0;
    return ws_layout(make_dims_h(h, B, T, L), nullptr).total_bytes;
}

}  // extern "C"

namespace msa {

// per-task inputs / outputs of a (grouped) forward
struct TaskIO {
    float* bn_stats;
    const int64_t *tokens, *tok_len, *mel_len, *spk_ids;
    const float *mels, *spk_vecs, *stop;
    const uint8_t* masks;
    float *mel_out, *mel_post_out, *gate_out, *align_out, *loss_out;
};

// Grouped chain launches (chain_mma.cu) are used when the configuration is inside what those kernels implement; otherwise the
// tasks of a group run the single-task persistent kernels one after the other (same results, no sharing of the hand-offs).
static bool use_mma_chains(const msa_handle* h, int G, int B, int T, int L, bool attn, bool bwd) {
    const msa_config& c = h->cfg;
    const bool have = attn ? (attn_chain_mma_supported(c, G, B, T, L, h->sm_count, h->smem_limit) &&
                              attn_chain_bwd_mma_supported(c, G, B, T, L, h->sm_count, h->smem_limit))
                           : chain_mma_supported(c, G, B, T, L, h->sm_count, h->smem_limit);
    if (!have) return false;
    // a batch the single-task kernels cannot hold (B > 8 at the default dimensions: their shared-memory staging of h) runs on the
    // grouped kernels under every policy -- their bf16x3 products are fp32-accurate to ~1e-5, well inside every stated tolerance
    const bool single_ok = attn ? attn_chain_single_ok(B, L, c.attn_rnn_dim, c.attn_dim, c.loc_filters, c.loc_kernel, h->sm_count,
                                                       h->smem_limit, c.forward_attn != 0)
                                : (lstm_rec_single_ok(B, c.dec_rnn_dim, 1, h->sm_count, h->smem_limit) &&
                                   lstm_rec_single_ok(B, c.enc_dim / 2, 2, h->sm_count, h->smem_limit));
    if (!single_ok) return true;
    if (h->mma_mode == 0 || (c.gemm_tf32 < 1 && h->mma_mode != 4)) return false;   // strict-fp32 policy: no tensor-core arithmetic
                                                                                    // (mode 4, tests: force the grouped kernels)
    // measured on B200 (profiles/r02_chain_mma_notes.txt): at 4 rows the tensor-core forward LSTM recurrences beat the fp32-FMA
    // ones (decoder RNN 0.70 vs 1.09 ms, encoder BiLSTM 0.17 vs 0.22 ms); the attention chain (2.52 vs 2.14 ms: its context term
    // streams MW from L2 where the single-task kernel keeps it in shared memory) and the backward kernels (a 16-step dz stream per
    // warp, latency-bound at few rows) do not: single passes take the forward LSTM kernels only
    if (h->mma_mode == 1 && G == 1 && (bwd || attn)) return false;
    if (h->mma_mode == 3 && attn) return false;                                     // development: LSTM recurrences only
    if (h->mma_mode == 5 && G == 1) return false;                                   // development: grouped launches only
    return true;
}

// Per-task-weight groups: the G tasks are cut into chunks of at most pt_group tasks, as even as possible (8 -> 4 + 4, 5 -> 3 + 2); a
// chunk of two or more tasks is one launch of the per-task-weight attention chain, a chunk of one task a single-task launch.
static std::vector<std::pair<int, int>> pt_chunks(const msa_handle* h, int G, int B, int T, int L, bool bwd) {
    std::vector<std::pair<int, int>> out;      // (first task, tasks)
    const bool want = G >= 2 && h->pt_group >= 2 && (!bwd || h->pt_bwd) && use_mma_chains(h, 2, B, T, L, true, bwd);
    // the largest chunk size <= pt_group whose launches fit (shared memory: the weight / dz rings grow with the tasks per launch)
    for (int cap = want ? h->pt_group : 1; cap >= 2; --cap) {
        const int n = (G + cap - 1) / cap;
        bool ok = true;
        int g0 = 0;
        out.clear();
        for (int i = 0; i < n; ++i) {
            const int sz = G / n + (i < G % n ? 1 : 0);
            out.push_back({g0, sz});
            g0 += sz;
            if (sz >= 2 && !attn_chain_pt_supported(h->cfg, sz, B, T, L, h->sm_count, h->smem_limit)) ok = false;
        }
        if (ok) return out;
    }
    out.clear();
    for (int g = 0; g < G; ++g) out.push_back({g, 1});
    return out;
}
static int pt_frag_ensure(msa_handle* h) {
    if (h->pt_frag) return 0;
    h->pt_frag_task_bytes = attn_chain_pt_frag_bytes(h->cfg, h->sm_count);
    MSA_CUDA(cudaMalloc(&h->pt_frag, h->pt_frag_task_bytes * (size_t)kPtGroupMax));
    return 0;
}

// Fork / join of the per-task stages of a grouped pass over the handle's side streams (created on first use).
struct StageFork {
    msa_handle* h; cudaStream_t main; int G; bool on;
    StageFork(msa_handle* h_, cudaStream_t main_, int G_) : h(h_), main(main_), G(G_), on(false) {
        if (G <= 1 || !h->group_streams) return;
        for (int g = 1; g < G; ++g)
            if (!h->aux[g] && cudaStreamCreateWithFlags(&h->aux[g], cudaStreamNonBlocking) != cudaSuccess) return;
        if (!h->ev_fork && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) return;
        for (int g = 1; g < G; ++g)
            if (!h->ev_join[g] && cudaEventCreateWithFlags(&h->ev_join[g], cudaEventDisableTiming) != cudaSuccess) return;
        on = true;
    }
    // stream of task g's stage (task 0 stays on the caller's stream); begin() before the first task, end() after the last one
    int begin() {
        if (!on) return 0;
        MSA_CUDA(cudaEventRecord(h->ev_fork, main));
        for (int g = 1; g < G; ++g) MSA_CUDA(cudaStreamWaitEvent(h->aux[g], h->ev_fork, 0));
        return 0;
    }
    cudaStream_t stream(int g) const { return on && g > 0 ? h->aux[g] : main; }
    int enter(int g) {
        cudaStream_t s = stream(g);
        h->cur_stream = s;
        MSA_BLAS(cublasSetStream(h->blas, s));
        return 0;
    }
    int end(const std::vector<Ws>& W) {
        if (on) {
            for (int g = 1; g < G; ++g) {
                MSA_CUDA(cudaEventRecord(h->ev_join[g], h->aux[g]));
                MSA_CUDA(cudaStreamWaitEvent(main, h->ev_join[g], 0));
            }
        }
        h->cur_stream = main;
        MSA_BLAS(cublasSetStream(h->blas, main));
        MSA_BLAS(cublasSetWorkspace(h->blas, W[0].blas_ws, W[0].blas_ws_bytes));
        return 0;
    }
};

// One teacher-forced forward pass for G tasks that share `params` (the theta_0 train-split passes of a meta-batch, maml.py:38-54):
// everything that is not a recurrence runs task by task in that task's own workspace slice, the three recurrences run ONCE for
// all G*B rows when the grouped kernels apply.
static int train_forward_impl(msa_handle* h, int G, char* wsp, size_t ws_stride, const float* const* params_g, const TaskIO* ios, int B, int T,
                              int L, cudaStream_t st) {
    MSA_TRY(train_check(h));
    const Dims d = make_dims_h(h, B, T, L);
    std::vector<Ws> W;
    for (int g = 0; g < G; ++g) W.push_back(ws_layout(d, wsp + (size_t)g * ws_stride));
    MSA_CUDA(cudaSetDevice(h->device));
    MSA_BLAS(cublasSetStream(h->blas, st));
    MSA_BLAS(cublasSetWorkspace(h->blas, W[0].blas_ws, W[0].blas_ws_bytes));
    h->fwd_valid = false;
    h->in_bwd = false;
    h->cur_stream = st;
    const msa_config& c = h->cfg;
    const auto secs = mask_sections(c, B, T, L);
    const int iPre = d.nEnc, iAttn = d.nEnc + 2, iDec = d.nEnc + 3, iPost = d.nEnc + 4;
    auto PG = [&](int g, const std::string& n) { return params_g[g] + h->off(n); };
    bool shared = true;      // one weight buffer for the whole group (the theta_0 train passes) or per-task weights (the test passes)
    for (int g = 1; g < G; ++g) shared = shared && params_g[g] == params_g[0];
    const int Gk = shared ? G : 1;      // tasks one recurrence launch can carry
    const bool mma = use_mma_chains(h, Gk, B, T, L, false, false), mma_attn = use_mma_chains(h, Gk, B, T, L, true, false);
    const int64_t tstride = (int64_t)(ws_stride / sizeof(float));
    StageFork fk(h, st, G);
    h->conv_split = h->conv_split_mode < 0 ? !fk.on : h->conv_split_mode != 0;
    const int H4e = 4 * d.Hh;
    const std::string at = "decoder.attention_layer.";
    const int H4a = 4 * d.Ha, H4d = 4 * d.Hd, ldA = d.Pd + d.E, ldD = d.Ha + d.E, ldP = d.Hd + d.E;
    const bool fa = c.forward_attn != 0, ta = fa && c.trans_agent != 0;

    // ---- stage 1: speaker vector, encoder convolutions, BiLSTM input projection ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < G; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        const TaskIO& io = ios[g];
        auto mk = [&](int i) { return io.masks + secs[i].off; };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        MSA_CUDA(cudaMemsetAsync(w.red_scr, 0, sizeof(unsigned int) * kRedTickets, st));      // tickets of the chunked column reductions
        // speaker vector (tacotron2nv.py:104-109)
        if (c.spk_mode == 0) {
            MSA_TRY(k_scale_copy(io.spk_vecs, w.spk_vec, (int64_t)B * d.Ds, 1.f, 0, st));
        } else if (c.spk_mode == 1) {
            MSA_TRY(k_fill_rows(w.spk_vec, P("speaker_lin.bias"), nullptr, B, d.Ds, st));
            MSA_TRY(gemm(h, false, true, B, d.Ds, d.Dsin, 1.f, io.spk_vecs, d.Dsin, P("speaker_lin.weight"), d.Dsin, 1.f, w.spk_vec, d.Ds));
        } else {
            MSA_TRY(k_embedding_fwd(P("speaker_embedder.weight"), io.spk_ids, w.spk_vec, B, d.Ds, c.num_speakers, st));
        }
        if (d.conv_tc) MSA_TRY(conv_repack_all(h, d, w, params, st));      // tap-major copies of this task's conv weights (forward + backward)
        // encoder (tacotron2nv.py:88, encoder.py:35-52)
        MSA_TRY(k_embedding_fwd(P("embedding.weight"), io.tokens, w.enc_x, (int)d.BL, d.C, c.n_symbols, st));
        const int64_t ex = d.BL * d.C;
        for (int i = 0; i < d.nEnc; ++i) {
            float* run = io.bn_stats ? io.bn_stats + h->bn_offs[i] : nullptr;
            MSA_TRY(conv_bn_fwd(h, st, params, "encoder.convolutions." + std::to_string(i), w.enc_x + i * ex, w.enc_y + i * ex,
                                w.enc_x + (i + 1) * ex, d.conv_tc ? nullptr : w.enc_col + i * d.BL * d.Kc * d.C,
                                d.conv_tc ? w.enc_wp + (int64_t)i * d.Kc * d.C * d.C : nullptr, w.conv_stats,
                                w.enc_bn + i * 2 * d.C, w.enc_bn + i * 2 * d.C + d.C, run, B, L, d.C, d.C, d.Kc, 1, mk(i), w.red_scr));
        }
        MSA_TRY(k_transpose01(w.enc_x + d.nEnc * ex, w.x3_tm, B, L, d.C, st));
        for (int dir = 0; dir < 2; ++dir) {
            const std::string sfx = dir ? "_reverse" : "";
            float* zx = w.enc_zx + (int64_t)dir * d.BL * H4e;
            MSA_TRY(k_fill_rows(zx, P("encoder.lstm.bias_ih_l0" + sfx), P("encoder.lstm.bias_hh_l0" + sfx), d.BL, H4e, st));
            MSA_TRY(gemm(h, false, true, d.BL, H4e, d.C, 1.f, w.x3_tm, d.C, P("encoder.lstm.weight_ih_l0" + sfx), d.C, 1.f, zx, H4e));
        }
    }
    MSA_TRY(fk.end(W));
    // ---- encoder BiLSTM (persistent) ----
    {
        auto make = [&](int g) {
            const Ws& w = W[g];
            LstmRecParams lp{};
            lp.T = L; lp.B = B; lp.H = d.Hh; lp.ndir = 2;
            lp.zin = w.enc_zx; lp.whh = PG(g, "encoder.lstm.weight_hh_l0");
            lp.whh_dir_stride = h->off("encoder.lstm.weight_hh_l0_reverse") - h->off("encoder.lstm.weight_hh_l0");
            lp.hout = w.enc_h; lp.cout = w.enc_c; lp.gates = w.enc_g; lp.mask = nullptr; lp.drop_scale = 1.f;
            lp.lengths = ios[g].tok_len; lp.abort_word = h->abort_dev; lp.prof = prof_ptr(h, w, PROF_ENC_LSTM_FWD);
            lp.trace = trace_ptr(h, w, PROF_ENC_LSTM_FWD); lp.trace_t0 = h->trace_t0; lp.flags = h->rec_flags >= 0 ? h->rec_flags : 1;
            return lp;
        };
        if (mma && shared) {
            ProfScope ps(h, PROF_ENC_LSTM_FWD + (G > 1 ? PROF_BASE : 0), st);
            LstmRecParams lp = make(0);
            lp.G = G; lp.tstride = tstride;
            for (int g = 0; g < G; ++g) lp.lengths_g[g] = ios[g].tok_len;
            MSA_TRY(launch_lstm_rec_fwd_mma(lp, h->sm_count, h->smem_limit, st));
        } else if (mma) {       // per-task weights: one launch per task (G = 1, that task's pointers)
            for (int g = 0; g < G; ++g) {
                ProfScope ps(h, PROF_ENC_LSTM_FWD, st);
                LstmRecParams lp = make(g);
                lp.G = 1; lp.tstride = 0;
                MSA_TRY(launch_lstm_rec_fwd_mma(lp, h->sm_count, h->smem_limit, st));
            }
        } else {
            for (int g = 0; g < G; ++g) {
                ProfScope ps(h, PROF_ENC_LSTM_FWD, st);
                MSA_TRY(launch_lstm_rec_fwd(make(g), h->sm_count, h->smem_limit, st));
            }
        }
    }
    // ---- stage 2: memory, decoder set-up (decoder.py:290-302, forward_attn.py:103-116) ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < G; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        const TaskIO& io = ios[g];
        auto mk = [&](int i) { return io.masks + secs[i].off; };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        MSA_TRY(k_build_memory(w.enc_h, w.spk_vec, w.memory, B, L, d.Hh, d.Ds, st));
        if (c.residual_encoder)      // tacotron2nv.py:94-96: + the embedded characters
            MSA_TRY(k_embedding_add(P("embedding.weight"), io.tokens, w.memory, (int)d.BL, d.C, d.E, c.n_symbols, st));
        MSA_TRY(gemm(h, false, true, d.BL, d.A, d.E, 1.f, w.memory, d.E, P(at + "inputs_layer.linear_layer.weight"), d.E, 0.f, w.pm, d.A));
        const float* Wia = P("decoder.attention_rnn.weight_ih");
        MSA_TRY(gemm(h, false, true, H4a, d.BL, d.E, 1.f, Wia + d.Pd, ldA, w.memory, d.E, 0.f, w.mw_rm, d.BL));
        MSA_TRY(k_prep_mels(io.mels, w.frames, w.target, B, d.M, T, st));
        const int64_t nfr = d.TB + B;
        MSA_TRY(gemm(h, false, true, nfr, d.Pd, d.M, 1.f, w.frames, d.M, P("decoder.prenet.layers.0.linear_layer.weight"), d.M, 0.f, w.p1, d.Pd));
        MSA_TRY(k_relu_drop_fwd(w.p1, mk(iPre), 2.f, nfr * d.Pd, st));
        MSA_TRY(gemm(h, false, true, nfr, d.Pd, d.Pd, 1.f, w.p1, d.Pd, P("decoder.prenet.layers.1.linear_layer.weight"), d.Pd, 0.f, w.xpre, d.Pd));
        MSA_TRY(k_relu_drop_fwd(w.xpre, mk(iPre + 1), 2.f, nfr * d.Pd, st));
        MSA_TRY(gemm(h, false, true, d.TB, H4a, d.Pd, 1.f, w.xpre, d.Pd, Wia, ldA, 0.f, w.xw, H4a, P("decoder.attention_rnn.bias_ih"),
                     P("decoder.attention_rnn.bias_hh")));
        if (ta)     // context half of the transition agent: ctx(t) = alpha(t) . memory  =>  W_ta[:E] . ctx(t) = alpha(t) . (memory . W_ta[:E])
            MSA_TRY(gemm(h, false, true, d.BL, 1, d.E, 1.f, w.memory, d.E, P(at + "ta.weight"), d.E + d.Ha, 0.f, w.mta, 1));
    }
    MSA_TRY(fk.end(W));
    // ---- attention chain (persistent) ----
    {
        auto make = [&](int g) {
            const Ws& w = W[g];
            AttnChainParams ap{};
            ap.T = T; ap.B = B; ap.L = L; ap.Ha = d.Ha; ap.A = d.A; ap.F = d.F; ap.Kl = d.Kl; ap.norm = c.attn_norm;
            ap.xw = w.xw; ap.whh = PG(g, "decoder.attention_rnn.weight_hh"); ap.mw_rm = w.mw_rm;
            ap.wq = PG(g, at + "query_layer.linear_layer.weight"); ap.pm = w.pm;
            ap.wloc = PG(g, at + "location_layer.location_conv1d.weight");
            ap.wld = PG(g, at + "location_layer.location_dense.linear_layer.weight");
            ap.v = PG(g, at + "v.linear_layer.weight"); ap.bv = PG(g, at + "v.linear_layer.bias");
            ap.mask = c.p_attn_dropout > 0.f ? ios[g].masks + secs[iAttn].off : nullptr;
            ap.drop_scale = 1.f / (1.f - c.p_attn_dropout);
            ap.ha = w.ha; ap.ca = w.ca; ap.ga = w.ga; ap.q = w.q; ap.align = w.align_tm; ap.cum = w.cum; ap.s = w.s;
            ap.fa = fa; ap.ta = ta; ap.aplain = w.aplain; ap.fsum = w.fsum; ap.ustash = w.ustash;
            if (ta) { ap.mta = w.mta; ap.wta_h = PG(g, at + "ta.weight") + d.E; ap.bta = PG(g, at + "ta.bias"); }
            ap.convf = w.convf; ap.znorm = w.znorm; ap.e = w.ebuf; ap.abort_word = h->abort_dev; ap.prof = prof_ptr(h, w, PROF_ATTN_FWD);
            ap.trace = trace_ptr(h, w, PROF_ATTN_FWD); ap.trace_t0 = h->trace_t0; ap.flags = h->rec_flags >= 0 ? h->rec_flags : 0;
            return ap;
        };
        if (mma_attn && shared) {
            ProfScope ps(h, PROF_ATTN_FWD + (G > 1 ? PROF_BASE : 0), st);
            AttnChainParams ap = make(0);
            ap.G = G; ap.tstride = tstride;
            for (int g = 0; g < G; ++g) ap.mask_g[g] = c.p_attn_dropout > 0.f ? ios[g].masks + secs[iAttn].off : nullptr;
            MSA_TRY(launch_attn_chain_fwd_mma(ap, h->sm_count, h->smem_limit, st));
        } else {       // per-task weights: chunks of tasks on the streamed-weight kernel, single tasks on the single-task kernels
            for (const auto& ch : pt_chunks(h, G, B, T, L, false)) {
                const int g0 = ch.first, Gc = ch.second;
                if (Gc >= 2) {
                    ProfScope ps(h, PROF_ATTN_FWD_PT, st);
                    MSA_TRY(pt_frag_ensure(h));
                    AttnChainParams ap = make(g0);
                    ap.G = Gc; ap.tstride = tstride; ap.pt = 1;
                    ap.wfrag = static_cast<const uint4*>(h->pt_frag); ap.wfrag_stride = (int64_t)(h->pt_frag_task_bytes / sizeof(uint4));
                    for (int i = 0; i < Gc; ++i) {
                        const int g = g0 + i;
                        ap.mask_g[i] = c.p_attn_dropout > 0.f ? ios[g].masks + secs[iAttn].off : nullptr;
                        ap.whh_g[i] = PG(g, "decoder.attention_rnn.weight_hh");
                        ap.wq_g[i] = PG(g, at + "query_layer.linear_layer.weight");
                        ap.wloc_g[i] = PG(g, at + "location_layer.location_conv1d.weight");
                        ap.wld_g[i] = PG(g, at + "location_layer.location_dense.linear_layer.weight");
                        ap.v_g[i] = PG(g, at + "v.linear_layer.weight");
                        ap.bv_g[i] = PG(g, at + "v.linear_layer.bias");
                    }
                    MSA_TRY(launch_attn_chain_fwd_mma(ap, h->sm_count, h->smem_limit, st));
                } else if (mma_attn) {
                    ProfScope ps(h, PROF_ATTN_FWD, st);
                    AttnChainParams ap = make(g0);
                    ap.G = 1; ap.tstride = 0;
                    MSA_TRY(launch_attn_chain_fwd_mma(ap, h->sm_count, h->smem_limit, st));
                } else {
                    ProfScope ps(h, PROF_ATTN_FWD, st);
                    MSA_TRY(launch_attn_chain_fwd(make(g0), h->sm_count, h->smem_limit, st));
                }
            }
        }
    }
    // ---- stage 3: context vectors, decoder-RNN input projection ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < G; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        // ctx[t][b] = a[t][b] . memory[b]  (forward_attn.py:217), batched over b
        MSA_TRY(gemm_batched(h, false, false, T, d.E, L, 1.f, w.align_tm, d.BL, L, w.memory, d.E, (int64_t)L * d.E, 0.f, w.ctx,
                             (int64_t)B * d.E, d.E, B));
        const float* Wid = P("decoder.decoder_rnn.weight_ih");
        MSA_TRY(gemm(h, false, true, d.TB, H4d, d.Ha, 1.f, w.ha, d.Ha, Wid, ldD, 0.f, w.zd, H4d, P("decoder.decoder_rnn.bias_ih"),
                     P("decoder.decoder_rnn.bias_hh")));
        MSA_TRY(gemm(h, false, true, d.TB, H4d, d.E, 1.f, w.ctx, d.E, Wid + d.Ha, ldD, 1.f, w.zd, H4d));
    }
    MSA_TRY(fk.end(W));
    // ---- decoder RNN chain (decoder.py:260-265, persistent) ----
    {
        auto make = [&](int g) {
            const Ws& w = W[g];
            LstmRecParams lp{};
            lp.T = T; lp.B = B; lp.H = d.Hd; lp.ndir = 1;
            lp.zin = w.zd; lp.whh = PG(g, "decoder.decoder_rnn.weight_hh"); lp.whh_dir_stride = 0;
            lp.hout = w.hd; lp.cout = w.cd; lp.gates = w.gd;
            lp.mask = c.p_dec_dropout > 0.f ? ios[g].masks + secs[iDec].off : nullptr;
            lp.drop_scale = 1.f / (1.f - c.p_dec_dropout);
            lp.lengths = nullptr; lp.abort_word = h->abort_dev; lp.prof = prof_ptr(h, w, PROF_DEC_LSTM_FWD);
            lp.trace = trace_ptr(h, w, PROF_DEC_LSTM_FWD); lp.trace_t0 = h->trace_t0; lp.flags = h->rec_flags >= 0 ? h->rec_flags : 1;
            return lp;
        };
        if (mma && shared) {
            ProfScope ps(h, PROF_DEC_LSTM_FWD + (G > 1 ? PROF_BASE : 0), st);
            LstmRecParams lp = make(0);
            lp.G = G; lp.tstride = tstride;
            for (int g = 0; g < G; ++g) lp.mask_g[g] = c.p_dec_dropout > 0.f ? ios[g].masks + secs[iDec].off : nullptr;
            MSA_TRY(launch_lstm_rec_fwd_mma(lp, h->sm_count, h->smem_limit, st));
        } else if (mma) {       // per-task weights: one launch per task (G = 1, that task's pointers)
            for (int g = 0; g < G; ++g) {
                ProfScope ps(h, PROF_DEC_LSTM_FWD, st);
                LstmRecParams lp = make(g);
                lp.G = 1; lp.tstride = 0;
                MSA_TRY(launch_lstm_rec_fwd_mma(lp, h->sm_count, h->smem_limit, st));
            }
        } else {
            for (int g = 0; g < G; ++g) {
                ProfScope ps(h, PROF_DEC_LSTM_FWD, st);
                MSA_TRY(launch_lstm_rec_fwd(make(g), h->sm_count, h->smem_limit, st));
            }
        }
    }
    // ---- stage 4: projections, postnet, outputs, loss ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < G; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        const TaskIO& io = ios[g];
        auto mk = [&](int i) { return io.masks + secs[i].off; };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        // mel / gate projections (decoder.py:267-270)
        const float* Wp = P("decoder.linear_projection.linear_layer.weight");
        const float* Wg = P("decoder.gate_layer.linear_layer.weight");
        MSA_TRY(gemm(h, false, true, d.TB, d.M, d.Hd, 1.f, w.hd, d.Hd, Wp, ldP, 0.f, w.mel_tm, d.M,
                     P("decoder.linear_projection.linear_layer.bias")));
        MSA_TRY(gemm(h, false, true, d.TB, d.M, d.E, 1.f, w.ctx, d.E, Wp + d.Hd, ldP, 1.f, w.mel_tm, d.M));
        MSA_TRY(k_fill_rows(w.gate_tm, P("decoder.gate_layer.linear_layer.bias"), nullptr, d.TB, 1, st));
        MSA_TRY(gemm(h, false, true, d.TB, 1, d.Hd, 1.f, w.hd, d.Hd, Wg, ldP, 1.f, w.gate_tm, 1));
        MSA_TRY(gemm(h, false, true, d.TB, 1, d.E, 1.f, w.ctx, d.E, Wg + d.Hd, ldP, 1.f, w.gate_tm, 1));
        // postnet (decoder.py:63-72, tacotron2nv.py:123-124)
        const int64_t px = d.BT * d.Cmax;
        MSA_TRY(k_transpose01(w.mel_tm, w.post_x, T, B, d.M, st));       // [T][B][M] -> [B][T][M]
        MSA_TRY(k_transpose01(w.gate_tm, w.gate_bt, T, B, 1, st));
        for (int i = 0; i < d.nPost; ++i) {
            const int ci = i == 0 ? d.M : d.Cp, co = i == d.nPost - 1 ? d.M : d.Cp;
            float* run = io.bn_stats ? io.bn_stats + h->bn_offs[d.nEnc + i] : nullptr;
            MSA_TRY(conv_bn_fwd(h, st, params, "postnet.convolutions." + std::to_string(i), w.post_x + i * px, w.post_y + i * px,
                                w.post_x + (i + 1) * px, d.conv_tc ? nullptr : w.post_col + i * d.BT * d.Kp * d.Cmax,
                                d.conv_tc ? w.post_wp + (int64_t)i * d.Kp * d.Cmax * d.Cmax : nullptr, w.conv_stats,
                                w.post_bn + i * 2 * d.Cmax, w.post_bn + i * 2 * d.Cmax + d.Cmax, run, B, T, ci, co, d.Kp,
                                i < d.nPost - 1 ? 2 : 0, mk(iPost + i), w.red_scr));
        }
        MSA_TRY(k_add(w.post_x, w.post_x + d.nPost * px, w.post_bt, d.BT * d.M, st));
        // outputs in the reference layouts
        if (io.mel_out) MSA_TRY(k_bt_to_ref(w.post_x, io.mel_out, B, T, d.M, st));
        if (io.mel_post_out) MSA_TRY(k_bt_to_ref(w.post_bt, io.mel_post_out, B, T, d.M, st));
        if (io.gate_out) MSA_TRY(k_scale_copy(w.gate_bt, io.gate_out, d.BT, 1.f, 0, st));
        if (io.align_out) MSA_TRY(k_transpose01(w.align_tm, io.align_out, T, B, L, st));
        // loss + d(loss)/d(outputs)
        if (io.stop) {
            MSA_TRY(k_loss(w.post_x, w.post_bt, w.gate_bt, w.target, io.stop, io.mel_len, B, T, d.M, c.loss_reduction,
                           c.loss_pos_weight, w.loss_part, w.loss, w.dpre, w.dpost, w.dgate, st));
            if (io.loss_out) MSA_TRY(k_scale_copy(w.loss, io.loss_out, 1, 1.f, 0, st));
        }
    }
    MSA_TRY(fk.end(W));
    h->d = d;
    h->G = G;
    h->ws_stride = ws_stride;
    for (int g = 0; g < G; ++g) {
        h->tokens[g] = ios[g].tokens; h->tok_len[g] = ios[g].tok_len; h->mel_len[g] = ios[g].mel_len; h->spk_ids[g] = ios[g].spk_ids;
        h->spk_in[g] = ios[g].spk_vecs; h->masks[g] = ios[g].masks;
    }
    h->fwd_valid = true;
    return 0;
}

}  // namespace msa

extern "C" {

int msa_train_forward(msa_handle* h, void* wsp, size_t ws_bytes, const float* params, float* bn_stats, const int64_t* tokens,
                      const int64_t* token_lengths, const float* mels, const int64_t* mel_lengths, const float* speaker_vecs,
                      const int64_t* speaker_ids, const float* stop_targets, const uint8_t* masks, int B, int T, int L,
                      float* mel_out, float* mel_post_out, float* gate_out, float* align_out, float* loss_out, void* stream) {
    MSA_CHECK(h && wsp && params && tokens && token_lengths && mels && mel_lengths && masks, MSA_E_ARG, "msa_train_forward: null argument");
    MSA_CHECK(B >= 1 && T >= 1 && L >= 1, MSA_E_ARG, "msa_train_forward: bad dims B=%d T=%d L=%d", B, T, L);
    MSA_CHECK(h->cfg.spk_mode == 2 ? speaker_ids != nullptr : speaker_vecs != nullptr, MSA_E_ARG, "msa_train_forward: speaker input missing");
    const size_t need = ws_layout(make_dims_h(h, B, T, L), nullptr).total_bytes;
    MSA_CHECK(ws_bytes >= need, MSA_E_WORKSPACE, "msa_train_forward: workspace %zu < %zu bytes", ws_bytes, need);
    MSA_CHECK(((uintptr_t)wsp & 255) == 0, MSA_E_ARG, "msa_train_forward: workspace must be 256-byte aligned");
    TaskIO io{bn_stats, tokens, token_lengths, mel_lengths, speaker_ids, mels, speaker_vecs, stop_targets, masks,
              mel_out, mel_post_out, gate_out, align_out, loss_out};
    return train_forward_impl(h, 1, static_cast<char*>(wsp), 0, &params, &io, B, T, L, (cudaStream_t)stream);
}

size_t msa_group_workspace_bytes(const msa_handle* h, int G, int B, int T, int L) {
    if (!h || G < 1 || G > kGroupMax || B <= 0 || T <= 0 || L <= 0) return 0;
    return (size_t)G * ((ws_layout(make_dims_h(h, B, T, L), nullptr).total_bytes + 255) / 256 * 256);
}

int msa_train_forward_group(msa_handle* h, int G, void* wsp, size_t ws_bytes, const float* const* params, float* const* bn_stats,
                            const int64_t* const* tokens, const int64_t* const* token_lengths, const float* const* mels,
                            const int64_t* const* mel_lengths, const float* const* speaker_vecs, const int64_t* const* speaker_ids,
                            const float* const* stop_targets, const uint8_t* const* masks, int B, int T, int L, float* loss_out,
                            void* stream) {
    MSA_CHECK(h && wsp && params && tokens && token_lengths && mels && mel_lengths && masks && stop_targets, MSA_E_ARG,
              "msa_train_forward_group: null argument");
    MSA_CHECK(G >= 1 && G <= kGroupMax, MSA_E_ARG, "msa_train_forward_group: G=%d outside [1,%d]", G, kGroupMax);
    MSA_CHECK(B >= 1 && T >= 1 && L >= 1, MSA_E_ARG, "msa_train_forward_group: bad dims B=%d T=%d L=%d", B, T, L);
    const size_t stride = msa_group_workspace_bytes(h, 1, B, T, L);
    MSA_CHECK(ws_bytes >= (size_t)G * stride, MSA_E_WORKSPACE, "msa_train_forward_group: workspace %zu < %zu bytes", ws_bytes, (size_t)G * stride);
    MSA_CHECK(((uintptr_t)wsp & 255) == 0, MSA_E_ARG, "msa_train_forward_group: workspace must be 256-byte aligned");
    TaskIO ios[kGroupMax];
    for (int g = 0; g < G; ++g) {
        MSA_CHECK(tokens[g] && token_lengths[g] && mels[g] && mel_lengths[g] && masks[g] && stop_targets[g], MSA_E_ARG,
                  "msa_train_forward_group: null input of task %d", g);
        const float* sv = speaker_vecs ? speaker_vecs[g] : nullptr;
        const int64_t* si = speaker_ids ? speaker_ids[g] : nullptr;
        MSA_CHECK(h->cfg.spk_mode == 2 ? si != nullptr : sv != nullptr, MSA_E_ARG, "msa_train_forward_group: speaker input of task %d missing", g);
        ios[g] = TaskIO{bn_stats ? bn_stats[g] : nullptr, tokens[g], token_lengths[g], mel_lengths[g], si, mels[g], sv, stop_targets[g],
                        masks[g], nullptr, nullptr, nullptr, nullptr, loss_out ? loss_out + g : nullptr};
    }
    for (int g = 0; g < G; ++g) MSA_CHECK(params[g] != nullptr, MSA_E_ARG, "msa_train_forward_group: null weights of task %d", g);
    return train_forward_impl(h, G, static_cast<char*>(wsp), stride, params, ios, B, T, L, (cudaStream_t)stream);
}

int msa_train_loss(msa_handle* h, void* wsp, const float* stop_targets, const int64_t* mel_lengths, int reduction, float pos_weight,
                   float* loss_out, void* stream) {
    MSA_CHECK(h && wsp && h->fwd_valid, MSA_E_STATE, "msa_train_loss: no forward pass in this workspace");
    MSA_CHECK(stop_targets && mel_lengths && loss_out, MSA_E_ARG, "msa_train_loss: null argument");
    MSA_CHECK(reduction == 0 || reduction == 1, MSA_E_ARG, "msa_train_loss: reduction must be 0 (none) or 1 (mean)");
    const Dims& d = h->d;
    const Ws w = ws_layout(d, wsp);
    cudaStream_t st = (cudaStream_t)stream;
    MSA_TRY(k_loss(w.post_x, w.post_bt, w.gate_bt, w.target, stop_targets, mel_lengths, d.B, d.T, d.M, reduction, pos_weight,
                   w.loss_part, w.loss, w.dpre, w.dpost, w.dgate, st));
    MSA_TRY(k_scale_copy(w.loss, loss_out, 1, 1.f, 0, st));
    return 0;
}

int msa_train_mcd(msa_handle* h, void* wsp, const int64_t* mel_lengths, int which, float* mcd_out, void* stream) {
    MSA_CHECK(h && wsp && h->fwd_valid, MSA_E_STATE, "msa_train_mcd: no forward pass in this workspace");
    MSA_CHECK(mel_lengths && mcd_out && (which == 0 || which == 1), MSA_E_ARG, "msa_train_mcd: bad argument");
    const Dims& d = h->d;
    const Ws w = ws_layout(d, wsp);
    cudaStream_t st = (cudaStream_t)stream;
    MSA_TRY(k_mcd(which == 0 ? w.post_x : w.post_bt, w.target, mel_lengths, d.B, d.T, d.M, w.loss_part, w.loss + 8, st));
    MSA_TRY(k_scale_copy(w.loss + 8, mcd_out, 1, 1.f, 0, st));
    return 0;
}

size_t msa_loss_scratch_floats(int B, int T, int n_mel) {
    return (size_t)7 * align_up((int64_t)B * T * n_mel, 64) + (size_t)2 * align_up((int64_t)B * T, 64) + 2048;
}
int msa_tacotron2_loss(const float* mel, const float* mel_post, const float* gate, const float* mel_target, const float* stop_targets,
                       const int64_t* mel_lengths, int B, int T, int n_mel, int reduction, float pos_weight, float* scratch,
                       float* loss_out, float* d_mel, float* d_mel_post, float* d_gate, void* stream) {
    MSA_CHECK(mel && mel_post && gate && mel_target && stop_targets && mel_lengths && scratch && loss_out, MSA_E_ARG,
              "msa_tacotron2_loss: null argument");
    MSA_CHECK(B >= 1 && T >= 1 && n_mel >= 1 && (reduction == 0 || reduction == 1), MSA_E_ARG, "msa_tacotron2_loss: bad dims / reduction");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = align_up((int64_t)B * T * n_mel, 64), ng = align_up((int64_t)B * T, 64);
    float *pre_bt = scratch, *post_bt = pre_bt + n, *tgt_bt = post_bt + n, *dpre = tgt_bt + n, *dpost = dpre + n, *dgate = dpost + n;
    float* part = dgate + ng;      // 1024 partials + the scalar
    float* lossb = part + 1024;
    MSA_TRY(k_ref_to_bt(mel, pre_bt, B, T, n_mel, st));
    MSA_TRY(k_ref_to_bt(mel_post, post_bt, B, T, n_mel, st));
    MSA_TRY(k_ref_to_bt(mel_target, tgt_bt, B, T, n_mel, st));
    MSA_TRY(k_loss(pre_bt, post_bt, gate, tgt_bt, stop_targets, mel_lengths, B, T, n_mel, reduction, pos_weight, part, lossb, dpre, dpost,
                   dgate, st));
    MSA_TRY(k_scale_copy(lossb, loss_out, 1, 1.f, 0, st));
    if (d_mel) MSA_TRY(k_bt_to_ref(dpre, d_mel, B, T, n_mel, st));
    if (d_mel_post) MSA_TRY(k_bt_to_ref(dpost, d_mel_post, B, T, n_mel, st));
    if (d_gate) MSA_TRY(k_scale_copy(dgate, d_gate, (int64_t)B * T, 1.f, 0, st));
    return 0;
}

int msa_loss_grads(msa_handle* h, void* wsp, float* d_mel, float* d_mel_post, float* d_gate, void* stream) {
    MSA_CHECK(h && wsp && h->fwd_valid, MSA_E_STATE, "msa_loss_grads: no forward pass in this workspace");
    const Dims& d = h->d;
    const Ws w = ws_layout(d, wsp);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_mel) MSA_TRY(k_bt_to_ref(w.dpre, d_mel, d.B, d.T, d.M, st));
    if (d_mel_post) MSA_TRY(k_bt_to_ref(w.dpost, d_mel_post, d.B, d.T, d.M, st));
    if (d_gate) MSA_TRY(k_scale_copy(w.dgate, d_gate, d.BT, 1.f, 0, st));
    return 0;
}

}  // extern "C"

namespace msa {

// grads of every parameter whose name starts with one of `prefixes` (or with none of them: invert) <- 0 unless accumulating:
// a detached sub-graph leaves p.grad = None in the reference (tacotron2nv.py:90-121), which mix_grad / the optimizers read as zero
static int zero_grads(const msa_handle* h, float* grads, const std::vector<std::string>& prefixes, bool invert, int acc, cudaStream_t st) {
    if (acc) return 0;
    for (size_t i = 0; i < h->names.size(); ++i) {
        bool hit = false;
        for (const auto& p : prefixes) hit = hit || h->names[i].compare(0, p.size(), p) == 0;
        if (hit != invert) MSA_CUDA(cudaMemsetAsync(grads + h->offs[i], 0, sizeof(float) * (size_t)h->numels[i], st));
    }
    return 0;
}

// Backward of train_forward_impl: the gradient of task g's loss w.r.t. the shared parameters goes to grads_g[g]
// (= autograd.grad(loss_g, fast_weights), maml.py:54 / 71-74); the three reverse-time recurrences run once for the whole group.
static int train_backward_impl(msa_handle* h, char* wsp, const float* const* params_g, const float* const* d_mel, const float* const* d_mel_post,
                               const float* const* d_gate, float* const* grads_g, int acc, float gs, cudaStream_t st) {
    const Dims d = h->d;
    const int NG = h->G;
    const size_t ws_stride = h->ws_stride;
    std::vector<Ws> W;
    for (int g = 0; g < NG; ++g) W.push_back(ws_layout(d, wsp + (size_t)g * ws_stride));
    const bool ext = d_mel != nullptr;
    MSA_CUDA(cudaSetDevice(h->device));
    MSA_BLAS(cublasSetStream(h->blas, st));
    MSA_BLAS(cublasSetWorkspace(h->blas, W[0].blas_ws, W[0].blas_ws_bytes));
    const msa_config& c = h->cfg;
    h->in_bwd = true;
    h->cur_stream = st;
    const int B = d.B, T = d.T, L = d.L;
    const auto secs = mask_sections(c, B, T, L);
    const int iPre = d.nEnc, iAttn = d.nEnc + 2, iDec = d.nEnc + 3, iPost = d.nEnc + 4;
    auto PG = [&](int g, const std::string& n) { return params_g[g] + h->off(n); };
    bool shared = true;      // one weight buffer for the whole group (the theta_0 train passes) or per-task weights (the test passes)
    for (int g = 1; g < NG; ++g) shared = shared && params_g[g] == params_g[0];
    const int Gk = shared ? NG : 1;      // tasks one recurrence launch can carry
    const float beta = acc ? 1.f : 0.f;
    const int H4a = 4 * d.Ha, H4d = 4 * d.Hd, H4e = 4 * d.Hh, ldA = d.Pd + d.E, ldD = d.Ha + d.E, ldP = d.Hd + d.E;
    const std::string at = "decoder.attention_layer.";
    const bool fa = c.forward_attn != 0, ta = fa && c.trans_agent != 0;
    const bool mma = use_mma_chains(h, Gk, B, T, L, false, true), mma_attn = use_mma_chains(h, Gk, B, T, L, true, true);
    const int64_t tstride = (int64_t)(ws_stride / sizeof(float));
    StageFork fk(h, st, NG);
    h->conv_split = h->conv_split_mode < 0 ? !fk.on : h->conv_split_mode != 0;
    // msa_backward_mark_event: recorded on the pass stream once the gradients of everything but the encoder are final (one-shot;
    // also on every early return, a waiter must never see an unrecorded event)
    auto mark_decoder_done = [&]() {
        if (h->bwd_ev) {
            cudaEventRecord(h->bwd_ev, st);
            h->bwd_ev = nullptr;
        }
    };

    // ---- stage 1: postnet and projections ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < NG; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        const float* Wia = P("decoder.attention_rnn.weight_ih");
        (void)Wia;
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        float* grads = grads_g[g];
        auto mk = [&](int i) { return h->masks[g] + secs[i].off; };
        auto G = [&](const std::string& n) { return grads + h->off(n); };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        (void)mk; (void)G;
        if (ext) {
            MSA_TRY(k_ref_to_bt(d_mel[g], w.dpre, B, T, d.M, st));
            MSA_TRY(k_ref_to_bt(d_mel_post[g], w.dpost, B, T, d.M, st));
            MSA_TRY(k_scale_copy(d_gate[g], w.dgate, d.BT, 1.f, 0, st));
        }
        // ---- postnet backward ----
        const int64_t px = d.BT * d.Cmax;
        const float* dcur = w.dpost;
        float* pingpong[2] = {w.bdx0, w.bdx1};
        for (int i = d.nPost - 1; i >= 0; --i) {
            const int ci = i == 0 ? d.M : d.Cp, co = i == d.nPost - 1 ? d.M : d.Cp;
            float* dx = pingpong[i & 1];
            MSA_TRY(conv_bn_bwd(h, st, params, grads, gs, acc, "postnet.convolutions." + std::to_string(i), w.post_x + i * px,
                                w.post_y + i * px, dcur, dx, w.bdy, d.conv_tc ? nullptr : w.post_col + i * d.BT * d.Kp * d.Cmax,
                                d.conv_tc ? w.post_wp + (int64_t)i * d.Kp * d.Cmax * d.Cmax : nullptr, w.bdcol, w.bn_scr,
                                w.post_bn + i * 2 * d.Cmax, w.post_bn + i * 2 * d.Cmax + d.Cmax, B, T, ci, co, d.Kp,
                                i < d.nPost - 1 ? 2 : 0, mk(iPost + i), true, w.red_scr));
            dcur = dx;
        }
        if (c.freeze_decoder) {      // tacotron2nv.py:118-121: mel / gate / alignments detached, nothing flows further back
            MSA_TRY(zero_grads(h, grads, {"postnet."}, true, acc, st));
            continue;
        }
        // d(pre-postnet mel) = loss term + residual + postnet input (tacotron2nv.py:123-124)
        MSA_TRY(k_add3(w.dpre, w.dpost, dcur, w.dmel_bt, d.BT * d.M, st));
        MSA_TRY(k_transpose01(w.dmel_bt, w.dmel_tm, B, T, d.M, st));
        MSA_TRY(k_transpose01(w.dgate, w.dgate_tm, B, T, 1, st));
        // ---- projections backward (decoder.py:267-270) ----
        const float* Wp = P("decoder.linear_projection.linear_layer.weight");
        const float* Wg = P("decoder.gate_layer.linear_layer.weight");
        float* gWp = G("decoder.linear_projection.linear_layer.weight");
        float* gWg = G("decoder.gate_layer.linear_layer.weight");
        MSA_TRY(gemm(h, false, false, d.TB, d.Hd, d.M, 1.f, w.dmel_tm, d.M, Wp, ldP, 0.f, w.dhd, d.Hd));
        MSA_TRY(gemm(h, false, false, d.TB, d.Hd, 1, 1.f, w.dgate_tm, 1, Wg, ldP, 1.f, w.dhd, d.Hd));
        MSA_TRY(gemm(h, false, false, d.TB, d.E, d.M, 1.f, w.dmel_tm, d.M, Wp + d.Hd, ldP, 0.f, w.dctx, d.E));
        MSA_TRY(gemm(h, false, false, d.TB, d.E, 1, 1.f, w.dgate_tm, 1, Wg + d.Hd, ldP, 1.f, w.dctx, d.E));
        MSA_TRY(gemm(h, true, false, d.M, d.Hd, d.TB, gs, w.dmel_tm, d.M, w.hd, d.Hd, beta, gWp, ldP));
        MSA_TRY(gemm(h, true, false, d.M, d.E, d.TB, gs, w.dmel_tm, d.M, w.ctx, d.E, beta, gWp + d.Hd, ldP));
        MSA_TRY(gemm(h, true, false, 1, d.Hd, d.TB, gs, w.dgate_tm, 1, w.hd, d.Hd, beta, gWg, ldP));
        MSA_TRY(gemm(h, true, false, 1, d.E, d.TB, gs, w.dgate_tm, 1, w.ctx, d.E, beta, gWg + d.Hd, ldP));
        MSA_TRY(k_colsum(w.dmel_tm, d.TB, d.M, d.M, G("decoder.linear_projection.linear_layer.bias"), gs, acc, nullptr, w.red_scr, st));
        MSA_TRY(k_colsum(w.dgate_tm, d.TB, 1, 1, G("decoder.gate_layer.linear_layer.bias"), gs, acc, nullptr, w.red_scr, st));
    }
    MSA_TRY(fk.end(W));
    if (c.freeze_decoder) { mark_decoder_done(); return 0; }
    // ---- decoder RNN chain backward (persistent) ----
    {
        auto make = [&](int g) {
            const Ws& w = W[g];
            LstmRecBwdParams bp{};
            bp.T = T; bp.B = B; bp.H = d.Hd; bp.ndir = 1;
            bp.whh = PG(g, "decoder.decoder_rnn.weight_hh"); bp.whh_dir_stride = 0;
            bp.gates = w.gd; bp.cout = w.cd; bp.dh_ext = w.dhd; bp.dz = w.dzd;
            bp.mask = c.p_dec_dropout > 0.f ? h->masks[g] + secs[iDec].off : nullptr;
            bp.drop_scale = 1.f / (1.f - c.p_dec_dropout);
            bp.lengths = nullptr; bp.abort_word = h->abort_dev; bp.prof = prof_ptr(h, w, PROF_DEC_LSTM_BWD);
            bp.trace = trace_ptr(h, w, PROF_DEC_LSTM_BWD); bp.trace_t0 = h->trace_t0; bp.flags = h->rec_flags >= 0 ? h->rec_flags : 1;
            return bp;
        };
        if (mma && shared) {
            ProfScope ps(h, PROF_DEC_LSTM_BWD + (NG > 1 ? PROF_BASE : 0), st);
            LstmRecBwdParams bp = make(0);
            bp.G = NG; bp.tstride = tstride;
            for (int g = 0; g < NG; ++g) bp.mask_g[g] = c.p_dec_dropout > 0.f ? h->masks[g] + secs[iDec].off : nullptr;
            MSA_TRY(launch_lstm_rec_bwd_mma(bp, h->sm_count, h->smem_limit, st));
        } else if (mma) {       // per-task weights: one launch per task (G = 1, that task's pointers)
            for (int g = 0; g < NG; ++g) {
                ProfScope ps(h, PROF_DEC_LSTM_BWD, st);
                LstmRecBwdParams bp = make(g);
                bp.G = 1; bp.tstride = 0;
                MSA_TRY(launch_lstm_rec_bwd_mma(bp, h->sm_count, h->smem_limit, st));
            }
        } else {
            for (int g = 0; g < NG; ++g) {
                ProfScope ps(h, PROF_DEC_LSTM_BWD, st);
                MSA_TRY(launch_lstm_rec_bwd(make(g), h->sm_count, h->smem_limit, st));
            }
        }
    }
    // ---- stage 2: decoder-RNN input gradients, context backward ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < NG; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        const float* Wia = P("decoder.attention_rnn.weight_ih");
        (void)Wia;
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        float* grads = grads_g[g];
        auto mk = [&](int i) { return h->masks[g] + secs[i].off; };
        auto G = [&](const std::string& n) { return grads + h->off(n); };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        (void)mk; (void)G;
        const float* Wid = P("decoder.decoder_rnn.weight_ih");
        float* gWid = G("decoder.decoder_rnn.weight_ih");
        MSA_TRY(gemm(h, false, false, d.TB, d.Ha, H4d, 1.f, w.dzd, H4d, Wid, ldD, 0.f, w.dha, d.Ha));
        MSA_TRY(gemm(h, false, false, d.TB, d.E, H4d, 1.f, w.dzd, H4d, Wid + d.Ha, ldD, 1.f, w.dctx, d.E));
        MSA_TRY(gemm(h, true, false, H4d, d.Ha, d.TB, gs, w.dzd, H4d, w.ha, d.Ha, beta, gWid, ldD));
        MSA_TRY(gemm(h, true, false, H4d, d.E, d.TB, gs, w.dzd, H4d, w.ctx, d.E, beta, gWid + d.Ha, ldD));
        float* gWhd = G("decoder.decoder_rnn.weight_hh");
        if (T > 1) {
            MSA_TRY(gemm(h, true, false, H4d, d.Hd, d.TB - B, gs, w.dzd + (int64_t)B * H4d, H4d, w.hd, d.Hd, beta, gWhd, d.Hd));
        } else if (!acc) {
            MSA_CUDA(cudaMemsetAsync(gWhd, 0, sizeof(float) * (size_t)H4d * d.Hd, st));
        }
        MSA_TRY(k_colsum(w.dzd, d.TB, H4d, H4d, G("decoder.decoder_rnn.bias_ih"), gs, acc, G("decoder.decoder_rnn.bias_hh"), w.red_scr, st));
        // ---- context backward: ctx = align . memory ----
        MSA_TRY(gemm_batched(h, false, true, T, L, d.E, 1.f, w.dctx, (int64_t)B * d.E, d.E, w.memory, d.E, (int64_t)L * d.E, 0.f,
                             w.da_ext, d.BL, L, B));
        MSA_TRY(gemm_batched(h, true, false, L, d.E, T, 1.f, w.align_tm, d.BL, L, w.dctx, (int64_t)B * d.E, d.E, 0.f, w.dmem, d.E,
                             (int64_t)L * d.E, B));
        // (W_ih[:, prenet:] . memory^T)^T for the attention chain's context-path backward
        MSA_TRY(gemm(h, false, true, d.BL, H4a, d.E, 1.f, w.memory, d.E, Wia + d.Pd, ldA, 0.f, w.mw_pm, H4a));
    }
    MSA_TRY(fk.end(W));
    // ---- attention chain backward (persistent) ----
    {
        auto make = [&](int g) {
            const Ws& w = W[g];
            AttnChainBwdParams bp{};
            bp.T = T; bp.B = B; bp.L = L; bp.Ha = d.Ha; bp.A = d.A; bp.F = d.F; bp.Kl = d.Kl; bp.norm = c.attn_norm;
            bp.whh = PG(g, "decoder.attention_rnn.weight_hh"); bp.mw_pm = w.mw_pm;
            bp.wq = PG(g, at + "query_layer.linear_layer.weight");
            bp.wloc = PG(g, at + "location_layer.location_conv1d.weight");
            bp.wld = PG(g, at + "location_layer.location_dense.linear_layer.weight");
            bp.v = PG(g, at + "v.linear_layer.weight");
            bp.mask = c.p_attn_dropout > 0.f ? h->masks[g] + secs[iAttn].off : nullptr;
            bp.drop_scale = 1.f / (1.f - c.p_attn_dropout);
            bp.ga = w.ga; bp.ca = w.ca; bp.align = w.align_tm; bp.s = w.s; bp.znorm = w.znorm;
            bp.dha_ext = w.dha; bp.da_ext = w.da_ext;
            bp.dza = w.dza; bp.dq = w.dq; bp.de = w.de; bp.ds = w.ds; bp.dconvf = w.dconvf; bp.dat = w.dat;
            bp.fa = fa; bp.ta = ta; bp.aplain = w.aplain; bp.fsum = w.fsum; bp.ustash = w.ustash; bp.dzu = w.dzu;
            if (ta) { bp.mta = w.mta; bp.wta_h = PG(g, at + "ta.weight") + d.E; }
            bp.abort_word = h->abort_dev; bp.prof = prof_ptr(h, w, PROF_ATTN_BWD); bp.trace = trace_ptr(h, w, PROF_ATTN_BWD);
            bp.trace_t0 = h->trace_t0; bp.flags = h->rec_flags >= 0 ? h->rec_flags : 0;
            return bp;
        };
        if (mma_attn && shared) {
            ProfScope ps(h, PROF_ATTN_BWD + (NG > 1 ? PROF_BASE : 0), st);
            AttnChainBwdParams bp = make(0);
            bp.G = NG; bp.tstride = tstride;
            for (int g = 0; g < NG; ++g) bp.mask_g[g] = c.p_attn_dropout > 0.f ? h->masks[g] + secs[iAttn].off : nullptr;
            MSA_TRY(launch_attn_chain_bwd_mma(bp, h->sm_count, h->smem_limit, st));
        } else {       // per-task weights: chunks of tasks on the streamed-weight kernel, single tasks on the single-task kernels
            for (const auto& ch : pt_chunks(h, NG, B, T, L, true)) {
                const int g0 = ch.first, Gc = ch.second;
                if (Gc >= 2) {
                    ProfScope ps(h, PROF_ATTN_BWD_PT, st);
                    MSA_TRY(pt_frag_ensure(h));
                    AttnChainBwdParams bp = make(g0);
                    bp.G = Gc; bp.tstride = tstride; bp.pt = 1;
                    bp.wfrag = static_cast<const uint4*>(h->pt_frag); bp.wfrag_stride = (int64_t)(h->pt_frag_task_bytes / sizeof(uint4));
                    for (int i = 0; i < Gc; ++i) {
                        const int g = g0 + i;
                        bp.mask_g[i] = c.p_attn_dropout > 0.f ? h->masks[g] + secs[iAttn].off : nullptr;
                        bp.whh_g[i] = PG(g, "decoder.attention_rnn.weight_hh");
                        bp.wq_g[i] = PG(g, at + "query_layer.linear_layer.weight");
                        bp.wloc_g[i] = PG(g, at + "location_layer.location_conv1d.weight");
                        bp.wld_g[i] = PG(g, at + "location_layer.location_dense.linear_layer.weight");
                        bp.v_g[i] = PG(g, at + "v.linear_layer.weight");
                    }
                    MSA_TRY(launch_attn_chain_bwd_mma(bp, h->sm_count, h->smem_limit, st));
                } else if (mma_attn) {
                    ProfScope ps(h, PROF_ATTN_BWD, st);
                    AttnChainBwdParams bp = make(g0);
                    bp.G = 1; bp.tstride = 0;
                    MSA_TRY(launch_attn_chain_bwd_mma(bp, h->sm_count, h->smem_limit, st));
                } else {
                    ProfScope ps(h, PROF_ATTN_BWD, st);
                    MSA_TRY(launch_attn_chain_bwd(make(g0), h->sm_count, h->smem_limit, st));
                }
            }
        }
    }
    // ---- stage 3: deferred attention / prenet / speaker gradients ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < NG; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        const float* Wia = P("decoder.attention_rnn.weight_ih");
        (void)Wia;
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        float* grads = grads_g[g];
        auto mk = [&](int i) { return h->masks[g] + secs[i].off; };
        auto G = [&](const std::string& n) { return grads + h->off(n); };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        (void)mk; (void)G;
        // ---- deferred attention / attention-RNN parameter gradients ----
        float* gWia = G("decoder.attention_rnn.weight_ih");
        if (ta) {
            // transition agent u(t) = sigmoid(W_ta . [ctx(t); h_a'(t)] + b) (forward_attn.py:222-224), dzu [T][B] from the chain
            float* gWta = G(at + "ta.weight");
            MSA_TRY(gemm(h, true, false, 1, d.E, d.TB, gs, w.dzu, 1, w.ctx, d.E, beta, gWta, d.E + d.Ha));
            MSA_TRY(gemm(h, true, false, 1, d.Ha, d.TB, gs, w.dzu, 1, w.ha, d.Ha, beta, gWta + d.E, d.E + d.Ha));
            MSA_TRY(k_colsum(w.dzu, d.TB, 1, 1, G(at + "ta.bias"), gs, acc, nullptr, w.red_scr, st));
            // d memory += sum_t alpha(t)^T . (dzu(t) (x) W_ta[:E])   (w.dctx is free after the context backward above)
            MSA_TRY(gemm(h, false, false, d.TB, d.E, 1, 1.f, w.dzu, 1, P(at + "ta.weight"), d.E + d.Ha, 0.f, w.dctx, d.E));
            MSA_TRY(gemm_batched(h, true, false, L, d.E, T, 1.f, w.align_tm, d.BL, L, w.dctx, (int64_t)B * d.E, d.E, 1.f, w.dmem, d.E,
                                 (int64_t)L * d.E, B));
        }
        const int64_t nfr = d.TB + B;
        MSA_TRY(gemm(h, false, false, d.TB, d.Pd, H4a, 1.f, w.dza, H4a, Wia, ldA, 0.f, w.dxp, d.Pd));
        MSA_CUDA(cudaMemsetAsync(w.dxp + d.TB * d.Pd, 0, sizeof(float) * (size_t)B * d.Pd, st));   // last prenet frame is unused (decoder.py:305)
        MSA_TRY(gemm(h, true, false, H4a, d.Pd, d.TB, gs, w.dza, H4a, w.xpre, d.Pd, beta, gWia, ldA));
        float* gWha = G("decoder.attention_rnn.weight_hh");
        if (T > 1) {
            MSA_TRY(gemm(h, true, false, H4a, d.Ha, d.TB - B, gs, w.dza + (int64_t)B * H4a, H4a, w.ha, d.Ha, beta, gWha, d.Ha));
            // dMW[b][l][r] = sum_t a[t][b][l] * dz_a[t+1][b][r]
            MSA_TRY(gemm_batched(h, true, false, L, H4a, T - 1, 1.f, w.align_tm, d.BL, L, w.dza + (int64_t)B * H4a, (int64_t)B * H4a, H4a,
                                 0.f, w.dmw, H4a, (int64_t)L * H4a, B));
        } else {
            if (!acc) MSA_CUDA(cudaMemsetAsync(gWha, 0, sizeof(float) * (size_t)H4a * d.Ha, st));
            MSA_CUDA(cudaMemsetAsync(w.dmw, 0, sizeof(float) * (size_t)d.BL * H4a, st));
        }
        MSA_TRY(k_colsum(w.dza, d.TB, H4a, H4a, G("decoder.attention_rnn.bias_ih"), gs, acc, G("decoder.attention_rnn.bias_hh"), w.red_scr, st));
        MSA_TRY(gemm(h, true, false, H4a, d.E, d.BL, gs, w.dmw, H4a, w.memory, d.E, beta, gWia + d.Pd, ldA));
        MSA_TRY(gemm(h, false, false, d.BL, d.E, H4a, 1.f, w.dmw, H4a, Wia + d.Pd, ldA, 1.f, w.dmem, d.E));
        MSA_TRY(gemm(h, true, false, d.A, d.Ha, d.TB, gs, w.dq, d.A, w.ha, d.Ha, beta, G(at + "query_layer.linear_layer.weight"), d.Ha));
        MSA_TRY(k_sum_over_t(w.ds, w.dpm, T, d.BL * d.A, st));
        MSA_TRY(gemm(h, true, false, d.A, d.E, d.BL, gs, w.dpm, d.A, w.memory, d.E, beta, G(at + "inputs_layer.linear_layer.weight"), d.E));
        MSA_TRY(gemm(h, false, false, d.BL, d.E, d.A, 1.f, w.dpm, d.A, P(at + "inputs_layer.linear_layer.weight"), d.E, 1.f, w.dmem, d.E));
        MSA_TRY(gemm(h, true, false, 1, d.A, d.TBL, gs, w.de, 1, w.s, d.A, beta, G(at + "v.linear_layer.weight"), d.A));
        MSA_TRY(k_dot_rows(w.de, nullptr, d.TBL, w.loss_part, G(at + "v.linear_layer.bias"), gs, acc, st));
        MSA_TRY(gemm(h, true, false, d.A, d.F, d.TBL, gs, w.ds, d.A, w.convf, d.F, beta,
                     G(at + "location_layer.location_dense.linear_layer.weight"), d.F));
        MSA_TRY(k_wloc_grad(w.dconvf, w.align_tm, w.cum, G(at + "location_layer.location_conv1d.weight"), w.wloc_part, T, B, L, d.F, d.Kl, gs,
                            acc, st));
        // ---- prenet backward (decoder.py:9-20) ----
        MSA_TRY(k_relu_drop_bwd(w.dxp, w.xpre, mk(iPre + 1), 2.f, nfr * d.Pd, st));
        MSA_TRY(gemm(h, true, false, d.Pd, d.Pd, nfr, gs, w.dxp, d.Pd, w.p1, d.Pd, beta, G("decoder.prenet.layers.1.linear_layer.weight"), d.Pd));
        MSA_TRY(gemm(h, false, false, nfr, d.Pd, d.Pd, 1.f, w.dxp, d.Pd, P("decoder.prenet.layers.1.linear_layer.weight"), d.Pd, 0.f, w.dp1, d.Pd));
        MSA_TRY(k_relu_drop_bwd(w.dp1, w.p1, mk(iPre), 2.f, nfr * d.Pd, st));
        MSA_TRY(gemm(h, true, false, d.Pd, d.M, nfr, gs, w.dp1, d.Pd, w.frames, d.M, beta, G("decoder.prenet.layers.0.linear_layer.weight"), d.M));
        // ---- memory -> encoder output + speaker path (tacotron2nv.py:104-111) ----
        MSA_TRY(k_split_dmemory(w.dmem, w.denc_h, c.spk_mode ? w.dspk : nullptr, B, L, d.Hh, d.Ds, st));
        if (c.spk_mode == 1) {
            MSA_TRY(gemm(h, true, false, d.Ds, d.Dsin, B, gs, w.dspk, d.Ds, h->spk_in[g], d.Dsin, beta, G("speaker_lin.weight"), d.Dsin));
            MSA_TRY(k_colsum(w.dspk, B, d.Ds, d.Ds, G("speaker_lin.bias"), gs, acc, nullptr, w.red_scr, st));
        } else if (c.spk_mode == 2) {
            MSA_TRY(k_embedding_bwd(w.dspk, h->spk_ids[g], G("speaker_embedder.weight"), B, d.Ds, c.num_speakers, gs, acc, st));
        }
    }
    MSA_TRY(fk.end(W));
    mark_decoder_done();
    if (c.freeze_encoder) {      // tacotron2nv.py:99-101: the encoder output (incl. the residual term) is detached
        for (int g = 0; g < NG; ++g) MSA_TRY(zero_grads(h, grads_g[g], {"encoder.", "embedding."}, false, acc, st));
        return 0;
    }
    // ---- encoder BiLSTM backward (persistent) ----
    {
        auto make = [&](int g) {
            const Ws& w = W[g];
            LstmRecBwdParams bp{};
            bp.T = L; bp.B = B; bp.H = d.Hh; bp.ndir = 2;
            bp.whh = PG(g, "encoder.lstm.weight_hh_l0");
            bp.whh_dir_stride = h->off("encoder.lstm.weight_hh_l0_reverse") - h->off("encoder.lstm.weight_hh_l0");
            bp.gates = w.enc_g; bp.cout = w.enc_c; bp.dh_ext = w.denc_h; bp.dz = w.dzx; bp.mask = nullptr; bp.drop_scale = 1.f;
            bp.lengths = h->tok_len[g]; bp.abort_word = h->abort_dev; bp.prof = prof_ptr(h, w, PROF_ENC_LSTM_BWD);
            bp.trace = trace_ptr(h, w, PROF_ENC_LSTM_BWD); bp.trace_t0 = h->trace_t0; bp.flags = h->rec_flags >= 0 ? h->rec_flags : 1;
            return bp;
        };
        if (mma && shared) {
            ProfScope ps(h, PROF_ENC_LSTM_BWD + (NG > 1 ? PROF_BASE : 0), st);
            LstmRecBwdParams bp = make(0);
            bp.G = NG; bp.tstride = tstride;
            for (int g = 0; g < NG; ++g) bp.lengths_g[g] = h->tok_len[g];
            MSA_TRY(launch_lstm_rec_bwd_mma(bp, h->sm_count, h->smem_limit, st));
        } else if (mma) {       // per-task weights: one launch per task (G = 1, that task's pointers)
            for (int g = 0; g < NG; ++g) {
                ProfScope ps(h, PROF_ENC_LSTM_BWD, st);
                LstmRecBwdParams bp = make(g);
                bp.G = 1; bp.tstride = 0;
                MSA_TRY(launch_lstm_rec_bwd_mma(bp, h->sm_count, h->smem_limit, st));
            }
        } else {
            for (int g = 0; g < NG; ++g) {
                ProfScope ps(h, PROF_ENC_LSTM_BWD, st);
                MSA_TRY(launch_lstm_rec_bwd(make(g), h->sm_count, h->smem_limit, st));
            }
        }
    }
    // ---- stage 4: encoder BiLSTM weight gradients, encoder convolutions, embedding ----
    MSA_TRY(fk.begin());
    for (int g = 0; g < NG; ++g) {
        const Ws& w = W[g];
        MSA_TRY(fk.enter(g));
        cudaStream_t st = fk.stream(g);      // this task's stage stream (shadows the pass stream)
        const float* params = params_g[g];
        auto P = [&](const std::string& n) { return params + h->off(n); };
        const float* Wia = P("decoder.attention_rnn.weight_ih");
        (void)Wia;
        MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
        float* grads = grads_g[g];
        auto mk = [&](int i) { return h->masks[g] + secs[i].off; };
        auto G = [&](const std::string& n) { return grads + h->off(n); };
        h->gemm_scratch = w.gemm_lo; h->gemm_scratch_floats = (size_t)w.n_gemm_lo;
        (void)mk; (void)G;
        for (int dir = 0; dir < 2; ++dir) {
            const std::string sfx = dir ? "_reverse" : "";
            const float* dzx = w.dzx + (int64_t)dir * d.BL * H4e;
            const float* eh = w.enc_h + (int64_t)dir * d.BL * d.Hh;
            MSA_TRY(gemm(h, false, false, d.BL, d.C, H4e, 1.f, dzx, H4e, P("encoder.lstm.weight_ih_l0" + sfx), d.C, dir ? 1.f : 0.f, w.dx3_tm, d.C));
            MSA_TRY(gemm(h, true, false, H4e, d.C, d.BL, gs, dzx, H4e, w.x3_tm, d.C, beta, G("encoder.lstm.weight_ih_l0" + sfx), d.C));
            float* gWhh = G("encoder.lstm.weight_hh_l0" + sfx);
            if (L > 1) {
                if (dir == 0)   // z(t) uses h(t-1)
                    MSA_TRY(gemm(h, true, false, H4e, d.Hh, d.BL - B, gs, dzx + (int64_t)B * H4e, H4e, eh, d.Hh, beta, gWhh, d.Hh));
                else            // reverse direction: z(t) uses h(t+1)
                    MSA_TRY(gemm(h, true, false, H4e, d.Hh, d.BL - B, gs, dzx, H4e, eh + (int64_t)B * d.Hh, d.Hh, beta, gWhh, d.Hh));
            } else if (!acc) {
                MSA_CUDA(cudaMemsetAsync(gWhh, 0, sizeof(float) * (size_t)H4e * d.Hh, st));
            }
            MSA_TRY(k_colsum(dzx, d.BL, H4e, H4e, G("encoder.lstm.bias_ih_l0" + sfx), gs, acc, G("encoder.lstm.bias_hh_l0" + sfx), w.red_scr, st));
        }
        // ---- encoder convolutions backward ----
        const int64_t ex = d.BL * d.C;
        float* epp[2] = {w.edx0, w.edx1};
        MSA_TRY(k_transpose01(w.dx3_tm, epp[d.nEnc & 1], L, B, d.C, st));   // [L][B][C] -> [B][L][C]
        const float* dcur = epp[d.nEnc & 1];
        for (int i = d.nEnc - 1; i >= 0; --i) {
            float* dx = epp[i & 1];
            MSA_TRY(conv_bn_bwd(h, st, params, grads, gs, acc, "encoder.convolutions." + std::to_string(i), w.enc_x + i * ex,
                                w.enc_y + i * ex, dcur, dx, w.edy, d.conv_tc ? nullptr : w.enc_col + i * d.BL * d.Kc * d.C,
                                d.conv_tc ? w.enc_wp + (int64_t)i * d.Kc * d.C * d.C : nullptr, w.edcol, w.bn_scr, w.enc_bn + i * 2 * d.C, w.enc_bn + i * 2 * d.C + d.C, B, L, d.C, d.C, d.Kc, 1, mk(i),
                                true, w.red_scr));
            dcur = dx;
        }
        if (c.freeze_charemb) {      // tacotron2nv.py:90-92: the embedded characters are detached (encoder input and residual term)
            MSA_TRY(zero_grads(h, grads, {"embedding."}, false, acc, st));
            continue;
        }
        if (c.residual_encoder)      // d(embedded) += d(encoder output) = the encoder columns of d(memory)
            MSA_TRY(k_add_cols(const_cast<float*>(dcur), w.dmem, d.BL, d.C, d.E, st));
        MSA_TRY(k_embedding_bwd(dcur, h->tokens[g], G("embedding.weight"), (int)d.BL, d.C, c.n_symbols, gs, acc, st));
    }
    MSA_TRY(fk.end(W));
    return 0;
}

}  // namespace msa

extern "C" {

int msa_train_backward(msa_handle* h, void* wsp, size_t ws_bytes, const float* params, const float* d_mel, const float* d_mel_post,
                       const float* d_gate, float* grads, int acc, float gs, void* stream) {
    MSA_CHECK(h && wsp && params && grads, MSA_E_ARG, "msa_train_backward: null argument");
    MSA_CHECK(h->fwd_valid, MSA_E_STATE, "msa_train_backward: call msa_train_forward first");
    MSA_CHECK(h->G == 1, MSA_E_STATE, "msa_train_backward: the last forward was a grouped one (use msa_train_backward_group)");
    MSA_CHECK(ws_bytes >= ws_layout(h->d, nullptr).total_bytes, MSA_E_WORKSPACE, "msa_train_backward: workspace too small");
    const bool ext = d_mel || d_mel_post || d_gate;
    MSA_CHECK(!ext || (d_mel && d_mel_post && d_gate), MSA_E_ARG, "msa_train_backward: pass all three upstream gradients or none");
    return train_backward_impl(h, static_cast<char*>(wsp), &params, ext ? &d_mel : nullptr, ext ? &d_mel_post : nullptr,
                               ext ? &d_gate : nullptr, &grads, acc, gs, (cudaStream_t)stream);
}

int msa_train_backward_group(msa_handle* h, void* wsp, size_t ws_bytes, const float* const* params, float* const* grads, int acc, float gs,
                             void* stream) {
    MSA_CHECK(h && wsp && params && grads, MSA_E_ARG, "msa_train_backward_group: null argument");
    MSA_CHECK(h->fwd_valid, MSA_E_STATE, "msa_train_backward_group: call msa_train_forward_group first");
    MSA_CHECK(ws_bytes >= (size_t)h->G * std::max<size_t>(h->ws_stride, 1), MSA_E_WORKSPACE, "msa_train_backward_group: workspace too small");
    for (int g = 0; g < h->G; ++g) MSA_CHECK(grads[g] != nullptr && params[g] != nullptr, MSA_E_ARG, "msa_train_backward_group: null buffer of task %d", g);
    return train_backward_impl(h, static_cast<char*>(wsp), params, nullptr, nullptr, nullptr, grads, acc, gs, (cudaStream_t)stream);
}

int msa_backward_mark_event(msa_handle* h, void* cuda_event, int64_t* prefix_floats) {
    MSA_CHECK(h, MSA_E_ARG, "msa_backward_mark_event: null handle");
    if (prefix_floats) {      // the encoder part is the prefix of the flat layout: embedding.weight, encoder.*
        int64_t end = 0;
        for (size_t i = 0; i < h->names.size(); ++i) {
            const std::string& n = h->names[i];
            if (n.compare(0, 10, "embedding.") != 0 && n.compare(0, 8, "encoder.") != 0) break;
            end = i + 1 < h->names.size() ? h->offs[i + 1] : h->total;
        }
        *prefix_floats = end;
    }
    h->bwd_ev = static_cast<cudaEvent_t>(cuda_event);
    return 0;
}

int msa_check_abort(msa_handle* h, void* wsp, void* stream) {
    (void)wsp;      // the word lives in the handle since round 2; the argument is kept for ABI stability
    MSA_CHECK(h, MSA_E_ARG, "msa_check_abort: null handle");
    unsigned int flag = 0;
    MSA_CUDA(cudaMemcpyAsync(&flag, h->abort_dev, sizeof(flag), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MSA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    MSA_CHECK(flag == 0, MSA_E_STATE, "a persistent kernel gave up waiting for data of another CTA (polling time-out): results are invalid");
    return 0;
}
int msa_abort_clear(msa_handle* h, void* stream) {
    MSA_CHECK(h, MSA_E_ARG, "msa_abort_clear: null handle");
    MSA_CUDA(cudaMemsetAsync(h->abort_dev, 0, 256, (cudaStream_t)stream));
    return 0;
}
int msa_abort_guard(msa_handle* h, float* sumsq, void* stream) {
    MSA_CHECK(h && sumsq, MSA_E_ARG, "msa_abort_guard: null argument");
    return k_abort_guard(h->abort_dev, sumsq, 0, (cudaStream_t)stream);
}
int msa_abort_read_async(msa_handle* h, uint32_t* host_flag, void* stream) {
    MSA_CHECK(h && host_flag, MSA_E_ARG, "msa_abort_read_async: null argument");
    MSA_CUDA(cudaMemcpyAsync(host_flag, h->abort_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}
int msa_debug_raise_abort(msa_handle* h, void* stream) {
    MSA_CHECK(h, MSA_E_ARG, "msa_debug_raise_abort: null handle");
    return k_abort_guard(h->abort_dev, nullptr, 1, (cudaStream_t)stream);
}

int msa_profile_trace(msa_handle* h, void* wsp, int id, int64_t* out, int ncta) {
    MSA_CHECK(h && wsp && out && h->fwd_valid, MSA_E_STATE, "msa_profile_trace: no forward pass in this workspace");
    MSA_CHECK(id >= 0 && id < PROF_N && ncta >= 1 && ncta <= 256, MSA_E_ARG, "msa_profile_trace: bad kernel id / cta count");
    id = prof_base_id(id);
    const Ws w = ws_layout(h->d, wsp);
    MSA_CUDA(cudaDeviceSynchronize());
    MSA_CUDA(cudaMemcpy(out, reinterpret_cast<const long long*>(w.trace) + (size_t)id * kTraceWordsPerKernel,
                        sizeof(int64_t) * (size_t)ncta * 16 * 4 * 12 * 2, cudaMemcpyDeviceToHost));
    return 0;
}
int msa_profile_trace_step(msa_handle* h, int t0) {
    MSA_CHECK(h, MSA_E_ARG, "msa_profile_trace_step: null handle");
    h->trace_t0 = t0;
    return 0;
}

int msa_profile_phases(msa_handle* h, void* wsp, int id, int64_t* out, int ncta) {
    MSA_CHECK(h && wsp && out && h->fwd_valid, MSA_E_STATE, "msa_profile_phases: no forward pass in this workspace");
    MSA_CHECK(id >= 0 && id < PROF_N && ncta >= 1 && ncta <= kProfCtas, MSA_E_ARG, "msa_profile_phases: bad kernel id / cta count");
    id = prof_base_id(id);
    const Ws w = ws_layout(h->d, wsp);
    MSA_CUDA(cudaDeviceSynchronize());
    MSA_CUDA(cudaMemcpy(out, reinterpret_cast<const long long*>(w.prof) + (size_t)id * kProfCtas * kProfSlots,
                        sizeof(int64_t) * (size_t)ncta * kProfSlots, cudaMemcpyDeviceToHost));
    return 0;
}

int msa_get_buffer(msa_handle* h, void* wsp, const char* name, void** ptr, int64_t* numel) {
    MSA_CHECK(h && wsp && name && h->fwd_valid, MSA_E_STATE, "msa_get_buffer: no forward pass in this workspace");
    const Ws w = ws_layout(h->d, wsp);
#define X(nm, n)                                   \
    if (!strcmp(name, #nm)) {                      \
        if (ptr) *ptr = w.nm;                      \
        if (numel) *numel = w.n_##nm;              \
        return 0;                                  \
    }
    WS_LIST(X)
#undef X
    MSA_CHECK(false, MSA_E_ARG, "msa_get_buffer: unknown buffer '%s'", name);
    return 0;
}

}  // extern "C"

// =====================================================================================================
// free-running inference (Tacotron2NV.infer, models/tacotron2nv.py:130-162; Decoder.infer, decoder.py:334-411)
namespace msa {

#define IWS_LIST(X)                                                                                       \
    X(spk_vec, d.B * d.Ds) X(enc_x, 2 * d.BL * d.C) X(enc_y, d.BL * d.C) X(enc_bn, 2 * std::max(d.C, d.Cmax)) \
    X(enc_col, d.BL * d.Kc * d.C) X(x3_tm, d.BL * d.C)              \
    X(enc_zx, 2 * d.BL * 4 * d.Hh) X(enc_g, 2 * d.BL * 4 * d.Hh) X(enc_c, 2 * d.BL * d.Hh)               \
    X(enc_h, 2 * d.BL * d.Hh) X(memory, d.BL * d.E) X(pm, d.BL * d.A)                                    \
    X(xin_a, d.B * (d.Pd + d.E)) X(p1, d.B * d.Pd) X(ha, 2 * d.B * d.Ha) X(ca, d.B * d.Ha)               \
    X(xin_d, d.B * (d.Ha + d.E)) X(hd, 2 * d.B * d.Hd) X(cd, d.B * d.Hd) X(xin_p, d.B * (d.Hd + d.E))    \
    X(mel_raw, d.B * d.M) X(gate_raw, d.B) X(frame, d.B * d.M) X(proj_part, 16 * d.B * (d.M + 1))        \
    X(prev, d.BL) X(cum, d.BL) X(fa_alpha, d.BL) X(fa_u, d.B + 4) X(attn_wloc_t, 2 * d.Kl * d.F + 4)      \
    X(attn_wld4, ((d.F + 3) / 4) * d.A * 4) X(mel_tm, (int64_t)d.T * d.B * d.M) X(ints, 64 + d.B)                      \
    X(post_x, 2 * d.BT * d.Cmax) X(post_y, d.BT * d.Cmax) X(post_col, d.BT * d.Kp * d.Cmax)              \
    X(post_bt, d.BT * d.M)

struct IWs {
#define X(name, n) float* name; int64_t n_##name;
    IWS_LIST(X)
#undef X
    unsigned int* abort_word;
    void* blas_ws;
    size_t blas_ws_bytes;
    size_t total_bytes;
};
static IWs iws_layout(const Dims& d, void* base) {
    IWs w;
    size_t o = 0;
    char* b = static_cast<char*>(base);
#define X(name, n)                                        \
    w.n_##name = (int64_t)(n);                            \
    w.name = reinterpret_cast<float*>(b + o);             \
    o += (size_t)align_up((int64_t)(n), 64) * sizeof(float);
    IWS_LIST(X)
#undef X
    w.abort_word = reinterpret_cast<unsigned int*>(b + o);
    o += 256;
    w.blas_ws = b + o;
    w.blas_ws_bytes = kBlasWs;
    o += kBlasWs;
    w.total_bytes = o;
    return w;
}

// conv1d ("same") + BatchNorm(eval: running statistics) + activation, channels-last rows = B*Tn (no dropout in eval)
static int conv_bn_eval(msa_handle* h, cudaStream_t st, const float* params, const std::string& pfx, const float* x, float* y,
                        float* xout, float* col, float* bn_mean, float* bn_invstd, const float* running, int B, int Tn,
                        int Ci, int Co, int K, int act) {
    const int64_t rows = (int64_t)B * Tn;
    MSA_TRY(k_im2col(x, col, B, Tn, Ci, K, st));
    MSA_TRY(k_fill_rows(y, params + h->off(pfx + ".0.conv.bias"), nullptr, rows, Co, st));
    MSA_TRY(gemm(h, false, true, rows, Co, (int64_t)K * Ci, 1.f, col, (int64_t)K * Ci, params + h->off(pfx + ".0.conv.weight"),
                 (int64_t)K * Ci, 1.f, y, Co));
    MSA_TRY(k_bn_eval_stats(running, Co, (int)align_up(Co), bn_mean, bn_invstd, st));
    MSA_TRY(k_bn_act_drop_fwd(y, bn_mean, bn_invstd, params + h->off(pfx + ".1.weight"), params + h->off(pfx + ".1.bias"), nullptr,
                              1.0f, act, xout, rows, Co, st));
    return 0;
}

static int infer_check(const msa_handle* h) {
    const msa_config& c = h->cfg;
    MSA_CHECK(!c.trans_agent || c.forward_attn, MSA_E_ARG, "msa_infer: trans_agent needs forward_attn (forward_attn.py:222)");
    return 0;
}

}  // namespace msa

extern "C" {

size_t msa_infer_workspace_bytes(const msa_handle* h, int B, int L, int max_steps) {
    if (!h || B <= 0 || L <= 0 || max_steps <= 0) return 0;
    return iws_layout(make_dims(h->cfg, B, max_steps, L), nullptr).total_bytes;
}

int msa_infer(msa_handle* h, void* wsp, size_t ws_bytes, const float* params, const float* bn_stats, const int64_t* tokens,
              const int64_t* token_lengths, const float* speaker_vecs, const int64_t* speaker_ids, const uint8_t* prenet_masks,
              int B, int L, int max_steps, float* mel_post_out, int32_t* mel_lengths_out, float* align_out, int32_t* n_steps_out,
              void* stream) {
    MSA_CHECK(h && wsp && params && bn_stats && tokens && token_lengths && prenet_masks && mel_post_out && mel_lengths_out &&
              align_out && n_steps_out, MSA_E_ARG, "msa_infer: null argument");
    MSA_CHECK(B >= 1 && L >= 1 && max_steps >= 1, MSA_E_ARG, "msa_infer: bad dims B=%d L=%d max_steps=%d", B, L, max_steps);
    MSA_CHECK(h->cfg.spk_mode == 2 ? speaker_ids != nullptr : speaker_vecs != nullptr, MSA_E_ARG, "msa_infer: speaker input missing");
    MSA_TRY(infer_check(h));
    const Dims d = make_dims(h->cfg, B, max_steps, L);
    const IWs w = iws_layout(d, wsp);
    MSA_CHECK(ws_bytes >= w.total_bytes, MSA_E_WORKSPACE, "msa_infer: workspace %zu < %zu bytes", ws_bytes, w.total_bytes);
    MSA_CHECK(((uintptr_t)wsp & 255) == 0, MSA_E_ARG, "msa_infer: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    MSA_CUDA(cudaSetDevice(h->device));
    MSA_BLAS(cublasSetStream(h->blas, st));
    MSA_BLAS(cublasSetWorkspace(h->blas, w.blas_ws, w.blas_ws_bytes));
    h->in_bwd = false;
    h->fwd_valid = false;
    h->gemm_scratch = nullptr; h->gemm_scratch_floats = 0; h->cur_stream = st;      // encoder / postnet GEMMs of inference stay on cuBLAS
    const msa_config& c = h->cfg;
    auto P = [&](const std::string& n) { return params + h->off(n); };

    // ---- speaker vector + encoder in eval mode (tacotron2nv.py:137-148, encoder.py:55-71) ----
    if (c.spk_mode == 0) {
        MSA_TRY(k_scale_copy(speaker_vecs, w.spk_vec, (int64_t)B * d.Ds, 1.f, 0, st));
    } else if (c.spk_mode == 1) {
        MSA_TRY(k_fill_rows(w.spk_vec, P("speaker_lin.bias"), nullptr, B, d.Ds, st));
        MSA_TRY(gemm(h, false, true, B, d.Ds, d.Dsin, 1.f, speaker_vecs, d.Dsin, P("speaker_lin.weight"), d.Dsin, 1.f, w.spk_vec, d.Ds));
    } else {
        MSA_TRY(k_embedding_fwd(P("speaker_embedder.weight"), speaker_ids, w.spk_vec, B, d.Ds, c.num_speakers, st));
    }
    const int64_t ex = d.BL * d.C;
    MSA_TRY(k_embedding_fwd(P("embedding.weight"), tokens, w.enc_x, (int)d.BL, d.C, c.n_symbols, st));
    for (int i = 0; i < d.nEnc; ++i)
        MSA_TRY(conv_bn_eval(h, st, params, "encoder.convolutions." + std::to_string(i), w.enc_x + (i & 1) * ex, w.enc_y,
                             w.enc_x + ((i + 1) & 1) * ex, w.enc_col, w.enc_bn, w.enc_bn + std::max(d.C, d.Cmax),
                             bn_stats + h->bn_offs[i], B, L, d.C, d.C, d.Kc, 1));
    MSA_TRY(k_transpose01(w.enc_x + (d.nEnc & 1) * ex, w.x3_tm, B, L, d.C, st));
    const int H4e = 4 * d.Hh;
    for (int dir = 0; dir < 2; ++dir) {
        const std::string sfx = dir ? "_reverse" : "";
        float* zx = w.enc_zx + (int64_t)dir * d.BL * H4e;
        MSA_TRY(k_fill_rows(zx, P("encoder.lstm.bias_ih_l0" + sfx), P("encoder.lstm.bias_hh_l0" + sfx), d.BL, H4e, st));
        MSA_TRY(gemm(h, false, true, d.BL, H4e, d.C, 1.f, w.x3_tm, d.C, P("encoder.lstm.weight_ih_l0" + sfx), d.C, 1.f, zx, H4e));
    }
    {
        LstmRecParams lp{};
        lp.T = L; lp.B = B; lp.H = d.Hh; lp.ndir = 2;
        lp.zin = w.enc_zx; lp.whh = P("encoder.lstm.weight_hh_l0");
        lp.whh_dir_stride = h->off("encoder.lstm.weight_hh_l0_reverse") - h->off("encoder.lstm.weight_hh_l0");
        lp.hout = w.enc_h; lp.cout = w.enc_c; lp.gates = w.enc_g; lp.mask = nullptr; lp.drop_scale = 1.f;
        lp.lengths = token_lengths; lp.abort_word = h->abort_dev; lp.prof = nullptr; lp.trace = nullptr; lp.trace_t0 = 0;
        lp.flags = h->rec_flags >= 0 ? h->rec_flags : 1;
        MSA_TRY(launch_lstm_rec_fwd(lp, h->sm_count, h->smem_limit, st));
    }
    MSA_TRY(k_build_memory(w.enc_h, w.spk_vec, w.memory, B, L, d.Hh, d.Ds, st));
    if (c.residual_encoder) MSA_TRY(k_embedding_add(P("embedding.weight"), tokens, w.memory, (int)d.BL, d.C, d.E, c.n_symbols, st));
    const std::string at = "decoder.attention_layer.";
    MSA_TRY(gemm(h, false, true, d.BL, d.A, d.E, 1.f, w.memory, d.E, P(at + "inputs_layer.linear_layer.weight"), d.E, 0.f, w.pm, d.A));

    // ---- decoder state (decoder.py:337-360, forward_attn.py:103-116) ----
    const int KA = d.Pd + d.E, KD = d.Ha + d.E, KP = d.Hd + d.E;
    int* ints = reinterpret_cast<int*>(w.ints);
    int* state = ints;                 // [0] step, [1] done, [2] steps produced
    int* not_finished = ints + 64;     // [B]
    MSA_CUDA(cudaMemsetAsync(w.xin_a, 0, sizeof(float) * (size_t)B * KA, st));
    MSA_CUDA(cudaMemsetAsync(w.ha, 0, sizeof(float) * (size_t)2 * B * d.Ha, st));
    MSA_CUDA(cudaMemsetAsync(w.ca, 0, sizeof(float) * (size_t)B * d.Ha, st));
    MSA_CUDA(cudaMemsetAsync(w.xin_d, 0, sizeof(float) * (size_t)B * KD, st));
    MSA_CUDA(cudaMemsetAsync(w.hd, 0, sizeof(float) * (size_t)2 * B * d.Hd, st));
    MSA_CUDA(cudaMemsetAsync(w.cd, 0, sizeof(float) * (size_t)B * d.Hd, st));
    MSA_CUDA(cudaMemsetAsync(w.xin_p, 0, sizeof(float) * (size_t)B * KP, st));
    MSA_CUDA(cudaMemsetAsync(w.frame, 0, sizeof(float) * (size_t)B * d.M, st));
    MSA_CUDA(cudaMemsetAsync(w.prev, 0, sizeof(float) * (size_t)d.BL, st));
    MSA_CUDA(cudaMemsetAsync(w.cum, 0, sizeof(float) * (size_t)d.BL, st));
    MSA_CUDA(cudaMemsetAsync(ints, 0, sizeof(int) * 64, st));
    MSA_CUDA(cudaMemsetAsync(mel_lengths_out, 0, sizeof(int32_t) * (size_t)B, st));
    MSA_TRY(k_fill_ones_i32(not_finished, B, st));
    int* win = ints + 16;                                   // [2] window index by step parity
    float* gmax = reinterpret_cast<float*>(ints + 18);
    MSA_TRY(k_init_fwd_attn(w.fa_alpha, w.fa_u, win, gmax, B, L, st));
    MSA_TRY(k_infer_attn_prep(P(at + "location_layer.location_conv1d.weight"), P(at + "location_layer.location_dense.linear_layer.weight"),
                              w.attn_wloc_t, w.attn_wld4, d.F, d.Kl, d.A, st));
    MSA_CUDA(cudaMemsetAsync(align_out, 0, sizeof(float) * (size_t)B * max_steps * L, st));
    MSA_CUDA(cudaMemsetAsync(mel_post_out, 0, sizeof(float) * (size_t)B * d.M * max_steps, st));

    // ---- one decoder step (decoder.py:362-399 -> decode 234-274) = six fused launches (infer_decode.cu); the step index
    // lives on the device.  The recurrent h of both LSTMCells ping-pongs between two buffers by step parity (a CTA may not
    // overwrite h(t-1) while another CTA still reads it); a second copy goes into the packed input of the next product. ----
    const float* Wia = P("decoder.attention_rnn.weight_ih");
    const float* Wid = P("decoder.decoder_rnn.weight_ih");
    unsigned int* counter = reinterpret_cast<unsigned int*>(ints + 8);
    // TMA path of the two LSTMCells: tensor maps of the (fixed) weight matrices and of the packed / ping-pong input buffers
    static const bool tma_env = !(getenv("MSA_INFER_TMA") && atoi(getenv("MSA_INFER_TMA")) == 0);
    const bool use_tma = tma_env && infer_lstm_tma_supported(d.Ha, KA, KA, d.Ha, d.Ha, h->sm_count) &&
                         infer_lstm_tma_supported(d.Hd, KD, KD, d.Hd, d.Hd, h->sm_count);
    alignas(64) unsigned char tmaps[10][128];
    if (use_tma) {
        MSA_TRY(infer_lstm_tma_map_x(tmaps[0], w.xin_a, B, KA, KA));
        MSA_TRY(infer_lstm_tma_map_x(tmaps[1], w.ha, B, d.Ha, d.Ha));
        MSA_TRY(infer_lstm_tma_map_x(tmaps[2], w.ha + (size_t)B * d.Ha, B, d.Ha, d.Ha));
        MSA_TRY(infer_lstm_tma_map_w(tmaps[3], Wia, d.Ha, KA, KA));
        MSA_TRY(infer_lstm_tma_map_w(tmaps[4], P("decoder.attention_rnn.weight_hh"), d.Ha, d.Ha, d.Ha));
        MSA_TRY(infer_lstm_tma_map_x(tmaps[5], w.xin_d, B, KD, KD));
        MSA_TRY(infer_lstm_tma_map_x(tmaps[6], w.hd, B, d.Hd, d.Hd));
        MSA_TRY(infer_lstm_tma_map_x(tmaps[7], w.hd + (size_t)B * d.Hd, B, d.Hd, d.Hd));
        MSA_TRY(infer_lstm_tma_map_w(tmaps[8], Wid, d.Hd, KD, KD));
        MSA_TRY(infer_lstm_tma_map_w(tmaps[9], P("decoder.decoder_rnn.weight_hh"), d.Hd, d.Hd, d.Hd));
    }
    auto step = [&](int s, bool first) -> int {
        const int cur = s & 1, nxt = cur ^ 1;
        // prenet, dropout always on (decoder.py:9-20,366): both layers in one cluster launch
        static const bool pn_env = !(getenv("MSA_INFER_PRENET") && atoi(getenv("MSA_INFER_PRENET")) == 0);
        if (pn_env && infer_prenet_supported(d.M, d.Pd)) {
            InferPrenetParams pn{};
            pn.B = B; pn.M = d.M; pn.Pd = d.Pd; pn.frame = w.frame; pn.w1 = P("decoder.prenet.layers.0.linear_layer.weight");
            pn.w2 = P("decoder.prenet.layers.1.linear_layer.weight"); pn.mask = prenet_masks; pn.out = w.xin_a; pn.ldo = KA;
            pn.state = state;
            MSA_TRY(k_infer_prenet(pn, st));
        } else {
            InferRowsParams rp{};
            rp.B = B; rp.state = state;
            rp.N = d.Pd; rp.nseg = 1; rp.epi = IR_EPI_RELU_DROP; rp.mask = prenet_masks;
            rp.x[0] = w.frame; rp.ldx[0] = d.M; rp.K[0] = d.M; rp.W[0] = P("decoder.prenet.layers.0.linear_layer.weight"); rp.ldw[0] = d.M;
            rp.out = w.p1; rp.ldo = d.Pd; rp.mask_layer = 0;
            MSA_TRY(k_infer_rows(rp, h->sm_count, st));
            rp.x[0] = w.p1; rp.ldx[0] = d.Pd; rp.K[0] = d.Pd; rp.W[0] = P("decoder.prenet.layers.1.linear_layer.weight"); rp.ldw[0] = d.Pd;
            rp.out = w.xin_a; rp.ldo = KA; rp.mask_layer = 1;
            MSA_TRY(k_infer_rows(rp, h->sm_count, st));
        }
        // attention LSTMCell on [prenet; ctx(t-1)] (decoder.py:253-255)
        InferRowsParams la{};
        la.B = B; la.state = state; la.N = 4 * d.Ha; la.H = d.Ha; la.nseg = 2; la.epi = IR_EPI_LSTM;
        la.x[0] = w.xin_a; la.ldx[0] = KA; la.K[0] = KA; la.W[0] = Wia; la.ldw[0] = KA;
        la.x[1] = w.ha + (size_t)cur * B * d.Ha; la.ldx[1] = d.Ha; la.K[1] = d.Ha; la.W[1] = P("decoder.attention_rnn.weight_hh"); la.ldw[1] = d.Ha;
        la.bias1 = P("decoder.attention_rnn.bias_ih"); la.bias2 = P("decoder.attention_rnn.bias_hh");
        la.c = w.ca; la.h1 = w.ha + (size_t)nxt * B * d.Ha; la.ldh1 = d.Ha; la.h2 = w.xin_d; la.ldh2 = KD;
        if (use_tma) {
            InferLstmTmaLaunch ta{};
            ta.B = B; ta.H = d.Ha; ta.tf32 = c.gemm_tf32 >= 2; ta.K0 = KA; ta.K1 = d.Ha; ta.map_x0 = tmaps[0]; ta.map_w0 = tmaps[3]; ta.map_x1 = tmaps[1 + cur];
            ta.map_w1 = tmaps[4]; ta.bias_ih = la.bias1; ta.bias_hh = la.bias2; ta.c = la.c; ta.h1 = la.h1; ta.ldh1 = la.ldh1;
            ta.h2 = la.h2; ta.ldh2 = la.ldh2; ta.state = state;
            MSA_TRY(k_infer_lstm_tma(ta, h->sm_count, st));
        } else {
            MSA_TRY(k_infer_rows(la, h->sm_count, st));
        }
        // attention (forward_attn.py:178-219), one CTA per row; writes ctx(t) into the three packed inputs
        InferAttnParams ap{};
        ap.B = B; ap.L = L; ap.Ha = d.Ha; ap.A = d.A; ap.F = d.F; ap.Kl = d.Kl; ap.E = d.E; ap.norm = c.attn_norm; ap.max_steps = max_steps;
        ap.h = w.xin_d; ap.ldh = KD; ap.wq = P(at + "query_layer.linear_layer.weight");
        ap.wloc_t = w.attn_wloc_t; ap.wld4 = w.attn_wld4;
        ap.v = P(at + "v.linear_layer.weight"); ap.bv = P(at + "v.linear_layer.bias");
        ap.pm = w.pm; ap.memory = w.memory; ap.prev = w.prev; ap.cum = w.cum;
        ap.ctx1 = w.xin_a + d.Pd; ap.ld1 = KA; ap.ctx2 = w.xin_d + d.Ha; ap.ld2 = KD; ap.ctx3 = w.xin_p + d.Hd; ap.ld3 = KP;
        ap.align_out = align_out; ap.state = state;
        ap.windowing = c.windowing; ap.forward_attn = c.forward_attn; ap.forward_attn_mask = c.forward_attn && c.forward_attn_mask;
        ap.trans_agent = c.forward_attn && c.trans_agent;
        ap.win = win; ap.gmax = gmax; ap.alpha = w.fa_alpha; ap.u = w.fa_u;
        if (ap.trans_agent) { ap.wta = P(at + "ta.weight"); ap.bta = P(at + "ta.bias"); }
        if (c.windowing && first) {
            // first step: attention[:, 0] = attention.max() spans the batch (forward_attn.py:148-149): one launch collects the
            // maximum, the second one is the step
            ap.phase = 1;
            MSA_TRY(k_infer_attention(ap, st));
            ap.phase = 2;
        }
        MSA_TRY(k_infer_attention(ap, st));
        // decoder LSTMCell on [h_a; ctx] (decoder.py:260-264)
        InferRowsParams ld{};
        ld.B = B; ld.state = state; ld.N = 4 * d.Hd; ld.H = d.Hd; ld.nseg = 2; ld.epi = IR_EPI_LSTM;
        ld.x[0] = w.xin_d; ld.ldx[0] = KD; ld.K[0] = KD; ld.W[0] = Wid; ld.ldw[0] = KD;
        ld.x[1] = w.hd + (size_t)cur * B * d.Hd; ld.ldx[1] = d.Hd; ld.K[1] = d.Hd; ld.W[1] = P("decoder.decoder_rnn.weight_hh"); ld.ldw[1] = d.Hd;
        ld.bias1 = P("decoder.decoder_rnn.bias_ih"); ld.bias2 = P("decoder.decoder_rnn.bias_hh");
        ld.c = w.cd; ld.h1 = w.hd + (size_t)nxt * B * d.Hd; ld.ldh1 = d.Hd; ld.h2 = w.xin_p; ld.ldh2 = KP;
        if (use_tma) {
            InferLstmTmaLaunch ta{};
            ta.B = B; ta.H = d.Hd; ta.tf32 = c.gemm_tf32 >= 2; ta.K0 = KD; ta.K1 = d.Hd; ta.map_x0 = tmaps[5]; ta.map_w0 = tmaps[8]; ta.map_x1 = tmaps[6 + cur];
            ta.map_w1 = tmaps[9]; ta.bias_ih = ld.bias1; ta.bias_hh = ld.bias2; ta.c = ld.c; ta.h1 = ld.h1; ta.ldh1 = ld.ldh1;
            ta.h2 = ld.h2; ta.ldh2 = ld.ldh2; ta.state = state;
            MSA_TRY(k_infer_lstm_tma(ta, h->sm_count, st));
        } else {
            MSA_TRY(k_infer_rows(ld, h->sm_count, st));
        }
        // projections + stop logic (decoder.py:267-270, 381-395)
        static const bool pj_env = !(getenv("MSA_INFER_PROJ") && atoi(getenv("MSA_INFER_PROJ")) == 0);
        if (pj_env && infer_proj_supported(B, d.M, KP, KP)) {
            InferProjParams pj{};
            pj.B = B; pj.M = d.M; pj.K = KP; pj.x = w.xin_p; pj.ldx = KP;
            pj.wp = P("decoder.linear_projection.linear_layer.weight"); pj.bp = P("decoder.linear_projection.linear_layer.bias");
            pj.wg = P("decoder.gate_layer.linear_layer.weight"); pj.bg = P("decoder.gate_layer.linear_layer.bias");
            pj.mel_tm = w.mel_tm; pj.frame = w.frame; pj.not_finished = not_finished; pj.mel_lengths = mel_lengths_out;
            pj.state = state; pj.state_rw = state; pj.early = c.early_stopping; pj.max_steps = max_steps; pj.threshold = c.gate_threshold;
            MSA_TRY(k_infer_proj(pj, st));
            return 0;
        }
        InferRowsParams pp{};
        pp.B = B; pp.state = state; pp.N = d.M + 1; pp.nseg = 1; pp.epi = IR_EPI_BIAS;
        pp.x[0] = w.xin_p; pp.ldx[0] = KP; pp.K[0] = KP; pp.W[0] = P("decoder.linear_projection.linear_layer.weight"); pp.ldw[0] = KP;
        pp.Wb = P("decoder.gate_layer.linear_layer.weight"); pp.nsplit = d.M;
        pp.bias1 = P("decoder.linear_projection.linear_layer.bias"); pp.bias2 = P("decoder.gate_layer.linear_layer.bias");
        pp.out = w.mel_raw; pp.ldo = d.M; pp.out2 = w.gate_raw; pp.ldo2 = 1;
        pp.finish = 1; pp.M = d.M; pp.early = c.early_stopping; pp.max_steps = max_steps; pp.threshold = c.gate_threshold;
        pp.mel_tm = w.mel_tm; pp.frame = w.frame; pp.not_finished = not_finished; pp.mel_lengths = mel_lengths_out; pp.state_rw = state;
        pp.counter = counter;
        // optional K split of the projection over blockIdx.z (MSA_PROJ_KSPLIT=n, <= 16): measured on B200 the kernel's 20 us are
        // launch ramp + the serial stop-logic tail of the last CTA, not the K loop (SM-active 4.4 us with 7 slices), and the extra
        // partial-sum pass made the step 5 us slower end to end, so the default is 1 (profiles/r01_infer_notes.txt)
        static const int ks_env = getenv("MSA_PROJ_KSPLIT") ? atoi(getenv("MSA_PROJ_KSPLIT")) : 1;
        pp.ksplit = std::max(1, std::min(std::min(16, (KP + 127) / 128), ks_env));
        pp.part = w.proj_part;
        MSA_TRY(k_infer_rows(pp, h->sm_count, st));
        return 0;
    };
    // optional: capture a PAIR of steps (both ping-pong parities) as a CUDA graph and replay it; steps enqueued beyond the end
    // are no-ops (every kernel returns at once when the done flag is set)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    int enqueued = 0;
    if (getenv("MSA_INFER_GRAPH") != nullptr && max_steps > 2) {     // measured on B200: replaying the captured step is not faster
                                                                      // than plain stream-ordered launches (the queue stays ahead)
        MSA_TRY(step(0, true));
        MSA_TRY(step(1, false));
        enqueued = 2;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int rc = step(0, false);
            if (rc == 0) rc = step(1, false);
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc != 0 || e != cudaSuccess || !graph || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                graph = nullptr;
                gexec = nullptr;
                cudaGetLastError();
                MSA_CHECK(rc == 0, MSA_E_STATE, "msa_infer: a decoder step failed during graph capture");
            }
        } else {
            cudaGetLastError();
        }
    }
    int host_state[4] = {0, 0, 0, 0};
    const int check_every = 128;
    for (int s = enqueued; s < max_steps; s += gexec ? 2 : 1) {
        if (gexec) MSA_CUDA(cudaGraphLaunch(gexec, st));
        else MSA_TRY(step(s, s == 0));
        if (c.early_stopping && (s % check_every) >= check_every - 2 && (gexec || (s % check_every) == check_every - 1)) {      // all rows finished: stop enqueueing no-op steps
            MSA_CUDA(cudaMemcpyAsync(host_state, state, sizeof(int) * 3, cudaMemcpyDeviceToHost, st));
            MSA_CUDA(cudaStreamSynchronize(st));
            if (host_state[1]) break;
        }
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    MSA_CUDA(cudaMemcpyAsync(host_state, state, sizeof(int) * 3, cudaMemcpyDeviceToHost, st));
    MSA_CUDA(cudaStreamSynchronize(st));      // the postnet needs T' on the host (its "same" padding ends at T')
    const int Tn = host_state[2];
    MSA_CHECK(Tn >= 1 && Tn <= max_steps, MSA_E_STATE, "msa_infer: decoder produced %d steps", Tn);
    MSA_CUDA(cudaMemcpyAsync(n_steps_out, state + 2, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));

    // ---- postnet in eval mode on [B, T', M] + residual (decoder.py:63-72, tacotron2nv.py:155-157) ----
    const int64_t px = (int64_t)B * Tn * d.Cmax;
    MSA_TRY(k_transpose01(w.mel_tm, w.post_bt, Tn, B, d.M, st));      // [T'][B][M] -> mel [B][T'][M]
    const float* src = w.post_bt;
    for (int i = 0; i < d.nPost; ++i) {
        const int ci = i == 0 ? d.M : d.Cp, co = i == d.nPost - 1 ? d.M : d.Cp;
        float* dst = w.post_x + (i & 1) * px;
        MSA_TRY(conv_bn_eval(h, st, params, "postnet.convolutions." + std::to_string(i), src, w.post_y, dst, w.post_col,
                             w.enc_bn, w.enc_bn + std::max(d.C, d.Cmax), bn_stats + h->bn_offs[d.nEnc + i], B, Tn, ci, co, d.Kp,
                             i < d.nPost - 1 ? 2 : 0));
        src = dst;
    }
    MSA_TRY(k_add(w.post_bt, src, w.post_y, (int64_t)B * Tn * d.M, st));
    MSA_TRY(k_bt_to_ref_ld(w.post_y, mel_post_out, B, Tn, d.M, max_steps, st));
    return 0;
}

}  // extern "C"
