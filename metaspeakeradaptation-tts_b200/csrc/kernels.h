// Internal launcher declarations (host side) for the hot-path kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct msa_config;

namespace msa {

// Task groups: the theta_0 train-split passes of up to kGroupMax speaker tasks share one launch of each recurrence.  Every float
// array of task g lives at (pointer of task 0) + g * tstride floats (the tasks' workspaces are identical slices of one allocation);
// masks and lengths are caller-owned and come as per-task pointers.  Row r of a grouped launch = (task r / B, batch row r % B).
constexpr int kGroupMax = 8;
constexpr int kPtGroupMax = 4;      // tasks per launch of the per-task-weight variants (their weight stream must stay L2-resident)

// ---------------- persistent LSTM recurrence (lstm_rec.cu) ----------------
struct LstmRecParams {
    int T, B, H, ndir;
    const float* zin;          // [ndir][T][B][4H]  W_ih.x + b_ih + b_hh, gate-major rows (i,f,g,o)
    const float* whh;          // [4H][H] per direction
    int64_t whh_dir_stride;    // floats between the two directions' W_hh
    float* hout;               // [ndir][T][B][H]   (post-dropout) hidden = recurrent state
    float* cout;               // [ndir][T][B][H]
    float* gates;              // [ndir][T][B][4H]  post-activation i,f,g,o (stash for backward)
    const uint8_t* mask;       // [T][B][H] keep-mask or nullptr
    float drop_scale;          // 1/(1-p)
    const int64_t* lengths;    // [B] packed-sequence lengths or nullptr
    unsigned int* abort_word;  // zeroed once per workspace; set by a kernel whose polling timed out
    long long* prof;           // optional [grid][8] per-phase cycle counters (nullptr = off)
    long long* trace;          // optional per-warp time stamps (rec_common.cuh)
    int trace_t0;
    int flags;                 // hand-off variant (kFlagGate | kFlagWarp0, rec_common.cuh)
    // task group (chain_mma.cu only; the single-task kernels ignore these): G tasks of B rows each
    int G;                     // 0 / 1 = a single task
    int64_t tstride;           // floats between the copies of every float array of consecutive tasks
    const uint8_t* mask_g[kGroupMax];
    const int64_t* lengths_g[kGroupMax];
};
struct LstmRecBwdParams {
    int T, B, H, ndir;
    const float* whh;
    int64_t whh_dir_stride;
    const float* gates;
    const float* cout;
    const float* dh_ext;       // [ndir][T][B][H] grad w.r.t. hout from everything but the recurrence
    float* dz;                 // [ndir][T][B][4H] grad w.r.t. gate pre-activations
    const uint8_t* mask;
    float drop_scale;
    const int64_t* lengths;
    unsigned int* abort_word;
    long long* prof;
    long long* trace;
    int trace_t0;
    int flags;
    // task group (chain_mma.cu only; the single-task kernels ignore these): G tasks of B rows each
    int G;                     // 0 / 1 = a single task
    int64_t tstride;           // floats between the copies of every float array of consecutive tasks
    const uint8_t* mask_g[kGroupMax];
    const int64_t* lengths_g[kGroupMax];
};
bool lstm_rec_single_ok(int B, int H, int ndir, int sm_count, size_t smem_limit);
size_t lstm_rec_fwd_smem(int B, int H);
size_t lstm_rec_bwd_smem(int B, int H);
int launch_lstm_rec_fwd(const LstmRecParams& p, int sm_count, size_t smem_limit, cudaStream_t st);
int launch_lstm_rec_bwd(const LstmRecBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st);

// ---------------- attention-RNN + location-sensitive attention chain (attn_chain.cu) ----------------
struct AttnChainParams {
    int T, B, L, Ha, A, F, Kl, norm;   // norm: 0 softmax, 1 sigmoid
    const float* xw;       // [T][B][4Ha]  W_ih[:, :prenet].x_t + b_ih + b_hh
    const float* whh;      // [4Ha][Ha]
    const float* mw_rm;    // [4Ha][B*L]   (W_ih[:, prenet:] . memory^T), row-major over gate rows
    const float* wq;       // [A][Ha]
    const float* pm;       // [B][L][A]    processed memory
    const float* wloc;     // [F][2][Kl]
    const float* wld;      // [A][F]
    const float* v;        // [A]
    const float* bv;       // [1]
    const uint8_t* mask;   // [T][B][Ha]
    float drop_scale;
    float* ha;             // [T][B][Ha]
    float* ca;             // [T][B][Ha]
    float* ga;             // [T][B][4Ha]
    float* q;              // [T][B][A]
    float* align;          // [T][B][L]
    float* cum;            // [T][B][L]   cumulative weights fed to the location conv at step t
    float* s;              // [T][B][L][A] tanh(q + loc + pm)
    float* convf;          // [T][B][L][F]
    float* znorm;          // [T][B]      normaliser (sum of sigmoids) for norm == 1
    float* e;              // [T][B][L]   energies (pre-normalisation)
    // forward attention (forward_attn.py:154-176,222-224); align then holds the forward-attended alignment alpha(t)
    int fa, ta;            // forward_attn, trans_agent
    const float* mta;      // [B][L]      memory . W_ta[:E]  (context half of the transition agent, ctx = alpha . memory)
    const float* wta_h;    // [Ha]        W_ta[E:]           (query half)
    const float* bta;      // [1]
    float* aplain;         // [T][B][L]   normalise(e(t)) before the forward-attention recursion
    float* fsum;           // [T][B]      sum_l alpha'(t)[l] (renormaliser)
    float* ustash;         // [T][B]      transition probability used at step t (= u(t-1); 0.5 without the agent)
    unsigned int* abort_word;
    long long* prof;       // optional [grid][8] per-phase cycle counters
    long long* trace;
    int trace_t0;
    int flags;
    // task group (chain_mma.cu only; the single-task kernels ignore these): G tasks of B rows each
    int G;                     // 0 / 1 = a single task
    int64_t tstride;           // floats between the copies of every float array of consecutive tasks
    const uint8_t* mask_g[kGroupMax];
    const int64_t* lengths_g[kGroupMax];
    // per-task weights (the test-split passes of a meta-batch: every task has its own adapted weights, maml.py:56-76): pt = 1,
    // G <= kPtGroupMax tasks; the recurrent weight slices are streamed every step from ready-made fragment arrays (L2-resident,
    // written by launch_attn_frag_fwd) instead of being resident in shared memory; the small attention weights per task
    int pt;
    const uint4* wfrag;        // [G][ncta][16 warps][KS][m tile][hi|lo][32] A fragments of (W_hh rows, W_q row)
    int64_t wfrag_stride;      // uint4 words between consecutive tasks
    const float* whh_g[kPtGroupMax];
    const float* wq_g[kPtGroupMax];
    const float* wloc_g[kPtGroupMax];
    const float* wld_g[kPtGroupMax];
    const float* v_g[kPtGroupMax];
    const float* bv_g[kPtGroupMax];
};
struct AttnChainBwdParams {
    int T, B, L, Ha, A, F, Kl, norm;
    const float* whh;
    const float* mw_pm;    // [B*L][4Ha]
    const float* wq;
    const float* wloc;
    const float* wld;
    const float* v;
    const uint8_t* mask;
    float drop_scale;
    const float* ga;
    const float* ca;
    const float* align;
    const float* s;
    const float* znorm;
    const float* dha_ext;  // [T][B][Ha]
    const float* da_ext;   // [T][B][L]
    float* dza;            // [T][B][4Ha]
    float* dq;             // [T][B][A]
    float* de;             // [T][B][L]
    float* ds;             // [T][B][L][A]
    float* dconvf;         // [T][B][L][F]
    float* dat;            // [T][B][L]   total d a(t);  forward attention: [T][2][B][L] = {d alpha(t), d a(t) through cum}
    int fa, ta;
    const float* mta; const float* wta_h;
    const float* aplain; const float* fsum; const float* ustash;
    float* dzu;            // [T][B]      gradient of the transition agent's pre-activation
    unsigned int* abort_word;
    long long* prof;
    long long* trace;
    int trace_t0;
    int flags;
    // task group (chain_mma.cu only; the single-task kernels ignore these): G tasks of B rows each
    int G;                     // 0 / 1 = a single task
    int64_t tstride;           // floats between the copies of every float array of consecutive tasks
    const uint8_t* mask_g[kGroupMax];
    const int64_t* lengths_g[kGroupMax];
    // per-task weights (see AttnChainParams): B fragments of W_hh^T streamed from wfrag, W_q^T and the small weights per task
    int pt;
    const uint4* wfrag;        // [G][ncta][16 warps][KS][32] {hi.x, hi.y, lo.x, lo.y}
    int64_t wfrag_stride;
    const float* whh_g[kPtGroupMax];
    const float* wq_g[kPtGroupMax];
    const float* wloc_g[kPtGroupMax];
    const float* wld_g[kPtGroupMax];
    const float* v_g[kPtGroupMax];
};
bool attn_chain_single_ok(int B, int L, int Ha, int A, int F, int Kl, int sm_count, size_t smem_limit, bool fa);
size_t attn_chain_fwd_smem(int B, int L, int Ha, int A, int F, int Kl, int sm_count, bool mw_resident, bool fa = false);
size_t attn_chain_bwd_smem(int B, int L, int Ha, int A, int F, int Kl, int sm_count, bool mwp_resident, bool fa = false);
int launch_attn_chain_fwd(const AttnChainParams& p, int sm_count, size_t smem_limit, cudaStream_t st);
int launch_attn_chain_bwd(const AttnChainBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st);

// ---------------- grouped recurrences with tensor-core gate products (chain_mma.cu) ----------------
// Same arithmetic as the kernels above for G tasks at once (R = G*B <= 32 rows per hand-off); the per-step gate products run on the
// tensor cores as bf16x3 split products (hi.hi + hi.lo + lo.hi, fp32 accumulation: ~1e-5 relative, inside the TF32-path tolerance).
bool chain_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit);        // LSTM recurrences
bool attn_chain_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit);   // attention chain, forward
bool attn_chain_bwd_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit);   // ... backward
int launch_lstm_rec_fwd_mma(const LstmRecParams& p, int sm_count, size_t smem_limit, cudaStream_t st);
int launch_lstm_rec_bwd_mma(const LstmRecBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st);
int launch_attn_chain_fwd_mma(const AttnChainParams& p, int sm_count, size_t smem_limit, cudaStream_t st);
int launch_attn_chain_bwd_mma(const AttnChainBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st);
// Per-task weights (params.pt = 1): G <= kPtGroupMax tasks of B <= 8 rows per launch, every task with its own weights (whh_g, wq_g,
// ...); the launchers first write the tasks' recurrent weight slices as ready-made fragments into params.wfrag (at least
// G * attn_chain_pt_frag_bytes; the kernels stream them from L2 every step) and then launch the chain.
bool attn_chain_pt_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit);
size_t attn_chain_pt_frag_bytes(const msa_config& cfg, int sm_count);      // per task; forward and backward fragments have this size

// ---------------- element-wise / layout / reduction kernels (model_kernels.cu) ----------------
int k_embedding_fwd(const float* w, const int64_t* tok, float* x, int rows, int C, int n_symbols, cudaStream_t st);
int k_embedding_add(const float* w, const int64_t* tok, float* x, int rows, int C, int ld, int n_symbols, cudaStream_t st);   // x[:, :C] += w[tok]
int k_add_cols(float* dst, const float* src, int64_t rows, int C, int ld, cudaStream_t st);                                  // dst += src[:, :C]
int k_embedding_bwd(const float* dx, const int64_t* tok, float* gw, int rows, int C, int n_symbols, float scale, int accumulate, cudaStream_t st);
int k_im2col(const float* x, float* col, int B, int T, int C, int K, cudaStream_t st);
int k_col2im(const float* dcol, float* dx, int B, int T, int C, int K, cudaStream_t st);
int k_fill_rows(float* y, const float* b1, const float* b2, int64_t rows, int N, cudaStream_t st);
// Reduction scratch (`red_scr`, may be null = single-chunk reductions): kRedTickets zero-initialised uint tickets followed by
// kRedChunks x kRedCols floats of partials; shared by the column reductions of one stream (they run back to back).
constexpr int kRedTickets = 256, kRedChunks = 16, kRedCols = 8192;
constexpr int64_t kRedScrFloats = kRedTickets + (int64_t)kRedChunks * kRedCols;
// x_round (optional, x itself or another [rows][ld] array): x written back rounded to the nearest TF32 value (the sum uses the unrounded values)
int k_colsum(const float* x, int64_t rows, int N, int ld, float* out, float scale, int accumulate, float* out2, float* red_scr, cudaStream_t st,
             float* x_round = nullptr);
int k_bn_stats(const float* y, int64_t rows, int C, float* mean, float* invstd, float* running, int Cpad, float* red_scr, cudaStream_t st);
// BatchNorm (train) from the slab statistics of the conv epilogue ([nslab][3][C]: count, mean, M2) + activation + dropout in ONE
// kernel: every block merges the slabs of its 32 channels (Chan's update, slab order), block row 0 also stores mean / invstd / running
int k_bn_slab_act_drop_fwd(const float* y, const float* slabs, int nslab, int64_t rows, int C, float* mean, float* invstd, float* running,
                           int Cpad, const float* gamma, const float* beta, const uint8_t* mask, float ds, int act, float* out, cudaStream_t st);
int k_bn_eval_stats(const float* running, int C, int Cpad, float* mean, float* invstd, cudaStream_t st);
// act: 0 none, 1 relu, 2 tanh
int k_bn_act_drop_fwd(const float* y, const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const uint8_t* mask, float drop_scale, int act, float* out, int64_t rows, int C, cudaStream_t st);
int k_bn_act_drop_bwd(const float* dout, const float* y, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, const uint8_t* mask, float drop_scale, int act, float* dy, float* ggamma,
                      float* gbeta, float* scratch, int64_t rows, int C, float scale, int accumulate, float* red_scr, cudaStream_t st);
int k_relu_drop_fwd(float* x, const uint8_t* mask, float drop_scale, int64_t n, cudaStream_t st);
int k_relu_drop_bwd(float* dx, const float* out, const uint8_t* mask, float drop_scale, int64_t n, cudaStream_t st);
int k_transpose01(const float* in, float* out, int D0, int D1, int C, cudaStream_t st);   // [D0][D1][C] -> [D1][D0][C]
int k_prep_mels(const float* mels, float* frames_tm, float* target_bt, int B, int M, int T, cudaStream_t st);
int k_build_memory(const float* enc_h, const float* spk, float* memory, int B, int L, int Hh, int Ds, cudaStream_t st);
int k_split_dmemory(const float* dmem, float* denc_h, float* dspk, int B, int L, int Hh, int Ds, cudaStream_t st);
int k_bt_to_ref(const float* x_bt, float* out, int B, int T, int M, cudaStream_t st);      // [B][T][M] -> [B][M][T]
int k_ref_to_bt(const float* x, float* out_bt, int B, int T, int M, cudaStream_t st);      // [B][M][T] -> [B][T][M]
int k_add(const float* a, const float* b, float* out, int64_t n, cudaStream_t st);
int k_add3(const float* a, const float* b, const float* c, float* out, int64_t n, cudaStream_t st);
int k_mcd(const float* out_bt, const float* target_bt, const int64_t* mel_len, int B, int T, int M, float* partials, float* result,
          cudaStream_t st);
int k_loss(const float* pre_bt, const float* post_bt, const float* gate_bt, const float* target_bt, const float* stop,
           const int64_t* mel_len, int B, int T, int M, int reduction, float pos_weight, float* partials, float* loss,
           float* dpre, float* dpost, float* dgate, cudaStream_t st);
int k_sum_over_t(const float* x, float* out, int T, int64_t n, cudaStream_t st);           // out[n] = sum_t x[t][n]
int wloc_grad_partials(int T, int B);   // number of partial tables of F*2*Kl floats k_wloc_grad needs as scratch
int k_wloc_grad(const float* dconvf, const float* align, const float* cum, float* gw, float* partials, int T, int B, int L, int F,
                int Kl, float scale, int accumulate, cudaStream_t st);
int k_dot_rows(const float* a, const float* b, int64_t n, float* partials, float* out, float scale, int accumulate, cudaStream_t st);
int k_masks_generate(uint8_t* masks, const int64_t* offsets, const int64_t* numels, const float* ps, int nsec, uint64_t seed, cudaStream_t st);
// ---------------- free-running inference, fused per-step kernels (infer_decode.cu, infer_kernels.cu) ----------------
enum { IR_EPI_BIAS = 0, IR_EPI_RELU_DROP = 1, IR_EPI_LSTM = 2 };
// out[b][n] = sum over the K segments of W_s[n][:] . x_s[b][:]  (+ epilogue), B rows in tiles of 32
struct InferRowsParams {
    int B, N, nseg, rotate;
    const float* x[2]; int ldx[2]; int K[2];      // inputs [B][ldx], K columns each (multiples of 4, 16-byte aligned rows)
    const float* W[2]; int ldw[2];                // weights [N][ldw]
    const float* Wb; int nsplit;                  // rows >= nsplit (if > 0) come from a second matrix Wb (segment 0 only)
    int epi;
    const float* bias1; const float* bias2;       // BIAS: bias of rows < nsplit / >= nsplit; LSTM: b_ih, b_hh
    float* out; int ldo; float* out2; int ldo2;   // BIAS / RELU_DROP outputs (out2: rows >= nsplit)
    const uint8_t* mask; int mask_layer;          // RELU_DROP: prenet masks [steps][2][B][N]
    float* c; float* h1; int ldh1; float* h2; int ldh2; int H;     // LSTM: cell state (in place), two copies of the new h
    const int* state;                             // [0] step, [1] done, [2] steps produced
    // end-of-step logic run by the last CTA (finish != 0): out = mel frame [B][M], out2 = gate [B]
    int finish, M, early, max_steps;
    float threshold;
    float* mel_tm; float* frame; int* not_finished; int* mel_lengths; int* state_rw; unsigned int* counter;
    int ksplit; float* part;                      // finish only: K split over blockIdx.z, partials [ksplit][B][N]
};
int k_infer_rows(const InferRowsParams& p, int sm_count, cudaStream_t st);
struct InferAttnParams {
    int B, L, Ha, A, F, Kl, E, norm, max_steps;
    const float* h; int ldh;          // [B][ldh] h_a'(t)
    const float* wq;                  // [A][Ha]
    const float* wloc_t;              // [2*Kl][F]            location conv weights, transposed (k_infer_attn_prep)
    const float* wld4;                // [ceil(F/4)][A][4]    location dense weights, blocked (k_infer_attn_prep)
    const float* v; const float* bv;
    const float* pm;
    const float* memory;
    float* prev; float* cum;
    float* ctx1; int ld1; float* ctx2; int ld2; float* ctx3; int ld3;
    float* align_out;
    const int* state;
    // eval-time variants (forward_attn.py:139-176,222-224)
    int windowing, forward_attn, forward_attn_mask, trans_agent;
    int phase;                        // 0 whole step; windowing, first step: 1 = contribute the batch-wide energy maximum, 2 = the step
    int* win;                         // [2] window index by step parity
    float* gmax;                      // batch-wide maximum of the first step's windowed energies
    float* alpha; float* u;           // [B][L], [B] forward-attention state
    const float* wta; const float* bta;
    int mem_res;                      // set by k_infer_attention: the memory tile is prefetched into shared memory
};
// LSTMCell of the inference step with TMA tensor-map boxes feeding the tensor cores (infer_lstm_tma.cu); maps are 128-byte
// CUtensorMap objects (opaque here), built once per msa_infer call
struct InferLstmTmaLaunch {
    int B, H, K0, K1;
    int tf32;                                      // 1: plain TF32 products (GEMM policy 2), 0: 3xTF32 (fp32-accurate)
    const void* map_x0; const void* map_w0;       // [input | ...] segment: x0 [B][K0], W_ih [4H][K0]
    const void* map_x1; const void* map_w1;       // recurrent segment: h [B][H], W_hh [4H][H]
    const float* bias_ih; const float* bias_hh;
    float* c; float* h1; int ldh1; float* h2; int ldh2;
    const int* state;
};
bool infer_lstm_tma_supported(int H, int K0, int ld0, int K1, int ld1, int sm_count);
int infer_lstm_tma_map_x(void* map_out, const float* x, int B, int K, int ld);
int infer_lstm_tma_map_w(void* map_out, const float* W, int H, int K, int ld);
int k_infer_lstm_tma(const InferLstmTmaLaunch& a, int sm_count, cudaStream_t st);
// both prenet layers in one cluster launch (infer_decode.cu)
struct InferPrenetParams {
    int B, M, Pd;
    const float* frame;               // [B][M] previous mel frame
    const float* w1; const float* w2; // [Pd][M], [Pd][Pd]
    const uint8_t* mask;              // [steps][2][B][Pd]
    float* out; int ldo;              // packed attention-LSTM input [B][ldo], first Pd columns
    const int* state;
};
bool infer_prenet_supported(int M, int Pd);
int k_infer_prenet(const InferPrenetParams& p, cudaStream_t st);
// mel / gate projection + stop logic in one cluster launch (infer_decode.cu), B <= 32
struct InferProjParams {
    int B, M, K;                      // batch rows, n_mel, columns of [h_d; ctx]
    const float* x; int ldx;          // [B][ldx]
    const float* wp; const float* bp; // linear_projection [M][K], [M]
    const float* wg; const float* bg; // gate_layer [1][K], [1]
    float* mel_tm; float* frame; int* not_finished; int* mel_lengths;
    const int* state; int* state_rw;
    int early, max_steps; float threshold;
};
bool infer_proj_supported(int B, int M, int K, int ldx);
int k_infer_proj(const InferProjParams& p, cudaStream_t st);
int k_init_fwd_attn(float* alpha, float* u, int* win, float* gmax, int B, int L, cudaStream_t st);
int k_fill_ones_i32(int* p, int n, cudaStream_t st);
size_t infer_attention_smem(int L, int Ha, int A, int F, int Kl, int E, bool mem_res);
int k_infer_attention(const InferAttnParams& p, cudaStream_t st);
int k_infer_attn_prep(const float* wloc, const float* wld, float* wloc_t, float* wld4, int F, int Kl, int A, cudaStream_t st);
int k_tm_to_ref_ld(const float* x_tm, float* out, int T, int B, int M, int ld, cudaStream_t st);
int k_bt_to_ref_ld(const float* x_bt, float* out, int B, int T, int M, int ld, cudaStream_t st);

// ---------------- tcgen05 / TMA tensor-core GEMM (gemm_tc.cu): C = alpha A.B^T + beta C, both operands K-contiguous ----------------
bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc);
size_t gemm_tc_scratch_floats(int64_t M, int64_t N, int64_t K);
int gemm_tc_nt(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
               float* C, int64_t ldc, int mode, float* scratch, cudaStream_t st, const float* bias1 = nullptr,
               const float* bias2 = nullptr);   // mode 0: 3xTF32 (fp32-accurate), 1: TF32
// general row-major form: ta: A is given as [K][M] (else [M][K]); tb: B is given as [N][K] (else [K][N])
int gemm_tc(bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
            float beta, float* C, int64_t ldc, int mode, float* scratch, cudaStream_t st, const float* bias1 = nullptr,
            const float* bias2 = nullptr);
// Implicit convolutions on the same kernel (channels-last activations [B][Tn][C], wp = tap-major weight copy [K][Co][Ci]); mode as above
bool conv_tc_supported(int B, int Tn, int Ci, int Co, int K, const float* x, const float* w);
size_t conv_tc_scratch_floats(int B, int Tn, int C, int N, int K);
int conv_tc_stat_slabs(int B, int Tn, int Ci, int Co, int K, bool have_scratch);
int conv_tc_fwd(const float* x, int B, int Tn, int Ci, const float* wp, int Co, int K, const float* bias, float* y, int mode, float* scratch,
                float* stats, cudaStream_t st);
int conv_tc_dx(const float* dy, int B, int Tn, int Co, const float* wp, int Ci, int K, float* dx, int mode, float* scratch, cudaStream_t st);
int conv_tc_dw(const float* dy, const float* x, int B, int Tn, int Co, int Ci, int K, float scale, int accumulate, float* dw, int mode,
               cudaStream_t st);
int conv_tc_repack(int n, const float* const* src, float* const* dst, const int* co, const int* ci, const int* k, cudaStream_t st);

int k_abort_guard(unsigned int* abort_word, float* sumsq, int raise, cudaStream_t st);   // raise: set the word; else NaN -> *sumsq if set
int k_fill_canary(float* p, int64_t n, cudaStream_t st);   // n floats (multiple of 4, 16-byte aligned) <- 0xFFFFFFFF
int k_scale_copy(const float* in, float* out, int64_t n, float scale, int accumulate, cudaStream_t st);

}  // namespace msa
