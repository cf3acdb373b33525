/*
 * msa_b200.h -- C ABI of the B200-native meta-training hot path of
 * HamedHemati/MetaSpeakerAdaptation-TTS (Tacotron2NV adapt-then-evaluate pass).
 *
 * The reference is pure Python/PyTorch and has NO FFI of its own;
 * every entry point below names the reference code it replaces (paths relative
 * to /root/reference/msa_tts/).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch types, no C++ exceptions.
 *   - every function returns int: 0 = ok, <0 = argument/shape/unsupported
 *     error (MSA_E_*), >0 = cudaError_t / (1000 + cublasStatus_t).
 *     msa_last_error_string() describes the last failure of the calling thread.
 *   - all device memory is owned by the caller (PyTorch caching allocator);
 *     the handle owns only its cuBLAS handle and small host-side tables.
 *   - all work is enqueued on the cudaStream_t passed in (as void*), nothing
 *     synchronises the host except where stated.
 *   - one host thread per handle; handles are independent.
 *   - there is no CPU fallback: without a CUDA device msa_create fails.
 *
 * Layouts (fp32 unless stated; "tm" = time-major)
 *   params / grads / m / v / fisher / means : one flat buffer, tensors in
 *       model.parameters() order, each starting on a 32-float boundary
 *       (msa_param_info gives name/offset/numel).
 *   bn_stats : per BatchNorm layer [running_mean(Cpad), running_var(Cpad)],
 *       Cpad = C rounded up to 32 (msa_bn_info).
 *   tokens int64 [B, L]; token_lengths int64 [B] sorted descending;
 *   mels [B, n_mel, T]; mel_lengths int64 [B]; speaker vectors [B, Ds];
 *   stop targets [B, T]  -- exactly the reference batch tuple
 *       (dataloaders/dataloader_meta.py:133-179, metatrainer.py:95-117).
 *   dropout keep-masks: one uint8 buffer (1 = keep), sections in the order of
 *       the reference's F.dropout calls, layouts given by msa_mask_info.
 */
#ifndef MSA_B200_H
#define MSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSA_OK 0
#define MSA_E_ARG (-1)          /* bad pointer / size / shape */
#define MSA_E_UNSUPPORTED (-2)  /* configuration outside the implemented path */
#define MSA_E_WORKSPACE (-3)    /* workspace too small */
#define MSA_E_NODEVICE (-4)     /* no sm_100 CUDA device */
#define MSA_E_STATE (-5)        /* call order (e.g. backward without forward) */

/* mirrors params["model"] of Tacotron2NV(params) (models/tacotron2nv.py:11-66) */
typedef struct msa_config {
    int32_t n_symbols;            /* n_symbols */
    int32_t enc_dim;              /* encoder_embedding_dim == symbols_embedding_dim */
    int32_t enc_kernel;           /* encoder_kernel_size */
    int32_t enc_n_convs;          /* encoder_n_convolutions */
    int32_t spk_mode;             /* 0 static, 1 static+linear, 2 learnable_lookup */
    int32_t spk_in_dim;           /* speaker_embedding_dim */
    int32_t spk_dim;              /* width appended to the encoder output */
    int32_t num_speakers;
    int32_t n_mel;                /* n_mel_channels (n_frames_per_step == 1) */
    int32_t prenet_dim;
    int32_t attn_rnn_dim;         /* hidden of decoder.attention_rnn (after the ctor swap) */
    int32_t dec_rnn_dim;          /* hidden of decoder.decoder_rnn */
    int32_t attn_dim;
    int32_t loc_filters;
    int32_t loc_kernel;
    int32_t post_dim;
    int32_t post_kernel;
    int32_t post_n_convs;
    int32_t attn_norm;            /* 0 softmax, 1 sigmoid (forward_attn.py:200-207) */
    int32_t forward_attn;         /* forward_attn.py:154-176 */
    int32_t trans_agent;
    int32_t windowing;            /* eval only, forward_attn.py:139-152 */
    int32_t forward_attn_mask;    /* eval only, forward_attn.py:163-173 */
    int32_t max_decoder_steps;
    int32_t early_stopping;       /* not decoder_no_early_stopping */
    int32_t loss_reduction;       /* 0 "none" (length-normalised masks), 1 "mean" */
    int32_t gemm_tf32;            /* batched-GEMM precision: 0 fp32 everywhere, 1 fp32 forward + TF32 backward, 2 TF32 everywhere */
    float p_attn_dropout;
    float p_dec_dropout;
    float gate_threshold;
    float loss_pos_weight;
    /* tacotron2nv.py:88-121 */
    int32_t freeze_charemb;       /* embedded characters detached: no gradient to embedding.weight */
    int32_t freeze_encoder;       /* encoder output detached: no gradient to encoder.* / embedding.weight */
    int32_t freeze_decoder;       /* decoder outputs detached: only the postnet receives gradients */
    int32_t residual_encoder;     /* use_residual_encoder: encoder output + embedded characters */
} msa_config;

typedef struct msa_handle msa_handle;

const char* msa_last_error_string(void);
int msa_version(void);

/* ---- handle ------------------------------------------------------------------------ */
int msa_create(const msa_config* cfg, int device, msa_handle** out);
int msa_destroy(msa_handle* h);
/* number of SMs / cooperative grid size the persistent kernels use */
int msa_sm_count(const msa_handle* h);

/* number of kernels of THIS library enqueued so far by the process (cuBLAS GEMMs are not counted) */
long long msa_launch_count(void);
/* per-kernel device timing of the persistent kernels with CUDA events on the launch stream:
 * enable (1), run passes, then read (synchronises) total milliseconds and launch counts per kernel id.
 * enable = 2 additionally switches the kernels to their instrumented variants (per-phase cycle counters and per-warp
 * traces, msa_profile_phases / msa_profile_trace): for the profiles/ scripts only, the kernels run slower. */
int msa_profile_enable(msa_handle* h, int enable);
int msa_profile_kernels(void);
const char* msa_profile_name(int id);
int msa_profile_read(msa_handle* h, double* total_ms, int64_t* counts);

/* ---- flat layout (replaces iterating model.parameters(), maml.py:71-76) -------------- */
int msa_param_count(const msa_handle* h);
int64_t msa_param_total(const msa_handle* h);           /* floats in a flat buffer (incl. padding) */
int msa_param_info(const msa_handle* h, int index, const char** name, int64_t* offset, int64_t* numel);
int msa_bn_count(const msa_handle* h);
int64_t msa_bn_total(const msa_handle* h);
int msa_bn_info(const msa_handle* h, int index, int64_t* offset, int32_t* channels);
/* dropout mask sections: index in call order; *rows x *cols uint8, layout string for docs */
int msa_mask_count(const msa_handle* h);
int64_t msa_mask_total(const msa_handle* h, int B, int T, int L);
int msa_mask_info(const msa_handle* h, int index, int B, int T, int L, const char** name, int64_t* offset,
                  int64_t* numel, float* p);
/* counter-based keep-mask generator (production path; parity tests inject masks instead) */
int msa_masks_generate(msa_handle* h, uint8_t* masks, int B, int T, int L, uint64_t seed, void* stream);

/* ---- one teacher-forced pass: Tacotron2NV.forward + Tacotron2Loss -------------------
 * replaces fmodel(**inputs) + criterion(...) (maml.py:50-53, reptile.py:52-55,
 * models/tacotron2nv.py:81-127, modules_tacotron2nv/tacotron2nv_loss.py:17-52).
 * Outputs are in the reference layouts: mel / mel_post [B, n_mel, T], gate [B, T],
 * align [B, T, L]; loss is one device float.  bn_stats (may be NULL) is updated in
 * place like BatchNorm1d in train mode.  Intermediates stay in `ws` for backward. */
size_t msa_workspace_bytes(const msa_handle* h, int B, int T, int L);
int msa_train_forward(msa_handle* h, void* ws, size_t ws_bytes, const float* params, float* bn_stats,
                      const int64_t* tokens, const int64_t* token_lengths, const float* mels,
                      const int64_t* mel_lengths, const float* speaker_vecs, const int64_t* speaker_ids,
                      const float* stop_targets, const uint8_t* masks, int B, int T, int L,
                      float* mel_out, float* mel_post_out, float* gate_out, float* align_out,
                      float* loss_out, void* stream);
/* replaces diffopt.step's autograd.grad / torch.autograd.grad(loss_test, ...) / loss.backward()
 * (maml.py:54,71-74; continual_ewc.py:355-356).  d_mel/d_mel_post/d_gate are upstream
 * gradients in the output layouts; pass NULL for all three to use d(loss)/d(outputs) of
 * the loss computed by the matching msa_train_forward.  grads (flat) is overwritten
 * (accumulate == 0) or accumulated into (accumulate != 0) with grad_scale * gradient. */
int msa_train_backward(msa_handle* h, void* ws, size_t ws_bytes, const float* params, const float* d_mel,
                       const float* d_mel_post, const float* d_gate, float* grads, int accumulate,
                       float grad_scale, void* stream);
/* Tacotron2Loss.__call__ (modules_tacotron2nv/tacotron2nv_loss.py:17-52) on the outputs of the last
 * msa_train_forward in `ws`, with the criterion's own reduction (0 "none", 1 "mean") and pos_weight
 * (metatrainer.py:83-86); also refreshes d(loss)/d(outputs) used by msa_train_backward(NULL,...). */
int msa_train_loss(msa_handle* h, void* ws, const float* stop_targets, const int64_t* mel_lengths, int reduction,
                   float pos_weight, float* loss_out, void* stream);
/* ---- task groups: the train-split passes that all tasks of a meta-batch take from the SAME weights theta_0
 * (maml.py:38-54, reptile.py:38-56: every `higher.innerloop_ctx` starts from `self.model`) as ONE pass over G tasks.
 * Everything that is not a recurrence runs task by task; the three persistent recurrences (encoder BiLSTM, attention chain,
 * decoder-RNN chain) run once for all G*B rows, so their per-step cross-CTA hand-offs are shared by the group.  Results per task
 * are those of G separate msa_train_forward / msa_train_backward calls (BatchNorm statistics, loss normalisation and parameter
 * gradients are per task).  All tasks must have the same B, T, L; G <= 8 and G*B <= 32 for the shared launches (otherwise, or
 * for configurations the grouped kernels do not implement, the recurrences silently run task by task -- same results).
 * ws: msa_group_workspace_bytes(h, G, B, T, L) bytes, 256-byte aligned; task g's slice (for msa_train_mcd / msa_get_buffer /
 * msa_train_loss) starts at ws + g * msa_group_workspace_bytes(h, 1, B, T, L).  Per-task arguments are HOST arrays of G device
 * pointers; loss_out: G floats on the device; grads[g]: flat gradient buffer of task g.
 * params[g]: the weights of task g.  All equal (the theta_0 train passes): the recurrences share one launch.  Different per task
 * (the test-split passes with the adapted weights, maml.py:56-76): every recurrence runs task by task, but the stages between
 * them still overlap across the tasks on the library's side streams (each task has its own workspace slice). */
size_t msa_group_workspace_bytes(const msa_handle* h, int G, int B, int T, int L);
int msa_train_forward_group(msa_handle* h, int G, void* ws, size_t ws_bytes, const float* const* params, float* const* bn_stats,
                            const int64_t* const* tokens, const int64_t* const* token_lengths, const float* const* mels,
                            const int64_t* const* mel_lengths, const float* const* speaker_vecs,
                            const int64_t* const* speaker_ids, const float* const* stop_targets,
                            const uint8_t* const* masks, int B, int T, int L, float* loss_out, void* stream);
int msa_train_backward_group(msa_handle* h, void* ws, size_t ws_bytes, const float* const* params, float* const* grads,
                             int accumulate, float grad_scale, void* stream);
/* Overlap of the meta-gradient allreduce with the tail of the last backward pass (parallel.py; the reference has no distributed code).
 * One-shot: the NEXT msa_train_backward / msa_train_backward_group records `cuda_event` (a cudaEvent_t, caller-owned) on its stream
 * as soon as every gradient outside the encoder is final -- before the encoder BiLSTM recurrence and the encoder convolutions run
 * (the backward pass produces the postnet, decoder and attention gradients first).  *prefix_floats (optional) receives the length of
 * the encoder part, which is the PREFIX of the flat layout (embedding.weight, encoder.*): the caller reduces grads[prefix:] when the
 * event fires and grads[:prefix] after the pass.  cuda_event == NULL only queries the prefix. */
int msa_backward_mark_event(msa_handle* h, void* cuda_event, int64_t* prefix_floats);
/* replaces utils/metrics.py:15-22 mcd_batch as called by the trainers' per-task logs (maml.py:78-82,
 * baseline.py, continual_*.py): K * mean_b mean_{t < len_b} ||mel_target - out||_2 with
 * K = 10 / ln(10) * sqrt(2), on the device, from the outputs of the last msa_train_forward in `ws`
 * (no device->host copy of the mel outputs, no host sync).  which = 0: the model's FIRST output
 * (pre-postnet mel -- the tensor the reference's logs call "out_post"), 1: mel_post.
 * mcd_out: 1 float on the device. */
int msa_train_mcd(msa_handle* h, void* ws, const int64_t* mel_lengths, int which, float* mcd_out, void* stream);
/* Tacotron2Loss.__call__ as a stand-alone operator on reference-layout tensors (tacotron2nv_loss.py:17-52):
 * mel / mel_post / mel_target [B, n_mel, T], gate / stop_targets [B, T], mel_lengths int64 [B]; writes the scalar loss and,
 * where the pointers are not NULL, d(loss)/d(mel, mel_post, gate).  scratch: msa_loss_scratch_floats(B, T, n_mel) floats. */
size_t msa_loss_scratch_floats(int B, int T, int n_mel);
int msa_tacotron2_loss(const float* mel, const float* mel_post, const float* gate, const float* mel_target,
                       const float* stop_targets, const int64_t* mel_lengths, int B, int T, int n_mel, int reduction,
                       float pos_weight, float* scratch, float* loss_out, float* d_mel, float* d_mel_post, float* d_gate,
                       void* stream);
/* d(loss)/d(mel, mel_post, gate) of the last forward, reference layouts (for autograd glue) */
int msa_loss_grads(msa_handle* h, void* ws, float* d_mel, float* d_mel_post, float* d_gate, void* stream);
/* The persistent kernels hand data between CTAs by polling (no grid barrier); a producer that never arrives makes
 * the consumers give up after a bounded spin and raise the handle's abort word instead of hanging the GPU.  The word
 * is STICKY: no pass clears it, so one check per meta-step covers every pass of that step (nothing comparable exists
 * in the reference -- its autograd graph cannot time out; call sites that consume the guarded results: maml.py:94-105,
 * reptile.py:82-89).
 *   msa_check_abort      synchronises `stream`, returns MSA_E_STATE if a kernel of this handle gave up (`ws` is unused)
 *   msa_abort_guard      device side, no sync: if the word is set, *sumsq <- NaN.  msa_flat_clip_sgd / msa_flat_clip_adam
 *                        skip their update when *sumsq is not finite, so garbage gradients never reach the weights or
 *                        the optimizer state
 *   msa_abort_read_async enqueues a 4-byte device->host copy of the word into (pinned) host memory; the trainers look
 *                        at it after the step's own synchronisation point and raise
 *   msa_abort_clear      resets the word (after the error has been reported)
 *   msa_debug_raise_abort test hook: raises the word exactly as a polling thread that timed out would */
int msa_check_abort(msa_handle* h, void* ws, void* stream);
int msa_abort_guard(msa_handle* h, float* sumsq, void* stream);
int msa_abort_read_async(msa_handle* h, uint32_t* host_flag, void* stream);
int msa_abort_clear(msa_handle* h, void* stream);
int msa_debug_raise_abort(msa_handle* h, void* stream);
/* cycles per phase recorded by thread 0 of every CTA of persistent kernel `id` (msa_profile_name) during the last
 * pass run with profiling enabled: out[ncta][8] (synchronises the device; profiles only) */
int msa_profile_phases(msa_handle* h, void* ws, int id, int64_t* out, int ncta);
/* development tracer: {SM clock, globaltimer ns} of lane 0 of every warp at every phase mark of the steps
 * [t0, t0+4) of persistent kernel `id`: out[ncta][16 warps][4 steps][12 tags][2] (synchronises the device) */
int msa_profile_trace_step(msa_handle* h, int t0);
int msa_profile_trace(msa_handle* h, void* ws, int id, int64_t* out, int ncta);
/* test hook: device pointer + element count of a named intermediate of the last pass */
int msa_get_buffer(msa_handle* h, void* ws, const char* name, void** ptr, int64_t* numel);

/* ---- free-running inference: Tacotron2NV.infer (models/tacotron2nv.py:130-162,
 * modules_tacotron2nv/decoder.py:334-411).  bn_stats are the (private) running statistics;
 * prenet_masks uint8 [max_steps, 2, B, prenet_dim] (dropout stays on).
 * Outputs: mel_post [B, n_mel, max_steps] (first *n_steps frames valid), mel_lengths
 * int32 [B], align [B, max_steps, L]; n_steps_out is one device int32. */
size_t msa_infer_workspace_bytes(const msa_handle* h, int B, int L, int max_steps);
int msa_infer(msa_handle* h, void* ws, size_t ws_bytes, const float* params, const float* bn_stats,
              const int64_t* tokens, const int64_t* token_lengths, const float* speaker_vecs,
              const int64_t* speaker_ids, const uint8_t* prenet_masks, int B, int L, int max_steps,
              float* mel_post_out, int32_t* mel_lengths_out, float* align_out, int32_t* n_steps_out,
              void* stream);

/* ---- fused multi-tensor kernels over flat buffers (n floats) -------------------------- */
/* higher's functional SGD step, p' = p - lr*(g + wd*p) with optional momentum buffer
 * (diffopt.step, maml.py:54 / reptile.py:56; torch.optim.SGD rule).  p_out may alias p. */
int msa_flat_sgd_step(const float* p, const float* g, float* p_out, float* momentum_buf, int64_t n, float lr,
                      float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                      void* stream);
/* mix_grad: acc = (init ? 0 : acc) + w * g   (utils/grad_utils.py:23-31, maml.py:94-98) */
int msa_flat_axpy(float* acc, const float* g, int64_t n, float w, int init, void* stream);
/* Reptile task "gradient" acc = (init ? 0 : acc) + w * -(p_T - p_0)   (reptile.py:73-77) */
int msa_flat_reptile_delta(float* acc, const float* p_T, const float* p_0, int64_t n, float w, int init,
                           void* stream);
/* apply_grad's norm (utils/grad_utils.py:8-20) / clip_grad_norm_'s total norm: out[0] = sum(g*g).
 * `partials` is caller scratch of msa_flat_partials() floats. */
int msa_flat_partials(void);
int msa_flat_sumsq(const float* g, int64_t n, float* partials, float* out, void* stream);
/* clip_grad_norm_ + outer_optimizer.step() (maml.py:101-105, reptile.py:85-89) fused:
 * coef = min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (max_norm <= 0: no clipping);
 * SGD:  p -= lr * coef * g  (+ momentum/weight decay as torch.optim.SGD)
 * Adam: torch.optim.Adam (no amsgrad), step = 1-based step index.
 * If *sumsq is not finite (NaN from msa_abort_guard, or a diverged gradient) the whole update is skipped: p and the
 * optimizer state stay as they are. */
int msa_flat_clip_sgd(float* p, const float* g, float* momentum_buf, const float* sumsq, int64_t n, float lr,
                      float max_norm, float momentum, float dampening, float weight_decay, int nesterov,
                      int first_step, void* stream);
int msa_flat_clip_adam(float* p, const float* g, float* m, float* v, const float* sumsq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int step, float max_norm,
                       void* stream);
/* Functional Adam step of the inner loop (higher's differentiable optimizer for a torch.optim.Adam inner optimizer -- the
 * reference builds the inner optimizer from YAML with any torch.optim class, utils/helpers.py:20-26; call sites maml.py:54,
 * reptile.py:56): torch.optim.Adam's rule (no amsgrad), p_out = p - lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps); p is not modified
 * (p_out may alias it); m, v are the per-context moment buffers, step is 1-based. */
int msa_flat_adam_step(const float* p, const float* g, float* p_out, float* m, float* v, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, void* stream);
/* EWC (continual_ewc.py): fisher += g*g / n_batches (59-82);
 * penalty = sum F (p-mu)^2 -> out[0] (84-89);
 * fused step p -= lr * (g + 2*lam*F*(p-mu)) and penalty in the same pass (345-357). */
int msa_ewc_fisher_accum(float* fisher, const float* g, int64_t n, float inv_n_batches, int init, void* stream);
int msa_ewc_penalty(const float* p, const float* mu, const float* fisher, int64_t n, float* partials,
                    float* out, void* stream);
int msa_ewc_sgd_step(float* p, const float* g, const float* mu, const float* fisher, int64_t n, float lr,
                     float lam, float* partials, float* penalty_out, void* stream);
/* The same penalty for any other optimizer (continual_ewc.py:213 builds self.optim with get_optimizer from ANY torch.optim class;
 * 345-357: loss += importance * penalty before loss.backward()): g += 2*lam*F*(p-mu) and the penalty at p in one pass; the
 * optimizer step (msa_flat_adam_step, msa_flat_sgd_step with momentum) follows on the completed gradient. */
int msa_ewc_penalty_grad(const float* p, float* g, const float* mu, const float* fisher, int64_t n, float lam,
                         float* partials, float* penalty_out, void* stream);

/* ---- tensor-core GEMM building block (exported for tests / profiles) ------------------------------------------
 * C[M,N] = alpha * A[M,K] . B[N,K]^T + beta * C, row-major, both operands K-contiguous (torch.nn.functional.linear's
 * contraction, the shape of every forward projection of the model).  tcgen05 TF32 MMAs fed by TMA; mode 0 = 3xTF32
 * split (fp32-accurate; the lo tiles are derived in shared memory), mode 1 = single TF32 product.  Output grids that
 * do not fill the GPU are split over K: `scratch` (msa_gemm_nt_scratch_floats(M, N, K) floats, 16-byte aligned)
 * holds the partial tiles, summed in a fixed order; scratch == NULL disables the K split.
 * Requires lda, ldb multiples of 4 floats and 16-byte aligned A, B. */
size_t msa_gemm_nt_scratch_floats(int64_t M, int64_t N, int64_t K);
int msa_gemm_nt(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
                float beta, float* C, int64_t ldc, int mode, float* scratch, void* stream);
/* The general row-major form C = alpha * op(A) . op(B) + beta * C on the same kernel: trans_a: A is given as [K][M] (else [M][K]),
 * trans_b: B is given as [N][K] (else [K][N]).  (0, 1) is msa_gemm_nt; (0, 0) is the input-gradient contraction dX = dY . W and
 * (1, 0) the weight-gradient contraction dW = dY^T . X of autograd's Linear / conv backward (maml.py:71-74 runs them through
 * torch.autograd.grad): operands whose contraction index is the ROW index are fed to tcgen05.mma as MN-major shared-memory tiles,
 * nothing is transposed in memory.  Leading dimensions multiples of 4 floats, 16-byte aligned operands. */
int msa_gemm(int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
             const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mode, float* scratch, void* stream);

/* ---- convolution building block (exported for tests / profiles) ---------------------------------------------------
 * The k-tap "same" convolutions of Encoder / Postnet (ConvNorm -> torch.nn.Conv1d, modules_tacotron2nv/encoder.py:36-37,
 * decoder.py:63-72, layers.py ConvNorm) and the two contractions of their autograd backward, as IMPLICIT GEMMs on the tcgen05
 * kernel: activations are channels-last [B][T][C]; the tap shift is a coordinate offset of a 3-D TMA tensor map whose
 * out-of-bounds zero fill is the convolution's padding, so no im2col matrix exists.  wp = tap-major copy [K][Cout][Cin] of the
 * native weight [Cout][Cin][K] (msa_conv1d_repack).  mode 0 = 3xTF32 (fp32-accurate), 1 = single TF32.
 *   msa_conv1d_fwd: y[B*T][Cout] = conv(x, w) + bias; stats (optional, msa_conv1d_stat_slabs x [3][Cout] floats): per 32-row slab
 *     of y the (count, mean, M2) of every channel -- BatchNorm's batch statistics without another pass over y;
 *   msa_conv1d_dx:  dx[B*T][Cin] = the transposed convolution of dy;
 *   msa_conv1d_dw:  dw[Cout][Cin][K] = (accumulate ? dw : 0) + scale * sum_{b,t} dy[b,t,co] x[b,t+k-pad,ci] (native layout).
 * Channels multiples of 4, K odd; scratch (msa_conv1d_scratch_floats floats, or NULL) holds K-split partial tiles. */
size_t msa_conv1d_scratch_floats(int B, int T, int Cin, int Cout, int K);
int msa_conv1d_stat_slabs(int B, int T, int Cin, int Cout, int K, int have_scratch);
int msa_conv1d_repack(const float* w, float* wp, int Cout, int Cin, int K, void* stream);
int msa_conv1d_fwd(const float* x, int B, int T, int Cin, const float* wp, int Cout, int K, const float* bias, float* y, int mode,
                   float* scratch, float* stats, void* stream);
int msa_conv1d_dx(const float* dy, int B, int T, int Cout, const float* wp, int Cin, int K, float* dx, int mode, float* scratch,
                  void* stream);
int msa_conv1d_dw(const float* dy, const float* x, int B, int T, int Cout, int Cin, int K, float scale, int accumulate, float* dw,
                  int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSA_B200_H */
