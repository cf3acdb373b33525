// Per-step kernels of free-running inference (Decoder.infer, modules_tacotron2nv/decoder.py:334-411).
//
// One decoder step is a fixed sequence of launches that reads the step index from DEVICE memory, so the sequence is
// captured once as a CUDA graph and replayed max_decoder_steps times with no host round trip per step (the reference
// synchronises the host every step for the stop test, decoder.py:385-395).  Every kernel returns immediately once the
// device-side `done` flag is set (early stopping, decoder.py:392-395).
//   state[0] = step index t, state[1] = done, state[2] = number of steps produced
#include "common.cuh"
#include "kernels.h"

namespace msa {

// out[b][j] = relu(out[b][j]) * keep / (1-p) on a [B][N] block with row stride ld; mask = prenet_masks[t][layer][B][N]
__global__ void ker_infer_relu_drop(float* x, int ld, const uint8_t* __restrict__ masks, int layer, int B, int N, const int* state) {
    if (state[1]) return;
    const int t = state[0];
    const uint8_t* mk = masks + ((size_t)t * 2 + layer) * B * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * N; i += gridDim.x * blockDim.x) {
        const int b = i / N, j = i - b * N;
        const float v = fmaxf(x[(size_t)b * ld + j], 0.f);
        x[(size_t)b * ld + j] = mk[i] ? v * 2.f : 0.f;
    }
}

// LSTMCell point-wise (eval: no dropout): z [B][4H] = W_ih.x + W_hh.h (no biases yet), gate order i,f,g,o
// h is written to two places (its own recurrent buffer and the packed input of the next GEMM)
__global__ void ker_infer_lstm_point(const float* __restrict__ z, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                     float* c, float* h1, int ld1, float* h2, int ld2, int B, int H, const int* state) {
    if (state[1]) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += gridDim.x * blockDim.x) {
        const int b = i / H, u = i - b * H;
        const float* zb = z + (size_t)b * 4 * H;
        const float gi = sigmoidf_(zb[u] + b_ih[u] + b_hh[u]);
        const float gf = sigmoidf_(zb[H + u] + b_ih[H + u] + b_hh[H + u]);
        const float gg = tanhf(zb[2 * H + u] + b_ih[2 * H + u] + b_hh[2 * H + u]);
        const float go = sigmoidf_(zb[3 * H + u] + b_ih[3 * H + u] + b_hh[3 * H + u]);
        const float cn = gf * c[i] + gi * gg;
        c[i] = cn;
        const float hv = go * tanhf(cn);
        h1[(size_t)b * ld1 + u] = hv;
        if (h2) h2[(size_t)b * ld2 + u] = hv;
    }
}

// Location-sensitive attention for one batch row per CTA (forward_attn.py:121-131,178-219, eval, no windowing / forward
// attention): loc = dense(conv([prev; cum])); e = v.tanh(q + loc + pm) + bv; a = softmax | sigmoid-norm; cum += a; prev = a;
// ctx = a.memory.  q = Wq.h_a comes from a GEMM before this kernel; the location weights are staged in shared memory
// (coalesced) so that every inner loop reads shared memory only.  ctx is written to the three packed GEMM inputs.
constexpr int kInferAttnThreads = 512;
__global__ void __launch_bounds__(kInferAttnThreads) ker_infer_attention(InferAttnParams p) {
    if (p.state[1]) return;
    extern __shared__ float sm[];
    const int t = p.state[0], b = blockIdx.x;
    const int L = p.L, A = p.A, F = p.F, Kl = p.Kl, pl = (Kl - 1) / 2, LH = L + Kl - 1, FP = F + 1;
    float* in_s = sm;                         // [2][LH] prev / cum with zero halo
    float* q_s = in_s + 2 * LH;               // [A]
    float* v_s = q_s + A;                     // [A]
    float* wloc_s = v_s + A;                  // [2*Kl][F]   wloc[f][c][k] -> [c*Kl+k][f]
    float* wldT_s = wloc_s + 2 * Kl * F;      // [F][A]      wld[d][f] -> [f][d]
    float* cf_s = wldT_s + (size_t)F * A;     // [L][F+1]
    float* e_s = cf_s + (size_t)L * FP;       // [L]
    float* red = e_s + L;                     // [40]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * LH; i += blockDim.x) {
        const int c = i / LH, l = i % LH - pl;
        in_s[i] = (l >= 0 && l < L) ? (c == 0 ? p.prev[(size_t)b * L + l] : p.cum[(size_t)b * L + l]) : 0.f;
    }
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        q_s[i] = p.q[(size_t)b * A + i];
        v_s[i] = __ldg(p.v + i);
    }
    for (int i = threadIdx.x; i < F * 2 * Kl; i += blockDim.x) {
        const int f = i / (2 * Kl), ck = i % (2 * Kl);
        wloc_s[ck * F + f] = __ldg(p.wloc + i);
    }
    for (int i = threadIdx.x; i < A * F; i += blockDim.x) {
        const int d = i / F, f = i % F;
        wldT_s[f * A + d] = __ldg(p.wld + i);
    }
    __syncthreads();
    // location conv: (l, f) items, lanes over f
    for (int i = threadIdx.x; i < L * F; i += blockDim.x) {
        const int l = i / F, f = i - l * F;
        float a0 = 0.f, a1 = 0.f;
        for (int k = 0; k < Kl; ++k) {
            a0 += wloc_s[k * F + f] * in_s[l + k];
            a1 += wloc_s[(Kl + k) * F + f] * in_s[LH + l + k];
        }
        cf_s[l * FP + f] = a0 + a1;
    }
    __syncthreads();
    // energies: one warp per position, lanes over attention dims
    const float bv = __ldg(p.bv);
    for (int l = w; l < L; l += nw) {
        float e = 0.f;
        for (int d = lane; d < A; d += 32) {
            float l0 = 0.f, l1 = 0.f;
            int f = 0;
            for (; f + 1 < F; f += 2) {
                l0 += wldT_s[f * A + d] * cf_s[l * FP + f];
                l1 += wldT_s[(f + 1) * A + d] * cf_s[l * FP + f + 1];
            }
            if (f < F) l0 += wldT_s[f * A + d] * cf_s[l * FP + f];
            e += v_s[d] * tanhf(q_s[d] + l0 + l1 + __ldg(p.pm + ((size_t)b * L + l) * A + d));
        }
        e = warp_sum(e);
        if (lane == 0) e_s[l] = e + bv;
    }
    __syncthreads();
    // normalise (block reduction over L)
    float m = -INFINITY;
    if (p.norm == 0) {
        for (int l = threadIdx.x; l < L; l += blockDim.x) m = fmaxf(m, e_s[l]);
        m = warp_max(m);
        if (lane == 0) red[w] = m;
        __syncthreads();
        m = red[0];
        for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
        __syncthreads();
    }
    float s = 0.f;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float x = p.norm == 0 ? expf(e_s[l] - m) : sigmoidf_(e_s[l]);
        e_s[l] = x;
        s += x;
    }
    s = block_sum(s, red);
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float a = e_s[l] / s;
        e_s[l] = a;
        p.prev[(size_t)b * L + l] = a;
        p.cum[(size_t)b * L + l] += a;
        p.align_out[((size_t)b * p.max_steps + t) * L + l] = a;
    }
    __syncthreads();
    // context: threads over memory channels (coalesced rows), 4 positions in flight
    for (int e = threadIdx.x; e < p.E; e += blockDim.x) {
        const float* mrow = p.memory + (size_t)b * L * p.E + e;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
        int l = 0;
        for (; l + 3 < L; l += 4) {
            c0 += e_s[l] * __ldg(mrow + (size_t)l * p.E);
            c1 += e_s[l + 1] * __ldg(mrow + (size_t)(l + 1) * p.E);
            c2 += e_s[l + 2] * __ldg(mrow + (size_t)(l + 2) * p.E);
            c3 += e_s[l + 3] * __ldg(mrow + (size_t)(l + 3) * p.E);
        }
        for (; l < L; ++l) c0 += e_s[l] * __ldg(mrow + (size_t)l * p.E);
        // fixed summation tree; matches the sequential reference to fp32 rounding
        const float acc = (c0 + c1) + (c2 + c3);
        p.ctx1[(size_t)b * p.ld1 + e] = acc;
        p.ctx2[(size_t)b * p.ld2 + e] = acc;
        p.ctx3[(size_t)b * p.ld3 + e] = acc;
    }
}

// End of a step: mel frame (+ bias) -> output [max_steps][B][M] and the next prenet input; stop-gate logic
// (decoder.py:381-395): dec = sigmoid(gate) <= threshold; not_finished *= dec; mel_lengths += not_finished.
__global__ void ker_infer_finish(const float* __restrict__ mel_raw, const float* __restrict__ bp, const float* __restrict__ gate_raw,
                                 const float* __restrict__ bg, float* mel_tm, float* frame, int* not_finished, int* mel_lengths,
                                 int B, int M, float threshold, int early, int max_steps, int* state) {
    if (state[1]) return;
    __shared__ int any;
    const int t = state[0];
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < B * M; i += blockDim.x) {
        const float v = mel_raw[i] + bp[i % M];
        mel_tm[(size_t)t * B * M + i] = v;
        frame[i] = v;
    }
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float g = gate_raw[b] + bg[0];
        const int dec = (1.f / (1.f + expf(-g))) <= threshold ? 1 : 0;
        const int nf = not_finished[b] * dec;
        not_finished[b] = nf;
        mel_lengths[b] += nf;
        if (nf) atomicOr(&any, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        state[2] = t + 1;
        if ((early && !any) || t + 1 >= max_steps) state[1] = 1;
        state[0] = t + 1;
    }
}

// [T'][B][M] -> [B][M][ld] (reference layout with row stride ld >= T')
__global__ void ker_tm_to_ref_ld(const float* __restrict__ x_tm, float* out, int T, int B, int M, int ld) {
    const int64_t n = (int64_t)T * B * M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), b = (int)((i / M) % B), t = (int)(i / ((int64_t)M * B));
        out[((size_t)b * M + m) * ld + t] = x_tm[i];
    }
}
// [B][T'][M] -> [B][M][ld]
__global__ void ker_bt_to_ref_ld(const float* __restrict__ x_bt, float* out, int B, int T, int M, int ld) {
    const int64_t n = (int64_t)B * T * M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), t = (int)((i / M) % T), b = (int)(i / ((int64_t)M * T));
        out[((size_t)b * M + m) * ld + t] = x_bt[i];
    }
}

__global__ void ker_fill_ones_i32(int* p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 1;
}
int k_fill_ones_i32(int* p, int n, cudaStream_t st) {
    ker_fill_ones_i32<<<cdiv(n, 256), 256, 0, st>>>(p, n);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_infer_relu_drop(float* x, int ld, const uint8_t* masks, int layer, int B, int N, const int* state, cudaStream_t st) {
    ker_infer_relu_drop<<<cdiv((int64_t)B * N, 256), 256, 0, st>>>(x, ld, masks, layer, B, N, state);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_infer_lstm_point(const float* z, const float* b_ih, const float* b_hh, float* c, float* h1, int ld1, float* h2, int ld2,
                       int B, int H, const int* state, cudaStream_t st) {
    ker_infer_lstm_point<<<cdiv((int64_t)B * H, 256), 256, 0, st>>>(z, b_ih, b_hh, c, h1, ld1, h2, ld2, B, H, state);
    MSA_LAUNCH_CHECK();
    return 0;
}
size_t infer_attention_smem(int L, int A, int F, int Kl) {
    return sizeof(float) * ((size_t)2 * (L + Kl - 1) + 2 * A + (size_t)2 * Kl * F + (size_t)F * A + (size_t)L * (F + 1) + L + 48);
}
int k_infer_attention(const InferAttnParams& p, cudaStream_t st) {
    const size_t smem = infer_attention_smem(p.L, p.A, p.F, p.Kl);
    MSA_CHECK(smem <= 200 * 1024, MSA_E_UNSUPPORTED, "infer attention: text length %d too long for the shared-memory tile", p.L);
    if (smem > 48 * 1024) MSA_CUDA(cudaFuncSetAttribute(ker_infer_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ker_infer_attention<<<p.B, kInferAttnThreads, smem, st>>>(p);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_infer_finish(const float* mel_raw, const float* bp, const float* gate_raw, const float* bg, float* mel_tm, float* frame,
                   int* not_finished, int* mel_lengths, int B, int M, float threshold, int early, int max_steps, int* state,
                   cudaStream_t st) {
    ker_infer_finish<<<1, 256, 0, st>>>(mel_raw, bp, gate_raw, bg, mel_tm, frame, not_finished, mel_lengths, B, M, threshold, early,
                                        max_steps, state);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_tm_to_ref_ld(const float* x_tm, float* out, int T, int B, int M, int ld, cudaStream_t st) {
    ker_tm_to_ref_ld<<<cdiv((int64_t)T * B * M, 256) > 2368 ? 2368 : cdiv((int64_t)T * B * M, 256), 256, 0, st>>>(x_tm, out, T, B, M, ld);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_bt_to_ref_ld(const float* x_bt, float* out, int B, int T, int M, int ld, cudaStream_t st) {
    ker_bt_to_ref_ld<<<cdiv((int64_t)T * B * M, 256) > 2368 ? 2368 : cdiv((int64_t)T * B * M, 256), 256, 0, st>>>(x_bt, out, B, T, M, ld);
    MSA_LAUNCH_CHECK();
    return 0;
}

}  // namespace msa
