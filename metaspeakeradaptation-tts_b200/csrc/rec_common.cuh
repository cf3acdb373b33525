// Device building blocks shared by the persistent recurrent kernels (lstm_rec.cu, attn_chain.cu).
//
// Work decomposition (forward): the 4H gate rows of an LSTMCell are split by HIDDEN UNIT across the
// co-resident CTAs of a cooperative launch; CTA c owns units [u0,u1) and keeps the 4*(u1-u0) rows
// of W_hh resident in shared memory (fp32) for all T steps.  Each step is a skinny mat-vec
// (rows x K) . (K x B) from shared memory, the LSTM point-wise update for the owned units and a
// store of the owned slice of h into the per-step stash array, which the other CTAs poll (common.cuh).
// Backward: CTA c owns the same units and keeps the matching COLUMNS of W_hh (i.e. rows of W_hh^T)
// resident; dh_rec[u] = sum_r W_hh[r][u] * dz[r] is computed with each thread owning a slice of r.
//
// Everything in the step loop is written for LATENCY, not throughput: operands are zero-padded in shared
// memory so that the inner loops have no bounds checks or divergent branches, all loads of a tile are issued
// before the first FMA, reductions are warp-shuffle trees, and every phase spreads over all 512 threads.
#pragma once
#include "common.cuh"

namespace msa {

constexpr int kRecThreads = 512;
constexpr int kRecWarps = kRecThreads / 32;
constexpr int kUMax = 8;        // max hidden units per CTA (=> 32 gate rows, one transpose-reduce)
constexpr int kBMax = 32;       // max batch rows of one pass through the recurrent kernels (also: 4*units*B <= threads)
constexpr int kBTiles = kBMax / 4;

__host__ __device__ inline int round_up_i(int n, int m) { return (n + m - 1) / m * m; }

// Development instrumentation (profiles/): off unless a buffer is passed.
//   prof : [grid][8]  cycles per phase summed over the steps, thread 0 of each CTA (accumulated in global memory so that it costs
//          no registers in the production path)
//   trace: [grid][kRecWarps][kTraceSteps][kTraceTags][2]  {clock64, globaltimer} of lane 0 of every warp at every mark of the
//          steps [trace_t0, trace_t0 + kTraceSteps)
constexpr int kProfSlots = 8, kTraceSteps = 4, kTraceTags = 12;
template <bool kOn>
struct ChainProf {
    long long* out;
    long long* trace;
    long long last;
    int t0;
    __device__ __forceinline__ void start(long long* prof_buf, long long* trace_buf, int trace_t0) {
        if (!kOn) return;
        out = threadIdx.x == 0 ? prof_buf : nullptr;
        trace = (threadIdx.x & 31) == 0 ? trace_buf : nullptr;
        t0 = trace_t0;
        if (out) {
            for (int i = 0; i < kProfSlots; ++i) out[(size_t)blockIdx.x * kProfSlots + i] = 0;
            last = clock64();
        }
    }
    __device__ __forceinline__ void mark(int i, int step) {
        if (!kOn) return;
        if (out) {
            const long long c = clock64();
            out[(size_t)blockIdx.x * kProfSlots + (i < kProfSlots ? i : kProfSlots - 1)] += c - last;
            last = c;
        }
        if (trace && step >= t0 && step < t0 + kTraceSteps) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            long long* e = trace + ((((size_t)blockIdx.x * kRecWarps + (threadIdx.x >> 5)) * kTraceSteps + (step - t0)) * kTraceTags + i) * 2;
            e[0] = clock64();
            e[1] = (long long)gt;
        }
    }
};

// exp2-based activations (ex2.approx + fast division): absolute error ~1e-7, far inside the fp32 parity tolerance,
// and a fraction of the dependent-instruction latency of expf/tanhf on the per-step critical path
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

struct LstmGrad {
    float di, df, dg, do_, dc_prev;
};
// dh: grad w.r.t. the (pre-dropout) hidden output; dc_in: grad carried from the later step.
// Returns gate pre-activation grads and the carry for the earlier step.
__device__ __forceinline__ LstmGrad lstm_point_bwd(float i, float f, float g, float o, float c, float cprev, float dh,
                                                   float dc_in) {
    LstmGrad r;
    const float tc = fast_tanh(c);
    r.do_ = dh * tc * o * (1.f - o);
    const float dc = dc_in + dh * o * (1.f - tc * tc);
    r.di = dc * g * i * (1.f - i);
    r.dg = dc * i * (1.f - g * g);
    r.df = dc * cprev * f * (1.f - f);
    r.dc_prev = dc * f;
    return r;
}

// Balanced partition of n items over parts: [lo, hi) of part i.
__device__ __host__ __forceinline__ int part_lo(int i, int n, int parts) { return (int)(((long long)i * n) / parts); }

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ---- forward mat-vec ------------------------------------------------------------------------------------------------
// part[((tile*KS + ks)*RG + rg)*32 + rr*4 + bl] = partial over the K-chunks (of 128 columns) j0 <= j < j1 that warp (rg, ks)
// handles (j = j0 + ks, j0 + ks + KS, ...) of   sum_k Wsm[(rg*8+rr)*KP + k] * hs[(tile*4+bl)*KP + k].
// Wsm [RG*8][KP] and hs [BP][KP] are zero-padded (KP % 128 == 0, BP % 4 == 0).  RG = ceil(rows/8), KS = (NT / 32) / RG.
// No __syncthreads inside: the caller synchronises and then sums the KS partials of each output (lstm_gate_sum).
template <int NT, bool kGlobalH = false>
__device__ __forceinline__ void cta_matvec_fwd(const float* __restrict__ Wsm, int RG, int KP, const float* __restrict__ hs, int B,
                                               float* __restrict__ part, int j0, int j1, int Hg = 0, SpinGuard* sg = nullptr) {
    // kGlobalH: hs is the [B][Hg] row block other CTAs are publishing in global memory (canary-polled straight into
    // registers, no staging pass through shared memory); otherwise the zero-padded [BP][KP] copy in shared memory
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int KS = (NT / 32) / RG;
    const int rg = w % RG, ks = w / RG;
    if (ks >= KS) return;
    const int KP4 = KP >> 2;
    const float4* W4 = reinterpret_cast<const float4*>(Wsm) + (size_t)(rg * 8) * KP4;
    const float4* h4p = reinterpret_cast<const float4*>(hs);
    const int ntile = (B + 3) >> 2;
    for (int tile = 0; tile < ntile; ++tile) {
        if (kGlobalH) {
            // one pass, 32 accumulators: every h word is fetched from L2 exactly once per warp
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            for (int j = j0 + ks; j < j1; j += KS) {
                const int c4 = j * 32 + lane;
                const bool in = c4 < (Hg >> 2);
                float4 h4[4];
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    h4[b] = (in && tile * 4 + b < B) ? ld_poll4(hs + (size_t)(tile * 4 + b) * Hg + (size_t)c4 * 4)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                unsigned int cmax = 0u;
#pragma unroll
                for (int b = 0; b < 4; ++b) cmax = umax_acc(cmax, h4[b]);
                if (cmax == kCanary) {      // some word arrived before it was published: exact test + re-poll per load
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (in && tile * 4 + b < B) {
                            sg->reset();
                            while (!ready4(h4[b])) {
                                if (sg->bail()) break;
                                h4[b] = ld_poll4(hs + (size_t)(tile * 4 + b) * Hg + (size_t)c4 * 4);
                            }
                        }
                    }
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 w4[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) w4[rr] = W4[(size_t)(half * 4 + rr) * KP4 + c4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[(half * 4 + rr) * 4 + b] += dot4(w4[rr], h4[b]);
                }
            }
            part[((size_t)(tile * KS + ks) * RG + rg) * 32 + lane] = warp_transpose_reduce32(acc);
            continue;
        }
        // two passes of 4 gate rows x 4 batch rows: 16 accumulators + 16 + 16 operand registers live at a time
        float tot = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            for (int j = j0 + ks; j < j1; j += KS) {
                const int c4 = j * 32 + lane;
                float4 h4[4], w4[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) h4[b] = h4p[(size_t)(tile * 4 + b) * KP4 + c4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) w4[rr] = W4[(size_t)(half * 4 + rr) * KP4 + c4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[rr * 4 + b] += dot4(w4[rr], h4[b]);
            }
            const float th = warp_transpose_reduce16(acc);      // lanes i, i+16: total of accumulator i
            if ((lane >> 4) == half) tot = th;                   // lane = (half*4 + rr)*4 + b, as with 32 accumulators
        }
        part[((size_t)(tile * KS + ks) * RG + rg) * 32 + lane] = tot;
    }
}
// pre-activation of gate row rl (local row index) for batch row b: sum of the KS partials
template <int NT>
__device__ __forceinline__ float lstm_gate_sum(const float* part, int RG, int rl, int b) {
    const int KS = (NT / 32) / RG;
    const float* p = part + ((size_t)((b >> 2) * KS) * RG + (rl >> 3)) * 32 + (rl & 7) * 4 + (b & 3);
    float s = 0.f;
    for (int q = 0; q < KS; ++q) s += p[(size_t)q * RG * 32];
    return s;
}
// Context term of the attention LSTM: zm[rl*BP + b] = sum_l MWs[(rl*B + b)*LP + l] * as[b*LP + l] for the nrows (multiple
// of 8) local gate rows.  Warp w takes rows w, w+16, ...; lane = (b-in-tile, l-eighth); 3-shuffle reduction.  No sync inside.
template <int NT>
__device__ __forceinline__ void cta_context_term(const float* __restrict__ MWs, const float* __restrict__ as_, int nrows, int B,
                                                 int BP, int LP, float* __restrict__ zm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int bl = lane >> 3, lq = lane & 7, LP4 = LP >> 2;
    for (int tile = 0; tile < ((B + 3) >> 2); ++tile) {
        const int b = tile * 4 + bl;
        for (int rl = w; rl < nrows; rl += (NT / 32)) {
            float acc = 0.f;
            if (b < B) {
                const float4* m4 = reinterpret_cast<const float4*>(MWs) + ((size_t)rl * B + b) * LP4;
                const float4* a4 = reinterpret_cast<const float4*>(as_) + (size_t)b * LP4;
                for (int l4 = lq; l4 < LP4; l4 += 8) acc += dot4(m4[l4], a4[l4]);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (lq == 0 && b < BP) zm[rl * BP + b] = acc;
        }
    }
}

// Copy rows x n4 float4 words that OTHER CTAs publish during this launch (canary-polled, common.cuh) from global
// (row stride n4 float4) to shared (row stride dst_stride4 float4).  All loads of a thread are issued before the first
// one is checked (memory-level parallelism while spinning).
template <int NT>
__device__ __forceinline__ void poll_copy_rows(float* dst_smem, int dst_stride4, const float* src, int rows, int n4, SpinGuard& sg) {
    constexpr int kBatch = 4;
    const int total = rows * n4;
    for (int base = threadIdx.x; base < total; base += kBatch * NT) {
        float4 v[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int idx = base + j * NT;
            if (idx < total) v[j] = ld_poll4(src + (size_t)idx * 4);
        }
        unsigned int cmax = 0u;
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
            if (base + j * NT < total) cmax = umax_acc(cmax, v[j]);
        if (cmax == kCanary) {      // some word arrived before it was published: exact test + re-poll per load
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int idx = base + j * NT;
                if (idx < total) {
                    sg.reset();
                    while (!ready4(v[j])) {
                        if (sg.bail()) break;
                        v[j] = ld_poll4(src + (size_t)idx * 4);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int idx = base + j * NT;
            if (idx < total) {
                const int r = idx / n4, c = idx - r * n4;
                reinterpret_cast<float4*>(dst_smem)[(size_t)r * dst_stride4 + c] = v[j];
            }
        }
    }
}

// Same term when the MW slice does not fit in shared memory: rows are read from the global [4Ha][B*L] matrix (gate-major rows,
// L2-resident after the first step).  grow(rl) = global row of local row rl.
template <int NT, class RowFn>
__device__ __forceinline__ void cta_context_term_global(const float* __restrict__ MW, RowFn grow, const float* __restrict__ as_,
                                                        int nrows, int B, int BP, int L, int LP, float* __restrict__ zm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int bl = lane >> 3, lq = lane & 7;
    for (int tile = 0; tile < ((B + 3) >> 2); ++tile) {
        const int b = tile * 4 + bl;
        for (int rl = w; rl < nrows; rl += (NT / 32)) {
            float acc = 0.f;
            if (b < B) {
                const float* m = MW + (size_t)grow(rl) * B * L + (size_t)b * L;
                for (int l = lq; l < L; l += 8) acc += __ldg(m + l) * as_[b * LP + l];
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (lq == 0 && b < BP) zm[rl * BP + b] = acc;
        }
    }
}

// ---- gated hand-offs ------------------------------------------------------------------------------------------------
// Spinning on not-yet-written data with all 512 threads of all 148 CTAs saturates the L2 request bandwidth and delays the
// very stores everybody is waiting for.  For the large gathers (h, dz) the first ncta threads therefore wait for ONE sentinel
// word per producer CTA -- the last word that producer stores -- while the other warps park at the following __syncthreads
// (no memory traffic), and then all threads fetch the bulk once (still canary-checked word by word: the sentinel is a hint,
// never the correctness argument).  Small gathers (energies, queries, d a, dq) can be fetched whole by warp 0 (kFlagWarp0).
template <class AddrFn>
__device__ __forceinline__ void gate_wait(int n, AddrFn addr, SpinGuard& sg) {
    // one sentinel per thread (the first n threads): every sentinel is polled independently, one L2 round trip each
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* a = addr(i);
        if (a) (void)poll1(a, sg);
    }
}
// warp 0 copies n words that other CTAs publish during this launch from global to shared, polling each word
__device__ __forceinline__ void gather_small(float* dst_smem, const float* src, int n, SpinGuard& sg) {
    if (threadIdx.x >= 32) return;
    constexpr int kBatch = 8;
    for (int base = 0; base < n; base += 32 * kBatch) {
        float v[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int i = base + j * 32 + threadIdx.x;
            if (i < n) v[j] = ld_poll(src + i);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int i = base + j * 32 + threadIdx.x;
            if (i < n) {
                sg.reset();
                while (is_canary(v[j])) {
                    if (sg.bail()) break;
                    v[j] = ld_poll(src + i);
                }
                dst_smem[i] = v[j];
            }
        }
    }
}

// n words published by other CTAs -> shared: by warp 0 alone (the others park at the caller's barrier) or by all threads
// (128-bit polls when the block is 16-byte aligned: a quarter of the L2 requests on the few hot lines)
template <int NT>
__device__ __forceinline__ void gather_words(float* dst_smem, const float* src, int n, bool warp0_only, SpinGuard& sg) {
    if (warp0_only) {
        gather_small(dst_smem, src, n, sg);
    } else if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_smem)) & 15) == 0) {
        for (int i = threadIdx.x; i < (n >> 2); i += NT) reinterpret_cast<float4*>(dst_smem)[i] = poll4(src + (size_t)i * 4, sg);
    } else {
        for (int i = threadIdx.x; i < n; i += NT) dst_smem[i] = poll1(src + i, sg);
    }
}
constexpr int kFlagGate = 1;    // large gathers: warp 0 waits for one sentinel per producer before the bulk fetch
constexpr int kFlagWarp0 = 2;   // small gathers: by warp 0 only

constexpr int kPairMax = 4;   // (b,l) positions per CTA whose context-path dot product rides along the backward mat-vec

// ---- backward mat-vec -----------------------------------------------------------------------------------------------
//   out[ul*BP + b] = sum_r WT[ul*R4 + r] * dz[b*R4 + r],  ul < 8 (rows of WT beyond the owned units are zero), b < B
// dz lives in global memory and is being published by the other CTAs (canary-polled).
// All threads must call.  `part` scratch: (NT / 32)*32 floats.
template <int NT>
__device__ __forceinline__ void cta_matvec_bwd(const float* __restrict__ WT, int R4, const float* dz, int B, float* part,
                                               float* out, int BP, SpinGuard& sg, int pass = 0, int npass = 1) {
    // pass / npass: the r range is split into npass interleaved slices (chunk index % npass == pass) so that the mat-vec can be
    // spread over the shadows of several hand-offs; pass 0 overwrites `out`, later passes accumulate into it
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nchunk = R4 >> 2;
    for (int bt = 0; bt < B; bt += 4) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        for (int ch = threadIdx.x * npass + pass; ch < nchunk; ch += NT * npass) {
            const int r = ch << 2;
            float4 d4[4];
#pragma unroll
            for (int b = 0; b < 4; ++b)
                d4[b] = (bt + b < B) ? ld_poll4(dz + (size_t)(bt + b) * R4 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned int cmax = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) cmax = umax_acc(cmax, d4[b]);
            if (cmax == kCanary) {      // some word arrived before it was published: exact test + re-poll per load
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (bt + b < B) {
                        sg.reset();
                        while (!ready4(d4[b])) {
                            if (sg.bail()) break;
                            d4[b] = ld_poll4(dz + (size_t)(bt + b) * R4 + r);
                        }
                    }
                }
            }
            float4 w4[kUMax];
#pragma unroll
            for (int ul = 0; ul < kUMax; ++ul) w4[ul] = *reinterpret_cast<const float4*>(WT + (size_t)ul * R4 + r);
#pragma unroll
            for (int ul = 0; ul < kUMax; ++ul)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[ul * 4 + b] += dot4(w4[ul], d4[b]);
        }
        const float tot = warp_transpose_reduce32(acc);
        part[w * 32 + lane] = tot;
        __syncthreads();
        if ((int)threadIdx.x < 32) {
            const int ul = threadIdx.x >> 2, b = threadIdx.x & 3;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < (NT / 32); ++q) s += part[q * 32 + threadIdx.x];
            if (bt + b < B) out[ul * BP + bt + b] = pass == 0 ? s : out[ul * BP + bt + b] + s;
        }
        __syncthreads();
    }
}
// Context-path recurrent term of the owned positions: pair_out[i] = sum_r MWp[i*mwp_stride + r] * dz[b_i*R4 + r] with
// b_i = (pair0+i)/L (MWp = rows pair0.. of the [B*L][R4] matrix, resident in shared memory or in global memory).
// All threads must call.  `red` scratch: (NT / 32)*kPairMax floats.
template <int NT>
__device__ __forceinline__ void cta_pair_dots(const float* __restrict__ MWp, size_t mwp_stride, const float* dz, int R4, int pair0,
                                              int npairs, int L, float* pair_out, float* red, SpinGuard& sg) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nchunk = R4 >> 2;
    for (int pbase = 0; pbase < npairs; pbase += kPairMax) {
        float accp[kPairMax];
#pragma unroll
        for (int i = 0; i < kPairMax; ++i) accp[i] = 0.f;
        for (int ch = threadIdx.x; ch < nchunk; ch += NT) {
            const int r = ch << 2;
            float4 d4[kPairMax];
#pragma unroll
            for (int i = 0; i < kPairMax; ++i)
                if (pbase + i < npairs) d4[i] = ld_poll4(dz + (size_t)((pair0 + pbase + i) / L) * R4 + r);
            unsigned int cmax = 0u;
#pragma unroll
            for (int i = 0; i < kPairMax; ++i)
                if (pbase + i < npairs) cmax = umax_acc(cmax, d4[i]);
            if (cmax == kCanary) {      // some word arrived before it was published: exact test + re-poll per load
#pragma unroll
                for (int i = 0; i < kPairMax; ++i) {
                    if (pbase + i < npairs) {
                        sg.reset();
                        while (!ready4(d4[i])) {
                            if (sg.bail()) break;
                            d4[i] = ld_poll4(dz + (size_t)((pair0 + pbase + i) / L) * R4 + r);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kPairMax; ++i) {
                if (pbase + i < npairs) {
                    accp[i] += dot4(*reinterpret_cast<const float4*>(MWp + (size_t)(pbase + i) * mwp_stride + r), d4[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kPairMax; ++i) {
            const float s = warp_sum(accp[i]);
            if (lane == 0) red[w * kPairMax + i] = s;
        }
        __syncthreads();
        if ((int)threadIdx.x < kPairMax && pbase + (int)threadIdx.x < npairs) {
            float s = 0.f;
            for (int q = 0; q < (NT / 32); ++q) s += red[q * kPairMax + threadIdx.x];
            pair_out[pbase + threadIdx.x] = s;
        }
        __syncthreads();
    }
}

}  // namespace msa
