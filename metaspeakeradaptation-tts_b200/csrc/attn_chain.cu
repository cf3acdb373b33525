// Persistent attention-RNN + location-sensitive attention chain over all T decoder steps.
//
// Restates the recurrent core of Decoder.decode (modules_tacotron2nv/decoder.py:253-258) and
// ForwardAttention.forward (forward_attn.py:121-131,178-219) for the teacher-forced pass:
//     z_a(t)  = XW[t] + Wc.ctx(t-1) + Whh.h_a'(t-1)           (attention LSTMCell, decoder.py:253-255)
//     h_a'(t) = dropout(h_a(t))                                (decoder.py:256, fed back)
//     e(t)    = v.tanh(Wq.h_a'(t) + loc(a(t-1), cum(t-1)) + PM) + b_v        (forward_attn.py:121-131)
//     a(t)    = softmax(e) | sigmoid(e)/sum                    (forward_attn.py:200-207; no padding mask, Q2)
// Because ctx(t-1) = sum_l a(t-1)[l] memory[l], the context term is evaluated as
// sum_l a(t-1)[l] * (Wc.memory[l]) with MW = Wc.memory^T precomputed by one GEMM: the context vector
// itself (and everything that consumes it: decoder-RNN input, mel/gate projection) leaves the
// sequential loop and becomes batched GEMMs (pass.cu).  What stays in the loop is exactly the
// dependency chain h_a -> q -> e -> a -> h_a.
//
// One cooperative launch (all CTAs co-resident), one CTA per SM.  CTA roles (every CTA plays all three):
//   unit owner : hidden units [u0,u1) of the attention LSTM, W_hh rows (+ their MW rows) resident in smem (fp32)
//   pair owner : (b,l) positions [p0,p1): location conv + dense + energy for those positions
//   query owner: attention dims [d0,d1): q[b][d] = Wq[d,:].h_a'[b,:]
// Three hand-offs per step (h_a, q, e), each ONE L2 round trip: the producer stores into the per-step stash array
// (ha[t], q[t], e[t] -- the arrays the backward pass reads anyway) and the consumers poll the data words themselves
// against the canary the host pre-filled (common.cuh).  No grid barrier, no atomics, no fences.
#include "rec_common.cuh"
#include "kernels.h"

namespace msa {

// thread count of the attention-chain kernels (the forward kernel carries a few register spills at 512 threads; 256 threads
// measured slower because the mat-vecs in the hand-off shadows take twice as long)
constexpr int kAttnThreads = 512;

struct AttnSmemFwd {
    size_t wsm, mws, hs, as_, ah, es, qs, part, part2, zm, wqs, wloc, wldT, vs, pm, pre, ctr, cf, apl, mta, wta, total;
    int KP, BP, LP, LH, CKP;
};
__host__ __device__ inline AttnSmemFwd attn_fwd_layout(int B, int L, int Ha, int A, int F, int Kl, int ncta, bool mw_res, bool fa = false) {
    AttnSmemFwd s;
    s.KP = round_up_i(Ha, 128);
    s.BP = (B + 3) & ~3;
    s.LP = round_up_i(L, 4);
    s.LH = round_up_i(L + Kl - 1, 4);
    s.CKP = 2 * Kl + 1;
    const int np_max = (B * L + ncta - 1) / ncta, nd_max = (A + ncta - 1) / ncta;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.wsm = take((size_t)4 * kUMax * s.KP);
    s.mws = take(mw_res ? (size_t)4 * kUMax * B * s.LP : 0);
    s.hs = take((size_t)s.BP * s.KP);
    s.as_ = take((size_t)B * s.LP);
    s.ah = take((size_t)2 * B * s.LH);
    s.es = take((size_t)B * L);
    s.qs = take((size_t)B * A);
    const int ntile = (B + 3) >> 2;
    s.part = take((size_t)ntile * kRecWarps * 32);
    s.part2 = take((size_t)ntile * kRecWarps * 32);
    s.zm = take((size_t)4 * kUMax * s.BP);
    s.wqs = take((size_t)nd_max * s.KP);
    s.wloc = take((size_t)F * s.CKP);
    s.wldT = take((size_t)F * A);
    s.vs = take(A);
    s.pm = take((size_t)np_max * A);
    s.pre = take((size_t)np_max * A);
    s.ctr = take((size_t)np_max * A);
    s.cf = take((size_t)np_max * F);
    s.apl = take(fa ? (size_t)B * L : 0);
    s.mta = take(fa ? (size_t)B * L : 0);
    s.wta = take(fa ? (size_t)s.KP : 0);
    s.total = o;
    return s;
}

template <bool kProf, int NT, bool kFa>
__global__ void __launch_bounds__(NT, 1) k_attn_chain_fwd(AttnChainParams p, int mw_res) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) float smem[];
    __shared__ float zn_s[kBMax];
    __shared__ float u_s[kBMax], uold_s[kBMax], fs_s[kBMax];     // forward attention: u(t-1) -> u(t), u used at t, sum alpha'
    const int T = p.T, B = p.B, L = p.L, Ha = p.Ha, A = p.A, F = p.F, Kl = p.Kl, H4 = 4 * Ha;
    const int BL = B * L, pl = (Kl - 1) / 2;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const AttnSmemFwd lay = attn_fwd_layout(B, L, Ha, A, F, Kl, ncta, mw_res != 0, kFa != 0);
    const int KP = lay.KP, KP4 = KP >> 2, LP = lay.LP, LH = lay.LH, CKP = lay.CKP;
    float* Wsm = smem + lay.wsm;       // [RG*8][KP]     W_hh rows of the owned units (local row = ul*4 + gate), zero-padded
    float* MWs = smem + lay.mws;       // [RG*8][B][LP]  (W_ih[:, prenet:] . memory^T) rows of the owned units
    float* hs = smem + lay.hs;         // [BP][KP]       h_a'(t-1), zero-padded
    float* as_ = smem + lay.as_;       // [B][LP]        a(t-1)
    float* ah = smem + lay.ah;         // [2][B][LH]     a(t-1) and cum(t-1) with a zero halo of pl on both sides (conv input)
    float* es = smem + lay.es;
    float* q_s = smem + lay.qs;
    float* part = smem + lay.part;     // W_hh . h_a'(t-1), first / second half of the K range (computed in the shadow
    float* part2 = smem + lay.part2;   //   of the q and e hand-offs of step t-1)
    float* zm = smem + lay.zm;         // [RG*8][BP] context term MW . a(t-1)
    float* wqs = smem + lay.wqs;
    float* wloc_s = smem + lay.wloc;   // [F][CKP]
    float* wldT = smem + lay.wldT;     // [F][A]
    float* vs = smem + lay.vs;
    float* pm_s = smem + lay.pm;       // [np][A] processed memory of the owned positions (constant over t)
    float* pre_s = smem + lay.pre;     // [np][A] loc + pm of the current step
    float* ctr_s = smem + lay.ctr;     // [np][A] v[d]*tanh(.) terms of the current step
    float* cf_s = smem + lay.cf;       // [np][F]
    float* apl_s = smem + lay.apl;     // [B][L]  forward attention: normalise(e(t)) of the current step
    float* mta_s = smem + lay.mta;     // [B][L]  memory . W_ta[:E]
    float* wta_s = smem + lay.wta;     // [KP]    W_ta[E:], zero-padded

    const int u0 = part_lo(cta, Ha, ncta), u1 = part_lo(cta + 1, Ha, ncta), U = u1 - u0, R = 4 * U;
    const int RG = R > 0 ? (R + 7) >> 3 : 1;
    const int p0 = part_lo(cta, BL, ncta), p1 = part_lo(cta + 1, BL, ncta), np = p1 - p0;
    const int d0 = part_lo(cta, A, ncta), d1 = part_lo(cta + 1, A, ncta), nd = d1 - d0;

    // ---- one-time staging of the resident operands ----
    for (int idx = threadIdx.x; idx < RG * 8 * KP4; idx += NT) {
        const int rl = idx / KP4, k4 = idx % KP4, g = rl & 3, ul = rl >> 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < R && k4 < (Ha >> 2)) v = __ldg(reinterpret_cast<const float4*>(p.whh + (size_t)(g * Ha + u0 + ul) * Ha) + k4);
        reinterpret_cast<float4*>(Wsm)[idx] = v;
    }
    if (mw_res) {
        for (int idx = threadIdx.x; idx < RG * 8 * B * LP; idx += NT) {
            const int rl = idx / (B * LP), j = idx % (B * LP), b = j / LP, l = j % LP, g = rl & 3, ul = rl >> 2;
            MWs[idx] = (rl < R && l < L) ? __ldg(p.mw_rm + (size_t)(g * Ha + u0 + ul) * BL + b * L + l) : 0.f;
        }
    }
    for (int idx = threadIdx.x; idx < nd * KP; idx += NT) {
        const int di = idx / KP, k = idx % KP;
        wqs[idx] = k < Ha ? __ldg(p.wq + (size_t)(d0 + di) * Ha + k) : 0.f;
    }
    for (int idx = threadIdx.x; idx < F * 2 * Kl; idx += NT) {
        const int f = idx / (2 * Kl), ck = idx % (2 * Kl);          // wloc[f][c][k] -> wloc_s[f][c*Kl+k]
        wloc_s[f * CKP + ck] = __ldg(p.wloc + idx);
    }
    for (int idx = threadIdx.x; idx < A * F; idx += NT) {
        const int d = idx / F, f = idx % F;                          // wld[d][f] -> wldT[f][d]
        wldT[f * A + d] = __ldg(p.wld + idx);
    }
    for (int idx = threadIdx.x; idx < A; idx += NT) vs[idx] = __ldg(p.v + idx);
    for (int idx = threadIdx.x; idx < np * A; idx += NT) pm_s[idx] = __ldg(p.pm + (size_t)p0 * A + idx);
    for (int idx = threadIdx.x; idx < B * LP; idx += NT) as_[idx] = 0.f;
    for (int idx = threadIdx.x; idx < 2 * B * LH; idx += NT) ah[idx] = 0.f;
    if (kFa) {
        for (int idx = threadIdx.x; idx < BL; idx += NT) mta_s[idx] = p.ta ? __ldg(p.mta + idx) : 0.f;
        for (int idx = threadIdx.x; idx < KP; idx += NT) wta_s[idx] = (p.ta && idx < Ha) ? __ldg(p.wta_h + idx) : 0.f;
        for (int idx = threadIdx.x; idx < kBMax; idx += NT) u_s[idx] = 0.5f;      // init_forward_attn (forward_attn.py:90-96)
    }
    __shared__ float bta_s;
    if (threadIdx.x == 0) bta_s = (kFa && p.ta) ? __ldg(p.bta) : 0.f;
    for (int idx = threadIdx.x; idx < lay.BP * KP; idx += NT) hs[idx] = 0.f;
    for (int i = threadIdx.x; i < np; i += NT) p.cum[p0 + i] = 0.f;       // cum fed to the conv at t = 0
    const float bv = __ldg(p.bv);

    // point-wise role of the LSTM: 4 lanes per cell (ul, b), lane g evaluates gate g; streaming inputs fetched one step ahead
    const bool pw = (int)threadIdx.x < 4 * U * B;
    const int g = threadIdx.x & 3, cell = threadIdx.x >> 2;
    const int ul = pw ? cell / B : 0, pb = pw ? cell % B : 0, u = u0 + ul;
    float xz = 0.f, cstate = 0.f;
    unsigned char mk = 1;
    auto fetch = [&](int t) {
        xz = __ldg(p.xw + ((size_t)t * B + pb) * H4 + (size_t)g * Ha + u);
        if (p.mask && g == 0) mk = p.mask[((size_t)t * B + pb) * Ha + u];
    };
    if (pw) fetch(0);
    SpinGuard sg(p.abort_word);
    ChainProf<kProf> prof;
    prof.start(p.prof, p.trace, p.trace_t0);
    __syncthreads();

    // Step schedule (critical path in CAPITALS, everything else runs in the shadow of a hand-off):
    //   CONTEXT TERM MW.a(t-1) -> POINT-WISE -> publish h(t) | location features of step t | GATHER h(t) -> Q -> publish q(t)
    //   | first half of W_hh.h(t) | GATHER q(t) -> ENERGIES -> publish e(t) | second half of W_hh.h(t) | GATHER e(t) -> SOFTMAX
    const int nK = KP >> 7, nKh = (nK + 1) >> 1, BPz = lay.BP;
    for (int idx = threadIdx.x; idx < ((B + 3) >> 2) * kRecWarps * 32; idx += NT) { part[idx] = 0.f; part2[idx] = 0.f; }
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        // ---- phase 1: attention LSTMCell for the owned units (decoder.py:253-256); publishes h_a'(t) ----
        if (U > 0) {
            if (mw_res) cta_context_term<NT>(MWs, as_, R, B, BPz, LP, zm);
            else cta_context_term_global<NT>(p.mw_rm, [&](int rl) { return (rl & 3) * Ha + u0 + (rl >> 2); }, as_, R, B, BPz, L, LP, zm);
        }
        prof.mark(0, t);
        __syncthreads();
        if (pw) {
            const size_t zb = ((size_t)t * B + pb) * H4;
            const int rl = ul * 4 + g;
            float z = lstm_gate_sum<NT>(part, RG, rl, pb) + lstm_gate_sum<NT>(part2, RG, rl, pb) + xz;
            z += zm[rl * BPz + pb];
            const float act = g == 2 ? fast_tanh(z) : fast_sigmoid(z);
            const unsigned int gm = 0xFu << (threadIdx.x & 28);
            const float ai = __shfl_sync(gm, act, 0, 4), af = __shfl_sync(gm, act, 1, 4);
            const float ag = __shfl_sync(gm, act, 2, 4), ao = __shfl_sync(gm, act, 3, 4);
            if (g == 0) {
                const float cn = af * cstate + ai * ag;
                cstate = cn;
                float hv = ao * fast_tanh(cn);
                if (p.mask) hv = mk ? hv * p.drop_scale : 0.f;
                st_pub(p.ha + ((size_t)t * B + pb) * Ha + u, hv);
                p.ca[((size_t)t * B + pb) * Ha + u] = cn;
            }
            p.ga[zb + (size_t)g * Ha + u] = act;
            if (t + 1 < T) fetch(t + 1);
        }
        prof.mark(1, t);
        // ---- phase 1a: location features of the owned (b,l) positions (forward_attn.py:121-127); overlaps the h hand-off ----
        // conv: item (pi, f, ks) sums taps ck = ks, ks+8, ... ; 8 consecutive lanes share (pi, f)
        for (int base = 0; base < np * F * 8; base += NT) {     // warp-uniform trip count (full-mask shuffles inside)
            const int it = base + threadIdx.x;
            const bool valid = it < np * F * 8;
            const int ks = it & 7, f = valid ? (it >> 3) % F : 0, pi = valid ? (it >> 3) / F : 0;
            const int pp = p0 + pi, b = pp / L, l = pp - b * L;
            float cf = 0.f;
            if (valid)
                for (int ck = ks; ck < 2 * Kl; ck += 8) {
                    const int c = ck >= Kl ? 1 : 0, k = ck - c * Kl;
                    cf += wloc_s[f * CKP + ck] * ah[(c * B + b) * LH + l + k];
                }
            cf += __shfl_xor_sync(0xffffffffu, cf, 1);
            cf += __shfl_xor_sync(0xffffffffu, cf, 2);
            cf += __shfl_xor_sync(0xffffffffu, cf, 4);
            if (valid && ks == 0) {
                cf_s[pi * F + f] = cf;
                p.convf[((size_t)t * BL + pp) * F + f] = cf;
            }
        }
        prof.mark(2, t);
        __syncthreads();
        for (int it = threadIdx.x; it < np * A; it += NT) {
            const int pi = it / A, d = it - pi * A;
            float l0 = 0.f, l1 = 0.f;
            int f = 0;
            for (; f + 1 < F; f += 2) {
                l0 += wldT[f * A + d] * cf_s[pi * F + f];
                l1 += wldT[(f + 1) * A + d] * cf_s[pi * F + f + 1];
            }
            if (f < F) l0 += wldT[f * A + d] * cf_s[pi * F + f];
            pre_s[it] = l0 + l1 + pm_s[it];
        }
        prof.mark(3, t);
        // ---- hand-off 1: all of h_a'(t) -> shared (operand of q now and of the mat-vec for step t+1) ----
        if (p.flags & kFlagGate) {
            const float* hrow = p.ha + ((size_t)t * B + (B - 1)) * Ha;
            gate_wait(ncta, [&](int c) { const int e = part_lo(c + 1, Ha, ncta); return e > 0 ? hrow + e - 1 : nullptr; }, sg);
            __syncthreads();
        }
        poll_copy_rows<NT>(hs, KP4, p.ha + (size_t)t * B * Ha, B, Ha >> 2, sg);
        prof.mark(4, t);
        __syncthreads();
        // ---- phase 2: query projection for the owned attention dims (forward_attn.py:125); publishes q(t) ----
        for (int it = NW - 1 - w; it < nd * B; it += NW) {     // top warps: warp 0 is the gatherer
            const int di = it / B, b = it - di * B;
            const float4* w4p = reinterpret_cast<const float4*>(wqs) + (size_t)di * KP4;
            const float4* h4p = reinterpret_cast<const float4*>(hs) + (size_t)b * KP4;
            float a0 = 0.f, a1 = 0.f;
            int k4 = lane;
            for (; k4 + 32 < KP4; k4 += 64) {
                a0 += dot4(w4p[k4], h4p[k4]);
                a1 += dot4(w4p[k4 + 32], h4p[k4 + 32]);
            }
            if (k4 < KP4) a0 += dot4(w4p[k4], h4p[k4]);
            const float sacc = warp_sum(a0 + a1);
            if (lane == 0) st_pub(p.q + ((size_t)t * B + b) * A + d0 + di, sacc);
        }
        prof.mark(5, t);
        // in the shadow of the q hand-off: first half of W_hh . h_a'(t) for step t+1
        if (U > 0) cta_matvec_fwd<NT>(Wsm, RG, KP, hs, B, part, 0, nKh);
        // ---- phase 3 (after hand-off 2: q): energies for the owned positions (forward_attn.py:128-131); publishes e(t) ----
        gather_words<NT>(q_s, p.q + (size_t)t * B * A, B * A, (p.flags & kFlagWarp0) != 0, sg);
        prof.mark(6, t);
        __syncthreads();
        for (int it = threadIdx.x; it < np * A; it += NT) {
            const int pi = it / A, d = it - pi * A, pp = p0 + pi, b = pp / L;
            const float sv = fast_tanh(q_s[b * A + d] + pre_s[it]);
            p.s[((size_t)t * BL + pp) * A + d] = sv;
            ctr_s[it] = vs[d] * sv;
        }
        prof.mark(7, t);
        __syncthreads();
        for (int pi = w; pi < np; pi += NW) {
            float e = 0.f;
            for (int d = lane; d < A; d += 32) e += ctr_s[pi * A + d];
            e = warp_sum(e);
            if (lane == 0) st_pub(p.e + (size_t)t * BL + p0 + pi, e + bv);
        }
        prof.mark(8, t);
        // in the shadow of the e hand-off: second half of W_hh . h_a'(t)
        if (U > 0) cta_matvec_fwd<NT>(Wsm, RG, KP, hs, B, part2, nKh, nK);
        // ---- hand-off 3: all energies of step t; a(t) = normalise(e(t)); cum += a(t) (forward_attn.py:200-210) ----
        gather_words<NT>(es, p.e + (size_t)t * BL, BL, (p.flags & kFlagWarp0) != 0, sg);
        prof.mark(9, t);
        __syncthreads();
        for (int b = w; b < B; b += NW) {     // one row per warp: softmax or sigmoid/sum
            float m = 0.f;
            if (p.norm == 0) {
                m = -INFINITY;
                for (int l = lane; l < L; l += 32) m = fmaxf(m, es[b * L + l]);
                m = warp_max(m);
            }
            float sum = 0.f;
            for (int l = lane; l < L; l += 32) {
                const float x = p.norm == 0 ? __expf(es[b * L + l] - m) : fast_sigmoid(es[b * L + l]);
                es[b * L + l] = x;
                sum += x;
            }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            if (!kFa) {
                for (int l = lane; l < L; l += 32) {
                    const float a = es[b * L + l] * inv;
                    as_[b * LP + l] = a;
                    ah[(0 * B + b) * LH + pl + l] = a;
                    ah[(1 * B + b) * LH + pl + l] += a;
                }
            } else {
                // forward attention (forward_attn.py:154-176, training: no mask): alpha' = ((1-u) alpha + u shift(alpha) + 1e-8) * a,
                // alpha(t) = alpha' / sum(alpha'); the location conv sees prev = alpha(t) and cum = sum of the PLAIN a (208-210)
                const float u = u_s[b];
                float S = 0.f;
                for (int l = lane; l < L; l += 32) {
                    const float a = es[b * L + l] * inv;
                    apl_s[b * L + l] = a;
                    ah[(1 * B + b) * LH + pl + l] += a;
                    // alpha(-1) = [1, 1e-7, ...] (as_ is zero before the first step)
                    const float al = t > 0 ? as_[b * LP + l] : (l == 0 ? 1.f : 1e-7f);
                    const float sh = l > 0 ? (t > 0 ? as_[b * LP + l - 1] : (l == 1 ? 1.f : 1e-7f)) : 0.f;
                    const float ap = (((1.f - u) * al + u * sh) + 1e-8f) * a;
                    es[b * L + l] = ap;
                    S += ap;
                }
                S = warp_sum(S);          // (also orders the reads of alpha(t-1) above before the writes below)
                float zu = 0.f;
                for (int l = lane; l < L; l += 32) {
                    const float al = es[b * L + l] / S;
                    as_[b * LP + l] = al;
                    ah[(0 * B + b) * LH + pl + l] = al;
                    zu += al * mta_s[b * L + l];
                }
                if (lane == 0) { uold_s[b] = u; fs_s[b] = S; }
                if (p.ta) {
                    // transition agent u(t) = sigmoid(W_ta . [ctx(t); h_a'(t)] + b) with ctx(t) = alpha(t) . memory (222-224)
                    for (int k = lane; k < KP; k += 32) zu += wta_s[k] * hs[(size_t)b * KP + k];
                    zu = warp_sum(zu);
                    if (lane == 0) u_s[b] = fast_sigmoid(zu + bta_s);
                }
            }
            if (lane == 0) zn_s[b] = sum;
        }
        prof.mark(10, t);
        __syncthreads();
        for (int i = threadIdx.x; i < np; i += NT) {
            const int pp = p0 + i, b = pp / L, l = pp - b * L;
            p.align[(size_t)t * BL + pp] = as_[b * LP + l];
            if (kFa) p.aplain[(size_t)t * BL + pp] = apl_s[pp];
            if (t + 1 < T) p.cum[(size_t)(t + 1) * BL + pp] = ah[(1 * B + b) * LH + pl + l];
        }
        if (cta == 0 && (int)threadIdx.x < B) {
            p.znorm[(size_t)t * B + threadIdx.x] = zn_s[threadIdx.x];
            if (kFa) {
                p.fsum[(size_t)t * B + threadIdx.x] = fs_s[threadIdx.x];
                p.ustash[(size_t)t * B + threadIdx.x] = uold_s[threadIdx.x];
            }
        }
        prof.mark(11, t);
    }
}

// =====================================================================================
// backward
struct AttnSmemBwd {
    size_t wt, mwp, part, dhs, wqT, das, als, des, tq, wldT, wloc, vs, dss, gcum, dprev, pout, dqs, qd, cpart, apl, apr, drec, dws, mta, total;
    int BP, AP;
};
__host__ __device__ inline AttnSmemBwd attn_bwd_layout(int B, int L, int Ha, int A, int F, int Kl, int ncta, bool mwp_res, bool fa = false) {
    AttnSmemBwd s;
    s.BP = (B + 3) & ~3;
    s.AP = A + 1;
    const int np_max = (B * L + ncta - 1) / ncta;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.wt = take((size_t)kUMax * 4 * Ha);
    s.mwp = take(mwp_res ? (size_t)np_max * 4 * Ha : 0);
    s.part = take(kRecWarps * 32);
    s.dhs = take(kUMax * s.BP);
    s.wqT = take((size_t)kUMax * A);
    s.das = take((size_t)B * L * (fa ? 2 : 1));
    s.als = take((size_t)B * L);
    s.des = take((size_t)B * L);
    s.tq = take((size_t)B * L);
    s.wldT = take((size_t)F * s.AP);
    s.wloc = take((size_t)F * (2 * Kl + 1));
    s.vs = take(A);
    s.dss = take((size_t)np_max * A);
    s.gcum = take(np_max);
    s.dprev = take(np_max);
    s.pout = take(np_max);
    s.dqs = take((size_t)B * A);
    s.qd = take((size_t)kUMax * kBMax);
    s.cpart = take((size_t)np_max * Kl * 2);
    const size_t nfa = fa ? (size_t)B * L : 0;
    s.apl = take(nfa);
    s.apr = take(nfa);
    s.drec = take(nfa);
    s.dws = take(nfa);
    s.mta = take(nfa);
    s.total = o;
    return s;
}

template <bool kProf, int NT, bool kFa>
__global__ void __launch_bounds__(NT, 1) k_attn_chain_bwd(AttnChainBwdParams p, int mwp_res) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[NW * kPairMax];
    __shared__ float zn_s[kBMax];
    __shared__ float us_s[kBMax], fs_s[kBMax], dzu_s[2][kBMax];      // forward attention: u used at t, sum alpha'(t), d zu by step parity
    const int T = p.T, B = p.B, L = p.L, Ha = p.Ha, A = p.A, F = p.F, Kl = p.Kl, H4 = 4 * Ha;
    const int BL = B * L, pl = (Kl - 1) / 2, CKP = 2 * Kl + 1;
    const int DS = kFa ? 2 * BL : BL;      // words of the d a hand-off per step
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const AttnSmemBwd lay = attn_bwd_layout(B, L, Ha, A, F, Kl, ncta, mwp_res != 0, kFa != 0);
    const int BP = lay.BP, AP = lay.AP;
    float* WT = smem + lay.wt;         // [kUMax][4Ha]  WT[ul][r] = W_hh[r][u0+ul], zero rows beyond the owned units
    float* MWp = smem + lay.mwp;       // [np][4Ha]     rows of (memory . Wc^T) of the owned positions
    float* part = smem + lay.part;
    float* dhs = smem + lay.dhs;
    float* wqT = smem + lay.wqT;       // [kUMax][A]    Wq columns of the owned units
    float* das = smem + lay.das;
    float* als = smem + lay.als;
    float* des = smem + lay.des;
    float* tq_s = smem + lay.tq;
    float* wldT = smem + lay.wldT;     // [F][A+1]
    float* wloc_s = smem + lay.wloc;   // [F][2Kl+1]
    float* vs = smem + lay.vs;
    float* ds_s = smem + lay.dss;      // [np][A]
    float* gcum_s = smem + lay.gcum;
    float* dprev_s = smem + lay.dprev;
    float* pout = smem + lay.pout;
    float* dq_s = smem + lay.dqs;
    float* qd_s = smem + lay.qd;       // [U*B]
    float* cpart = smem + lay.cpart;   // [np*Kl][2]
    float* apl_s = smem + lay.apl;     // forward attention: [B][L] plain a(t), alpha(t-1), d alpha(t) from step t+1, d w scratch, mta
    float* apr_s = smem + lay.apr;
    float* drec_s = smem + lay.drec;
    float* dws_s = smem + lay.dws;
    float* mta_s = smem + lay.mta;

    const int u0 = part_lo(cta, Ha, ncta), u1 = part_lo(cta + 1, Ha, ncta), U = u1 - u0;
    const int p0 = part_lo(cta, BL, ncta), p1 = part_lo(cta + 1, BL, ncta), np = p1 - p0;
    const int d0 = part_lo(cta, A, ncta), d1 = part_lo(cta + 1, A, ncta), nd = d1 - d0;

    for (int idx = threadIdx.x; idx < kUMax * H4; idx += NT) {
        const int ul = idx / H4, r = idx % H4;
        WT[idx] = ul < U ? __ldg(p.whh + (size_t)r * Ha + u0 + ul) : 0.f;
    }
    if (mwp_res)
        for (int idx = threadIdx.x; idx < np * (H4 >> 2); idx += NT)
            reinterpret_cast<float4*>(MWp)[idx] = __ldg(reinterpret_cast<const float4*>(p.mw_pm + (size_t)p0 * H4) + idx);
    const float* mwp_src = mwp_res ? MWp : p.mw_pm + (size_t)p0 * H4;
    for (int idx = threadIdx.x; idx < U * A; idx += NT) {
        const int ul = idx / A, d = idx % A;
        wqT[idx] = __ldg(p.wq + (size_t)d * Ha + u0 + ul);
    }
    for (int idx = threadIdx.x; idx < A * F; idx += NT) {
        const int d = idx / F, f = idx % F;
        wldT[f * AP + d] = __ldg(p.wld + idx);
    }
    for (int idx = threadIdx.x; idx < F * 2 * Kl; idx += NT) {
        const int f = idx / (2 * Kl), ck = idx % (2 * Kl);
        wloc_s[f * CKP + ck] = __ldg(p.wloc + idx);
    }
    for (int idx = threadIdx.x; idx < A; idx += NT) vs[idx] = __ldg(p.v + idx);
    for (int idx = threadIdx.x; idx < kUMax * BP; idx += NT) dhs[idx] = 0.f;
    for (int idx = threadIdx.x; idx < np; idx += NT) { gcum_s[idx] = 0.f; dprev_s[idx] = 0.f; pout[idx] = 0.f; }
    if (kFa) {
        for (int idx = threadIdx.x; idx < BL; idx += NT) {
            mta_s[idx] = p.ta ? __ldg(p.mta + idx) : 0.f;
            drec_s[idx] = 0.f;
        }
        for (int idx = threadIdx.x; idx < 2 * kBMax; idx += NT) (&dzu_s[0][0])[idx] = 0.f;
        if (cta == 0 && (int)threadIdx.x < B) p.dzu[(size_t)(T - 1) * B + threadIdx.x] = 0.f;      // u(T-1) is never used
    }

    // point-wise role of the LSTM backward: thread (ul, b); forward stash + external gradient fetched one step ahead
    const bool pw = (int)threadIdx.x < U * B;
    const int ul = pw ? threadIdx.x / B : 0, pb = pw ? threadIdx.x % B : 0, u = u0 + ul;
    const float wta_u = (kFa && p.ta && pw) ? __ldg(p.wta_h + u) : 0.f;      // query half of the transition agent, owned unit
    float gi[4] = {0.f, 0.f, 0.f, 0.f}, cc = 0.f, cp = 0.f, dhe = 0.f, dcarry = 0.f;
    unsigned char mk = 1;
    // streaming forward stash of the attention part, fetched one step ahead as well:
    //   sv_own : s[t][owned position][d] for item = threadIdx.x (< np*A), da_own: da_ext[t][owned position]
    //   sv_q   : s[t][(b,l) = threadIdx.x][d0] for the first owned attention dim
    float sv_own = 0.f, sv_q = 0.f, da_own = 0.f;
    auto fetch = [&](int t) {
        if (pw) {
            const size_t zb = ((size_t)t * B + pb) * H4, hb = ((size_t)t * B + pb) * Ha + u;
#pragma unroll
            for (int g = 0; g < 4; ++g) gi[g] = __ldg(p.ga + zb + (size_t)g * Ha + u);
            cc = __ldg(p.ca + hb);
            cp = t > 0 ? __ldg(p.ca + hb - (size_t)B * Ha) : 0.f;
            dhe = __ldg(p.dha_ext + hb);
            if (p.mask) mk = p.mask[hb];
        }
        if ((int)threadIdx.x < np * A) sv_own = __ldg(p.s + ((size_t)t * BL + p0) * A + threadIdx.x);
        if (nd > 0 && (int)threadIdx.x < BL) sv_q = __ldg(p.s + ((size_t)t * BL + threadIdx.x) * A + d0);
        if ((int)threadIdx.x < np) da_own = __ldg(p.da_ext + (size_t)t * BL + p0 + threadIdx.x);
    };
    fetch(T - 1);
    SpinGuard sg(p.abort_word);
    ChainProf<kProf> prof;
    prof.start(p.prof, p.trace, p.trace_t0);
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        // ---- P1: context-path recurrent term of the owned positions from dz_a(t+1); publishes dat(t) ----
        if (t < T - 1 && (p.flags & kFlagGate)) {
            const float* zrow = p.dza + ((size_t)(t + 1) * B + (B - 1)) * H4 + (size_t)3 * Ha;
            gate_wait(ncta, [&](int c) { const int e = part_lo(c + 1, Ha, ncta); return e > 0 ? zrow + e - 1 : nullptr; }, sg);
            __syncthreads();
        }
        if (t < T - 1) cta_pair_dots<NT>(mwp_src, (size_t)H4, p.dza + (size_t)(t + 1) * B * H4, H4, p0, np, L, pout, red, sg);
        prof.mark(0, T - 1 - t);
        for (int i = threadIdx.x; i < np; i += NT) {
            const float dae = i == (int)threadIdx.x ? da_own : __ldg(p.da_ext + (size_t)t * BL + p0 + i);
            if (!kFa) {
                st_pub(p.dat + (size_t)t * DS + p0 + i, dae + pout[i] + gcum_s[i] + dprev_s[i]);
            } else {      // d alpha(t) (context, MW and "previous alignment" paths) and d a(t) (cumulative path) separately
                st_pub(p.dat + (size_t)t * DS + p0 + i, dae + pout[i] + dprev_s[i]);
                st_pub(p.dat + (size_t)t * DS + BL + p0 + i, gcum_s[i]);
            }
        }
        // forward stash of this step that the next phase needs (plain data of an earlier kernel)
        for (int idx = threadIdx.x; idx < BL; idx += NT) als[idx] = __ldg(p.align + (size_t)t * BL + idx);
        if ((int)threadIdx.x < B) zn_s[threadIdx.x] = __ldg(p.znorm + (size_t)t * B + threadIdx.x);
        if (kFa) {
            for (int idx = threadIdx.x; idx < BL; idx += NT) {
                apl_s[idx] = __ldg(p.aplain + (size_t)t * BL + idx);
                const int l = idx % L;
                apr_s[idx] = t > 0 ? __ldg(p.align + (size_t)(t - 1) * BL + idx) : (l == 0 ? 1.f : 1e-7f);
            }
            if ((int)threadIdx.x < B) {
                us_s[threadIdx.x] = __ldg(p.ustash + (size_t)t * B + threadIdx.x);
                fs_s[threadIdx.x] = __ldg(p.fsum + (size_t)t * B + threadIdx.x);
            }
        }
        // in the shadow of the d a(t) hand-off: first half of W_hh^T . dz_a(t+1) for the owned units (needed by the point-wise phase).
        // A hand-off that is waited on directly costs ~2.5 us from the last publish to the release (measured,
        // profiles/r01_trace_attn_chain_bwd_v7.txt; independent of polling back-off), one that is given >= 2 us of independent
        // work costs nothing, so the mat-vec is split over the two hand-offs that have nothing else to hide behind.
        if (t < T - 1) cta_matvec_bwd<NT>(WT, H4, p.dza + (size_t)(t + 1) * B * H4, B, part, dhs, BP, sg, 0, 2);
        prof.mark(1, T - 1 - t);
        // ---- hand-off 1: d a(t) of every position ----
        gather_words<NT>(das, p.dat + (size_t)t * DS, DS, (p.flags & kFlagWarp0) != 0, sg);
        __syncthreads();
        prof.mark(2, T - 1 - t);
        // ---- P2: normalisation backward, dS, d(conv features), dq; publishes dconvf(t), dq(t) ----
        for (int b = w; b < B; b += NW) {
            const float* arow = kFa ? apl_s + b * L : als + b * L;      // normalise(e(t))
            if (kFa) {
                // forward-attention backward (forward_attn.py:154-176): alpha = alpha'/S, alpha' = w * a,
                // w = (1-u) alpha(t-1) + u shift(alpha(t-1)) + 1e-8, u = u(t-1) = sigmoid(zu(t-1))
                const float uu = us_s[b], S = fs_s[b];
                float sd = 0.f;
                for (int l = lane; l < L; l += 32) sd += (das[b * L + l] + drec_s[b * L + l]) * als[b * L + l];
                sd = warp_sum(sd);
                float du = 0.f;
                for (int l = lane; l < L; l += 32) {
                    const float dap = ((das[b * L + l] + drec_s[b * L + l]) - sd) / S;
                    const float al = apr_s[b * L + l], sh = l > 0 ? apr_s[b * L + l - 1] : 0.f;
                    const float wl = ((1.f - uu) * al + uu * sh) + 1e-8f;
                    das[b * L + l] = dap * wl + das[BL + b * L + l];          // d a(t): recursion path + cumulative path
                    const float dw = dap * arow[l];
                    dws_s[b * L + l] = dw;
                    du += dw * (sh - al);
                }
                du = warp_sum(du);
                __syncwarp();
                const float dzu = (p.ta && t > 0) ? du * uu * (1.f - uu) : 0.f;   // u(-1) = 0.5 is a constant
                for (int l = lane; l < L; l += 32)      // d alpha(t-1): recursion + context half of the transition agent u(t-1)
                    drec_s[b * L + l] = (1.f - uu) * dws_s[b * L + l] + uu * (l + 1 < L ? dws_s[b * L + l + 1] : 0.f) + dzu * mta_s[b * L + l];
                if (lane == 0 && t > 0) {
                    dzu_s[(t - 1) & 1][b] = dzu;
                    if (cta == 0) p.dzu[(size_t)(t - 1) * B + b] = dzu;
                }
                __syncwarp();
            }
            float sd = 0.f;
            for (int l = lane; l < L; l += 32) sd += arow[l] * das[b * L + l];
            sd = warp_sum(sd);
            for (int l = lane; l < L; l += 32) {
                const float a = arow[l];
                float de = a * (das[b * L + l] - sd);                       // softmax backward
                if (p.norm == 1) de = de * (1.f - a * zn_s[b]);             // sigmoid/sum backward: s = a*Z
                des[b * L + l] = de;
            }
        }
        __syncthreads();
        // dq first: it is the next hand-off, everything else of this phase is local work
        if (nd > 0)       // terms of dq for the first owned attention dim (prefetched column of s)
            for (int idx = threadIdx.x; idx < BL; idx += NT) {
                const float sv = idx == (int)threadIdx.x ? sv_q : __ldg(p.s + ((size_t)t * BL + idx) * A + d0);
                tq_s[idx] = des[idx] * (1.f - sv * sv);
            }
        for (int i = threadIdx.x; i < np; i += NT) p.de[(size_t)t * BL + p0 + i] = des[p0 + i];
        __syncthreads();
        for (int it = NW - 1 - w; it < nd * B; it += NW) {
            const int di = it / B, b = it - di * B, d = d0 + di;
            float acc = 0.f;
            if (di == 0) {
                for (int l = lane; l < L; l += 32) acc += tq_s[b * L + l];
            } else {
                for (int l = lane; l < L; l += 32) {
                    const float sv = __ldg(p.s + ((size_t)t * BL + b * L + l) * A + d);
                    acc += des[b * L + l] * (1.f - sv * sv);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) st_pub(p.dq + ((size_t)t * B + b) * A + d, vs[d] * acc);
        }
        for (int it = threadIdx.x; it < np * A; it += NT) {
            const int pi = it / A, d = it - pi * A;
            const float sv = it == (int)threadIdx.x ? sv_own : __ldg(p.s + ((size_t)t * BL + p0) * A + it);
            const float dS = des[p0 + pi] * vs[d] * (1.f - sv * sv);
            p.ds[((size_t)t * BL + p0) * A + it] = dS;
            ds_s[it] = dS;
        }
        __syncthreads();
        // d(conv features) of the owned positions: item (pi, f, js) sums d = js, js+8, ...; 8 consecutive lanes share (pi, f)
        for (int base = 0; base < np * F * 8; base += NT) {     // warp-uniform trip count (full-mask shuffles inside)
            const int it = base + threadIdx.x;
            const bool valid = it < np * F * 8;
            const int js = it & 7, f = valid ? (it >> 3) % F : 0, pi = valid ? (it >> 3) / F : 0;
            float acc = 0.f;
            if (valid)
                for (int d = js; d < A; d += 8) acc += ds_s[pi * A + d] * wldT[f * AP + d];
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (valid && js == 0) st_pub(p.dconvf + ((size_t)t * BL + p0 + pi) * F + f, acc);
        }
        prof.mark(3, T - 1 - t);
        // in the shadow of the dq hand-off: second half of W_hh^T . dz_a(t+1)
        if (t < T - 1) cta_matvec_bwd<NT>(WT, H4, p.dza + (size_t)(t + 1) * B * H4, B, part, dhs, BP, sg, 1, 2);
        prof.mark(7, T - 1 - t);
        // ---- hand-off 2: dq(t) ----
        gather_words<NT>(dq_s, p.dq + (size_t)t * B * A, B * A, (p.flags & kFlagWarp0) != 0, sg);
        __syncthreads();
        prof.mark(4, T - 1 - t);
        // ---- P3: LSTM point-wise backward (units); publishes dz_a(t) ----
        // q-path term of dh: qd[cell] = sum_d Wq[d][u] dq[b][d], 16 lanes per cell
        for (int base = 0; base < U * B * 16; base += NT) {     // warp-uniform trip count (full-mask shuffles inside)
            const int it = base + threadIdx.x;
            const bool valid = it < U * B * 16;
            const int js = it & 15, cell = valid ? it >> 4 : 0, cu = cell / B, cb = cell - cu * B;
            float acc = 0.f;
            if (valid)
                for (int d = js; d < A; d += 16) acc += wqT[cu * A + d] * dq_s[cb * A + d];
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            if (valid && js == 0) qd_s[cell] = acc;
        }
        __syncthreads();
        prof.mark(8, T - 1 - t);
        if (pw) {
            const size_t zb = ((size_t)t * B + pb) * H4;
            float dh = dhe + (t < T - 1 ? dhs[ul * BP + pb] : 0.f) + qd_s[threadIdx.x];
            if (kFa) dh += dzu_s[t & 1][pb] * wta_u;      // query half of the transition agent u(t)
            if (p.mask) dh = mk ? dh * p.drop_scale : 0.f;
            const LstmGrad gr = lstm_point_bwd(gi[0], gi[1], gi[2], gi[3], cc, cp, dh, dcarry);
            dcarry = gr.dc_prev;
            st_pub(p.dza + zb + 0 * Ha + u, gr.di);
            st_pub(p.dza + zb + 1 * Ha + u, gr.df);
            st_pub(p.dza + zb + 2 * Ha + u, gr.dg);
            st_pub(p.dza + zb + 3 * Ha + u, gr.do_);
        }
        if (t > 0) fetch(t - 1);
        prof.mark(5, T - 1 - t);
        // ---- hand-off 3 (dconvf of the +-pad neighbours) + location-conv backward for the owned positions ----
        // item (pi, k): one warp, lane = filter; output position lo = l - k + pad is fed by input l through tap k
        for (int base = w; base < np * Kl; base += 4 * NW) {      // 4 items per warp in flight: the polls overlap
            float dv[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int it = base + j * NW, pi = it / Kl, k = it - pi * Kl;
                const int pp = p0 + pi, b = pp / L, l = pp - b * L, lo = l - k + pl;
                ok[j] = it < np * Kl && lo >= 0 && lo < L && lane < F;
                dv[j] = ok[j] ? ld_poll(p.dconvf + ((size_t)t * BL + b * L + lo) * F + lane) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int it = base + j * NW, pi = it / Kl, k = it - pi * Kl;
                const int pp = p0 + pi, b = pp / L, l = pp - b * L, lo = l - k + pl;
                float a0 = 0.f, a1 = 0.f;
                if (ok[j]) {
                    sg.reset();
                    while (is_canary(dv[j])) {
                        if (sg.bail()) break;
                        dv[j] = ld_poll(p.dconvf + ((size_t)t * BL + b * L + lo) * F + lane);
                    }
                    a0 = wloc_s[lane * CKP + k] * dv[j];
                    a1 = wloc_s[lane * CKP + Kl + k] * dv[j];
                }
                a0 = warp_sum(a0);
                a1 = warp_sum(a1);
                if (lane == 0 && it < np * Kl) {
                    cpart[it * 2 + 0] = a0;
                    cpart[it * 2 + 1] = a1;
                }
            }
        }
        __syncthreads();
        prof.mark(9, T - 1 - t);
        for (int i = w; i < np; i += NW) {            // one warp per owned position, lanes over the taps (fixed order)
            float a0 = 0.f, a1 = 0.f;
            for (int k = lane; k < Kl; k += 32) {
                a0 += cpart[(i * Kl + k) * 2 + 0];
                a1 += cpart[(i * Kl + k) * 2 + 1];
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
            if (lane == 0) {
                dprev_s[i] = a0;          // d/d a(t-1) through the "previous alignment" channel
                gcum_s[i] += a1;          // running d/d cum(t-1)
            }
        }
        __syncthreads();
        prof.mark(6, T - 1 - t);
    }
}

size_t attn_chain_fwd_smem(int B, int L, int Ha, int A, int F, int Kl, int sm_count, bool mw_resident, bool fa) {
    return attn_fwd_layout(B, L, Ha, A, F, Kl, sm_count, mw_resident, fa).total * sizeof(float);
}
size_t attn_chain_bwd_smem(int B, int L, int Ha, int A, int F, int Kl, int sm_count, bool mwp_resident, bool fa) {
    return attn_bwd_layout(B, L, Ha, A, F, Kl, sm_count, mwp_resident, fa).total * sizeof(float);
}

static int attn_check(int B, int Ha, int A, int F, int sm_count) {
    MSA_CHECK(Ha % 4 == 0, MSA_E_UNSUPPORTED, "attn_chain: attention_rnn_dim %d must be a multiple of 4", Ha);
    MSA_CHECK(F <= 32, MSA_E_UNSUPPORTED, "attn_chain: attention_location_n_filters %d > 32", F);
    MSA_CHECK(B >= 1 && B <= kBMax, MSA_E_UNSUPPORTED, "attn_chain: batch %d outside [1,%d]", B, kBMax);
    MSA_CHECK((Ha + sm_count - 1) / sm_count <= kUMax, MSA_E_UNSUPPORTED,
              "attn_chain: attention_rnn_dim %d needs more than %d units per CTA on %d CTAs", Ha, kUMax, sm_count);
    MSA_CHECK(4 * ((Ha + sm_count - 1) / sm_count) * B <= kAttnThreads, MSA_E_UNSUPPORTED,
              "attn_chain: %d units x %d batch rows per CTA exceed the point-wise thread budget", (Ha + sm_count - 1) / sm_count, B);
    (void)A;
    return 0;
}

// can the single-task kernels run this shape at all? (pass.cu: the grouped tensor-core kernels take over when they cannot)
bool attn_chain_single_ok(int B, int L, int Ha, int A, int F, int Kl, int sm_count, size_t smem_limit, bool fa) {
    if (Ha % 4 != 0 || F > 32 || B < 1 || B > kBMax) return false;
    const int umax = (Ha + sm_count - 1) / sm_count;
    if (umax > kUMax || 4 * umax * B > kAttnThreads) return false;
    return attn_chain_fwd_smem(B, L, Ha, A, F, Kl, sm_count, false, fa) <= smem_limit &&
           attn_chain_bwd_smem(B, L, Ha, A, F, Kl, sm_count, false, fa) <= smem_limit;
}

int launch_attn_chain_fwd(const AttnChainParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(attn_check(p.B, p.Ha, p.A, p.F, sm_count));
    int mw_res = 1;
    size_t smem = attn_chain_fwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, true, p.fa != 0);
    if (smem > smem_limit) {
        mw_res = 0;
        smem = attn_chain_fwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, false, p.fa != 0);
    }
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "attn_chain_fwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    void* fn = p.fa ? (p.prof ? (void*)k_attn_chain_fwd<true, kAttnThreads, true> : (void*)k_attn_chain_fwd<false, kAttnThreads, true>)
                    : (p.prof ? (void*)k_attn_chain_fwd<true, kAttnThreads, false> : (void*)k_attn_chain_fwd<false, kAttnThreads, false>);
    MSA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // canaries of the three hand-off arrays (common.cuh)
    const size_t TB = (size_t)p.T * p.B;
    MSA_TRY(k_fill_canary(p.ha, (int64_t)(TB * p.Ha), st));
    MSA_TRY(k_fill_canary(p.q, (int64_t)(TB * p.A), st));
    MSA_TRY(k_fill_canary(p.e, (int64_t)(TB * p.L), st));
    AttnChainParams pp = p;
    void* args[] = {&pp, &mw_res};
    MSA_CUDA(cudaLaunchCooperativeKernel(fn, dim3(sm_count), dim3(kAttnThreads), args, smem, st));
    count_launch();
    return 0;
}

int launch_attn_chain_bwd(const AttnChainBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(attn_check(p.B, p.Ha, p.A, p.F, sm_count));
    int mwp_res = 1;
    size_t smem = attn_chain_bwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, true, p.fa != 0);
    if (smem > smem_limit) {
        mwp_res = 0;
        smem = attn_chain_bwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, false, p.fa != 0);
    }
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "attn_chain_bwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    void* fn = p.fa ? (p.prof ? (void*)k_attn_chain_bwd<true, kRecThreads, true> : (void*)k_attn_chain_bwd<false, kRecThreads, true>)
                    : (p.prof ? (void*)k_attn_chain_bwd<true, kRecThreads, false> : (void*)k_attn_chain_bwd<false, kRecThreads, false>);
    MSA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t TB = (size_t)p.T * p.B;
    MSA_TRY(k_fill_canary(p.dza, (int64_t)(TB * 4 * p.Ha), st));
    MSA_TRY(k_fill_canary(p.dq, (int64_t)(TB * p.A), st));
    MSA_TRY(k_fill_canary(p.dat, (int64_t)(TB * p.L * (p.fa ? 2 : 1)), st));
    MSA_TRY(k_fill_canary(p.dconvf, (int64_t)(TB * p.L * p.F), st));
    AttnChainBwdParams pp = p;
    void* args[] = {&pp, &mwp_res};
    MSA_CUDA(cudaLaunchCooperativeKernel(fn, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

}  // namespace msa
