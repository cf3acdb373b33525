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
// One cooperative launch, one CTA per SM.  CTA roles (every CTA plays all three):
//   unit owner : hidden units [u0,u1) of the attention LSTM, W_hh rows resident in smem (fp32)
//   pair owner : (b,l) positions [p0,p1): location conv + dense + energy for those positions
//   query owner: attention dims [d0,d1): q[b][d] = Wq[d,:].h_a'[b,:]
// Three grid barriers per step (h_a all-gather, q all-gather, e all-gather).
#include "rec_common.cuh"
#include "kernels.h"

namespace msa {

struct AttnSmemFwd {
    size_t wsm, mws, hs, as_, cums, part, zs, cs, wqs, wlocT, wldT, vs, pre, cf, total;
};
__host__ __device__ inline AttnSmemFwd attn_fwd_layout(int B, int L, int Ha, int A, int F, int Kl, int ncta, bool mw_res) {
    AttnSmemFwd s;
    const int BP = (B + 3) & ~3;
    const int np_max = (B * L + ncta - 1) / ncta, nd_max = (A + ncta - 1) / ncta;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.wsm = take((size_t)4 * kUMax * Ha);
    s.mws = take(mw_res ? (size_t)4 * kUMax * B * L : 0);
    s.hs = take((size_t)BP * Ha);
    s.as_ = take((size_t)B * L);
    s.cums = take((size_t)B * L);
    s.part = take(kRecWarps * 32);
    s.zs = take(32 * BP);
    s.cs = take(kUMax * BP);
    s.wqs = take((size_t)nd_max * Ha);
    s.wlocT = take((size_t)2 * Kl * F);
    s.wldT = take((size_t)F * A);
    s.vs = take(A);
    s.pre = take((size_t)np_max * A);
    s.cf = take(kRecWarps * 32);
    s.total = o;
    return s;
}

// normalise energies (row per warp): softmax or sigmoid/sum.  e_src may be global (ld_cg) .
__device__ __forceinline__ void normalise_rows(const float* e_src, float* a_dst, float* zn_dst, int B, int L, int norm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = w; b < B; b += kRecWarps) {
        if (norm == 0) {
            float m = -INFINITY;
            for (int l = lane; l < L; l += 32) m = fmaxf(m, ld_cg(e_src + b * L + l));
            m = warp_max(m);
            float sum = 0.f;
            for (int l = lane; l < L; l += 32) {
                const float x = expf(ld_cg(e_src + b * L + l) - m);
                a_dst[b * L + l] = x;
                sum += x;
            }
            sum = warp_sum(sum);
            for (int l = lane; l < L; l += 32) a_dst[b * L + l] = a_dst[b * L + l] / sum;
            if (lane == 0 && zn_dst) zn_dst[b] = sum;
        } else {
            float sum = 0.f;
            for (int l = lane; l < L; l += 32) {
                const float x = sigmoidf_(ld_cg(e_src + b * L + l));
                a_dst[b * L + l] = x;
                sum += x;
            }
            sum = warp_sum(sum);
            for (int l = lane; l < L; l += 32) a_dst[b * L + l] = a_dst[b * L + l] / sum;
            if (lane == 0 && zn_dst) zn_dst[b] = sum;
        }
    }
}

__global__ void __launch_bounds__(kRecThreads, 1) k_attn_chain_fwd(AttnChainParams p, int mw_res) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float zn_s[16];
    const int T = p.T, B = p.B, L = p.L, Ha = p.Ha, A = p.A, F = p.F, Kl = p.Kl, H4 = 4 * Ha;
    const int BP = (B + 3) & ~3, BL = B * L, pl = (Kl - 1) / 2;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const AttnSmemFwd lay = attn_fwd_layout(B, L, Ha, A, F, Kl, ncta, mw_res != 0);
    float* Wsm = smem + lay.wsm;
    float* MWs = smem + lay.mws;
    float* hs = smem + lay.hs;
    float* as_ = smem + lay.as_;
    float* cums = smem + lay.cums;
    float* part = smem + lay.part;
    float* zs = smem + lay.zs;
    float* cs = smem + lay.cs;
    float* wqs = smem + lay.wqs;
    float* wlocT = smem + lay.wlocT;
    float* wldT = smem + lay.wldT;
    float* vs = smem + lay.vs;
    float* pre_s = smem + lay.pre;
    float* cf_s = smem + lay.cf;

    const int u0 = part_lo(cta, Ha, ncta), u1 = part_lo(cta + 1, Ha, ncta), U = u1 - u0, R = 4 * U;
    const int p0 = part_lo(cta, BL, ncta), p1 = part_lo(cta + 1, BL, ncta), np = p1 - p0;
    const int d0 = part_lo(cta, A, ncta), d1 = part_lo(cta + 1, A, ncta), nd = d1 - d0;

    // ---- one-time staging of the resident operands ----
    for (int idx = threadIdx.x; idx < R * (Ha >> 2); idx += kRecThreads) {
        const int rl = idx / (Ha >> 2), k4 = idx % (Ha >> 2), g = rl & 3, ul = rl >> 2;
        reinterpret_cast<float4*>(Wsm)[(size_t)rl * (Ha >> 2) + k4] =
            __ldg(reinterpret_cast<const float4*>(p.whh + (size_t)(g * Ha + u0 + ul) * Ha) + k4);
    }
    if (mw_res) {
        for (int idx = threadIdx.x; idx < R * BL; idx += kRecThreads) {
            const int rl = idx / BL, j = idx % BL, g = rl & 3, ul = rl >> 2;
            MWs[idx] = __ldg(p.mw_rm + (size_t)(g * Ha + u0 + ul) * BL + j);
        }
    }
    for (int idx = threadIdx.x; idx < nd * Ha; idx += kRecThreads) wqs[idx] = __ldg(p.wq + (size_t)d0 * Ha + idx);
    for (int idx = threadIdx.x; idx < F * 2 * Kl; idx += kRecThreads) {
        const int f = idx / (2 * Kl), ck = idx % (2 * Kl);          // wloc[f][c][k] -> wlocT[c*Kl+k][f]
        wlocT[ck * F + f] = __ldg(p.wloc + idx);
    }
    for (int idx = threadIdx.x; idx < A * F; idx += kRecThreads) {
        const int d = idx / F, f = idx % F;                          // wld[d][f] -> wldT[f][d]
        wldT[f * A + d] = __ldg(p.wld + idx);
    }
    for (int idx = threadIdx.x; idx < A; idx += kRecThreads) vs[idx] = __ldg(p.v + idx);
    for (int idx = threadIdx.x; idx < kUMax * BP; idx += kRecThreads) cs[idx] = 0.f;
    for (int idx = threadIdx.x; idx < BL; idx += kRecThreads) { as_[idx] = 0.f; cums[idx] = 0.f; }
    const float bv = __ldg(p.bv);
    GridBarrier gb;
    gb.init(p.barrier);
    __syncthreads();

    for (int t = 0; t <= T; ++t) {
        // ---- phase 0: a(t-1) = normalise(e(t-1)); cum(t-1) += a(t-1) (forward_attn.py:200-210) ----
        if (t > 0) {
            normalise_rows(p.ebuf, as_, zn_s, B, L, p.norm);
            __syncthreads();
            for (int idx = threadIdx.x; idx < BL; idx += kRecThreads) {
                const float a = as_[idx];
                cums[idx] += a;
                if (cta == 0) p.align[(size_t)(t - 1) * BL + idx] = a;
            }
            if (cta == 0 && (int)threadIdx.x < B) p.znorm[(size_t)(t - 1) * B + threadIdx.x] = zn_s[threadIdx.x];
        }
        if (t == T) break;
        if (cta == 0)
            for (int idx = threadIdx.x; idx < BL; idx += kRecThreads) p.cum[(size_t)t * BL + idx] = cums[idx];
        __syncthreads();

        // ---- phase 1a: location features for the owned (b,l) positions (forward_attn.py:121-127) ----
        for (int pi = w; pi < np; pi += kRecWarps) {
            const int pp = p0 + pi, b = pp / L, l = pp % L;
            float cf = 0.f;
            if (lane < F) {
                for (int k = 0; k < Kl; ++k) {
                    const int ll = l + k - pl;
                    if (ll >= 0 && ll < L)
                        cf += wlocT[(0 * Kl + k) * F + lane] * as_[b * L + ll] + wlocT[(1 * Kl + k) * F + lane] * cums[b * L + ll];
                }
                p.convf[((size_t)t * BL + pp) * F + lane] = cf;
            }
            cf_s[w * 32 + lane] = cf;
            __syncwarp();
            for (int d = lane; d < A; d += 32) {
                float loc = 0.f;
                for (int f = 0; f < F; ++f) loc += wldT[f * A + d] * cf_s[w * 32 + f];
                pre_s[pi * A + d] = loc + __ldg(p.pm + (size_t)pp * A + d);
            }
            __syncwarp();
        }

        // ---- phase 1b: attention LSTMCell for the owned units (decoder.py:253-256) ----
        if (U > 0) {
            for (int idx = threadIdx.x; idx < B * (Ha >> 2); idx += kRecThreads) {
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t > 0) v4 = ld_cg4(p.ha + (size_t)(t - 1) * B * Ha + (size_t)idx * 4);
                reinterpret_cast<float4*>(hs)[idx] = v4;
            }
            __syncthreads();
            if (mw_res)
                cta_matvec_fwd<true>(Wsm, R, Ha, hs, MWs, BL, 0, 0, L, as_, B, part, zs, BP);
            else
                cta_matvec_fwd<true>(Wsm, R, Ha, hs, p.mw_rm, BL, Ha, u0, L, as_, B, part, zs, BP);
            if ((int)threadIdx.x < U * B) {
                const int ul = threadIdx.x / B, b = threadIdx.x % B, u = u0 + ul;
                const size_t zb = ((size_t)t * B + b) * H4;
                LstmPoint r = lstm_point_fwd(zs[(ul * 4 + 0) * BP + b] + __ldg(p.xw + zb + 0 * Ha + u),
                                             zs[(ul * 4 + 1) * BP + b] + __ldg(p.xw + zb + 1 * Ha + u),
                                             zs[(ul * 4 + 2) * BP + b] + __ldg(p.xw + zb + 2 * Ha + u),
                                             zs[(ul * 4 + 3) * BP + b] + __ldg(p.xw + zb + 3 * Ha + u), cs[ul * BP + b]);
                cs[ul * BP + b] = r.c;
                float hv = r.h;
                if (p.mask) hv = p.mask[((size_t)t * B + b) * Ha + u] ? hv * p.drop_scale : 0.f;
                p.ga[zb + 0 * Ha + u] = r.i;
                p.ga[zb + 1 * Ha + u] = r.f;
                p.ga[zb + 2 * Ha + u] = r.g;
                p.ga[zb + 3 * Ha + u] = r.o;
                p.ca[((size_t)t * B + b) * Ha + u] = r.c;
                p.ha[((size_t)t * B + b) * Ha + u] = hv;
            }
        }
        gb.sync();

        // ---- phase 2: query projection for the owned attention dims (forward_attn.py:125) ----
        for (int di = w; di < nd; di += kRecWarps) {
            for (int bt = 0; bt < B; bt += 4) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = lane * 4; k < Ha; k += 128) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wqs + (size_t)di * Ha + k);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (bt + b < B) {
                            const float4 h4 = ld_cg4(p.ha + ((size_t)t * B + bt + b) * Ha + k);
                            acc[b] += w4.x * h4.x + w4.y * h4.y + w4.z * h4.z + w4.w * h4.w;
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float sacc = warp_sum(acc[b]);
                    if (lane == 0 && bt + b < B) p.q[((size_t)t * B + bt + b) * A + d0 + di] = sacc;
                }
            }
        }
        gb.sync();

        // ---- phase 3: energies for the owned positions (forward_attn.py:128-131) ----
        for (int pi = w; pi < np; pi += kRecWarps) {
            const int pp = p0 + pi, b = pp / L;
            float e = 0.f;
            for (int d = lane; d < A; d += 32) {
                const float sv = tanhf(ld_cg(p.q + ((size_t)t * B + b) * A + d) + pre_s[pi * A + d]);
                p.s[((size_t)t * BL + pp) * A + d] = sv;
                e += vs[d] * sv;
            }
            e = warp_sum(e);
            if (lane == 0) p.ebuf[pp] = e + bv;
        }
        gb.sync();
    }
}

// =====================================================================================
// backward
struct AttnSmemBwd {
    size_t wt, part, dhs, dcs, wqT, das, als, des, wld, wlocT, vs, dss, gcum, dprev, pout, dqs, total;
};
__host__ __device__ inline AttnSmemBwd attn_bwd_layout(int B, int L, int Ha, int A, int F, int Kl, int ncta) {
    AttnSmemBwd s;
    const int BP = (B + 3) & ~3;
    const int np_max = (B * L + ncta - 1) / ncta;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.wt = take((size_t)kUMax * 4 * Ha);
    s.part = take(kRecWarps * 32);
    s.dhs = take(kUMax * BP);
    s.dcs = take(kUMax * BP);
    s.wqT = take((size_t)kUMax * A);
    s.das = take((size_t)B * L);
    s.als = take((size_t)B * L);
    s.des = take((size_t)B * L);
    s.wld = take((size_t)A * F);
    s.wlocT = take((size_t)2 * Kl * F);
    s.vs = take(A);
    s.dss = take((size_t)kRecWarps * A);
    s.gcum = take(np_max);
    s.dprev = take(np_max);
    s.pout = take(np_max);
    s.dqs = take((size_t)B * A);
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(kRecThreads, 1) k_attn_chain_bwd(AttnChainBwdParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[33];
    __shared__ float zn_s[16];
    const int T = p.T, B = p.B, L = p.L, Ha = p.Ha, A = p.A, F = p.F, Kl = p.Kl, H4 = 4 * Ha;
    const int BP = (B + 3) & ~3, BL = B * L, pl = (Kl - 1) / 2;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const AttnSmemBwd lay = attn_bwd_layout(B, L, Ha, A, F, Kl, ncta);
    float* WT = smem + lay.wt;
    float* part = smem + lay.part;
    float* dhs = smem + lay.dhs;
    float* dcs = smem + lay.dcs;
    float* wqT = smem + lay.wqT;
    float* das = smem + lay.das;
    float* als = smem + lay.als;
    float* des = smem + lay.des;
    float* wld_s = smem + lay.wld;
    float* wlocT = smem + lay.wlocT;
    float* vs = smem + lay.vs;
    float* ds_s = smem + lay.dss;
    float* gcum_s = smem + lay.gcum;
    float* dprev_s = smem + lay.dprev;
    float* pout = smem + lay.pout;
    float* dq_s = smem + lay.dqs;

    const int u0 = part_lo(cta, Ha, ncta), u1 = part_lo(cta + 1, Ha, ncta), U = u1 - u0;
    const int p0 = part_lo(cta, BL, ncta), p1 = part_lo(cta + 1, BL, ncta), np = p1 - p0;
    const int d0 = part_lo(cta, A, ncta), d1 = part_lo(cta + 1, A, ncta), nd = d1 - d0;

    for (int idx = threadIdx.x; idx < U * H4; idx += kRecThreads) {
        const int ul = idx / H4, r = idx % H4;
        WT[idx] = __ldg(p.whh + (size_t)r * Ha + u0 + ul);
    }
    for (int idx = threadIdx.x; idx < U * A; idx += kRecThreads) {
        const int ul = idx / A, d = idx % A;
        wqT[idx] = __ldg(p.wq + (size_t)d * Ha + u0 + ul);
    }
    for (int idx = threadIdx.x; idx < A * F; idx += kRecThreads) wld_s[idx] = __ldg(p.wld + idx);
    for (int idx = threadIdx.x; idx < F * 2 * Kl; idx += kRecThreads) {
        const int f = idx / (2 * Kl), ck = idx % (2 * Kl);
        wlocT[ck * F + f] = __ldg(p.wloc + idx);
    }
    for (int idx = threadIdx.x; idx < A; idx += kRecThreads) vs[idx] = __ldg(p.v + idx);
    for (int idx = threadIdx.x; idx < kUMax * BP; idx += kRecThreads) { dcs[idx] = 0.f; dhs[idx] = 0.f; }
    for (int idx = threadIdx.x; idx < np; idx += kRecThreads) { gcum_s[idx] = 0.f; dprev_s[idx] = 0.f; pout[idx] = 0.f; }
    GridBarrier gb;
    gb.init(p.barrier);
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        // ---- P1: recurrent terms from dz_a(t+1): W_hh^T.dz (units) and (Wc.memory[l]).dz (positions) ----
        if (t < T - 1)
            cta_matvec_bwd(WT, U, H4, p.dza + (size_t)(t + 1) * B * H4, B, part, dhs, BP, p.mw_pm, p0, np, L, pout, red);
        for (int i = threadIdx.x; i < np; i += kRecThreads)
            p.dat[p0 + i] = __ldg(p.da_ext + (size_t)t * BL + p0 + i) + pout[i] + gcum_s[i] + dprev_s[i];
        gb.sync();

        // ---- P2: normalisation backward, dS, d(conv features), dq ----
        for (int idx = threadIdx.x; idx < BL; idx += kRecThreads) {
            das[idx] = ld_cg(p.dat + idx);
            als[idx] = __ldg(p.align + (size_t)t * BL + idx);
        }
        if ((int)threadIdx.x < B) zn_s[threadIdx.x] = __ldg(p.znorm + (size_t)t * B + threadIdx.x);
        __syncthreads();
        for (int b = w; b < B; b += kRecWarps) {
            float sd = 0.f;
            for (int l = lane; l < L; l += 32) sd += als[b * L + l] * das[b * L + l];
            sd = warp_sum(sd);
            for (int l = lane; l < L; l += 32) {
                const float a = als[b * L + l];
                float de = a * (das[b * L + l] - sd);                       // softmax backward
                if (p.norm == 1) de = de * (1.f - a * zn_s[b]);             // sigmoid/sum backward: s = a*Z
                des[b * L + l] = de;
                if (cta == 0) p.de[(size_t)t * BL + b * L + l] = de;
            }
        }
        __syncthreads();
        for (int pi = w; pi < np; pi += kRecWarps) {
            const int pp = p0 + pi;
            const float de = des[pp];
            for (int d = lane; d < A; d += 32) {
                const float sv = __ldg(p.s + ((size_t)t * BL + pp) * A + d);
                const float dS = de * vs[d] * (1.f - sv * sv);
                p.ds[((size_t)t * BL + pp) * A + d] = dS;
                ds_s[w * A + d] = dS;
            }
            __syncwarp();
            if (lane < F) {
                float acc = 0.f;
                for (int d = 0; d < A; ++d) acc += ds_s[w * A + d] * wld_s[d * F + lane];
                p.dconvf[((size_t)t * BL + pp) * F + lane] = acc;
            }
            __syncwarp();
        }
        for (int di = w; di < nd; di += kRecWarps) {
            const int d = d0 + di;
            for (int b = 0; b < B; ++b) {
                float acc = 0.f;
                for (int l = lane; l < L; l += 32) {
                    const float sv = __ldg(p.s + ((size_t)t * BL + b * L + l) * A + d);
                    acc += des[b * L + l] * (1.f - sv * sv);
                }
                acc = warp_sum(acc);
                if (lane == 0) p.dq[((size_t)t * B + b) * A + d] = vs[d] * acc;
            }
        }
        gb.sync();

        // ---- P3: LSTM point-wise backward (units) and location-conv backward (positions) ----
        for (int idx = threadIdx.x; idx < B * A; idx += kRecThreads) dq_s[idx] = ld_cg(p.dq + (size_t)t * B * A + idx);
        __syncthreads();
        if ((int)threadIdx.x < U * B) {
            const int ul = threadIdx.x / B, b = threadIdx.x % B, u = u0 + ul;
            const size_t zb = ((size_t)t * B + b) * H4;
            float dh = __ldg(p.dha_ext + ((size_t)t * B + b) * Ha + u) + (t < T - 1 ? dhs[ul * BP + b] : 0.f);
            float qd = 0.f;
            for (int d = 0; d < A; ++d) qd += wqT[ul * A + d] * dq_s[b * A + d];
            dh += qd;
            if (p.mask) dh = p.mask[((size_t)t * B + b) * Ha + u] ? dh * p.drop_scale : 0.f;
            const float cprev = t > 0 ? __ldg(p.ca + ((size_t)(t - 1) * B + b) * Ha + u) : 0.f;
            LstmGrad g = lstm_point_bwd(__ldg(p.ga + zb + 0 * Ha + u), __ldg(p.ga + zb + 1 * Ha + u),
                                        __ldg(p.ga + zb + 2 * Ha + u), __ldg(p.ga + zb + 3 * Ha + u),
                                        __ldg(p.ca + ((size_t)t * B + b) * Ha + u), cprev, dh, dcs[ul * BP + b]);
            dcs[ul * BP + b] = g.dc_prev;
            p.dza[zb + 0 * Ha + u] = g.di;
            p.dza[zb + 1 * Ha + u] = g.df;
            p.dza[zb + 2 * Ha + u] = g.dg;
            p.dza[zb + 3 * Ha + u] = g.do_;
        }
        for (int pi = w; pi < np; pi += kRecWarps) {
            const int pp = p0 + pi, b = pp / L, l = pp % L;
            float a0 = 0.f, a1 = 0.f;
            if (lane < F) {
                for (int k = 0; k < Kl; ++k) {
                    const int lo = l - k + pl;                               // output position fed by input l through tap k
                    if (lo >= 0 && lo < L) {
                        const float dv = ld_cg(p.dconvf + ((size_t)t * BL + b * L + lo) * F + lane);
                        a0 += wlocT[(0 * Kl + k) * F + lane] * dv;
                        a1 += wlocT[(1 * Kl + k) * F + lane] * dv;
                    }
                }
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
            if (lane == 0) {
                dprev_s[pi] = a0;          // d/d a(t-1) through the "previous alignment" channel
                gcum_s[pi] += a1;          // running d/d cum(t-1)
            }
        }
        if (t > 0) gb.sync();
    }
}

size_t attn_chain_fwd_smem(int B, int L, int Ha, int A, int F, int Kl, int sm_count, bool mw_resident) {
    return attn_fwd_layout(B, L, Ha, A, F, Kl, sm_count, mw_resident).total * sizeof(float);
}
size_t attn_chain_bwd_smem(int B, int L, int Ha, int A, int F, int Kl, int sm_count) {
    return attn_bwd_layout(B, L, Ha, A, F, Kl, sm_count).total * sizeof(float);
}

static int attn_check(int B, int Ha, int A, int F, int sm_count) {
    MSA_CHECK(Ha % 4 == 0, MSA_E_UNSUPPORTED, "attn_chain: attention_rnn_dim %d must be a multiple of 4", Ha);
    MSA_CHECK(F <= 32, MSA_E_UNSUPPORTED, "attn_chain: attention_location_n_filters %d > 32", F);
    MSA_CHECK(B >= 1 && B <= 16, MSA_E_UNSUPPORTED, "attn_chain: batch %d outside [1,16]", B);
    MSA_CHECK((Ha + sm_count - 1) / sm_count <= kUMax, MSA_E_UNSUPPORTED,
              "attn_chain: attention_rnn_dim %d needs more than %d units per CTA on %d CTAs", Ha, kUMax, sm_count);
    (void)A;
    return 0;
}

int launch_attn_chain_fwd(const AttnChainParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(attn_check(p.B, p.Ha, p.A, p.F, sm_count));
    int mw_res = 1;
    size_t smem = attn_chain_fwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, true);
    if (smem > smem_limit) {
        mw_res = 0;
        smem = attn_chain_fwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, false);
    }
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "attn_chain_fwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    MSA_CUDA(cudaFuncSetAttribute(k_attn_chain_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSA_CUDA(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), st));
    AttnChainParams pp = p;
    void* args[] = {&pp, &mw_res};
    MSA_CUDA(cudaLaunchCooperativeKernel((void*)k_attn_chain_fwd, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

int launch_attn_chain_bwd(const AttnChainBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(attn_check(p.B, p.Ha, p.A, p.F, sm_count));
    const size_t smem = attn_chain_bwd_smem(p.B, p.L, p.Ha, p.A, p.F, p.Kl, sm_count);
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "attn_chain_bwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    MSA_CUDA(cudaFuncSetAttribute(k_attn_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSA_CUDA(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), st));
    AttnChainBwdParams pp = p;
    void* args[] = {&pp};
    MSA_CUDA(cudaLaunchCooperativeKernel((void*)k_attn_chain_bwd, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

}  // namespace msa
