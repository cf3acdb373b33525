// Device building blocks shared by the persistent recurrent kernels (lstm_rec.cu, attn_chain.cu).
//
// Work decomposition (forward): the 4H gate rows of an LSTMCell are split by HIDDEN UNIT across the
// co-resident CTAs of a cooperative launch; CTA c owns units [u0,u1) and keeps the 4*(u1-u0) rows
// of W_hh resident in shared memory (fp32) for all T steps.  Each step is a skinny mat-vec
// (rows x K) . (K x B) from shared memory, the LSTM point-wise update for the owned units, a
// store of the owned slice of h, and one grid barrier (the all-gather of h).
// Backward: CTA c owns the same units and keeps the matching COLUMNS of W_hh (i.e. rows of W_hh^T)
// resident; dh_rec[u] = sum_r W_hh[r][u] * dz[r] is computed with each thread owning a slice of r.
#pragma once
#include "common.cuh"

namespace msa {

constexpr int kRecThreads = 512;
constexpr int kRecWarps = kRecThreads / 32;
constexpr int kUMax = 8;  // max hidden units per CTA (=> 32 gate rows, one transpose-reduce)

struct LstmPoint {
    float i, f, g, o, c, h;
};
__device__ __forceinline__ LstmPoint lstm_point_fwd(float zi, float zf, float zg, float zo, float cprev) {
    LstmPoint r;
    r.i = sigmoidf_(zi);
    r.f = sigmoidf_(zf);
    r.g = tanhf(zg);
    r.o = sigmoidf_(zo);
    r.c = r.f * cprev + r.i * r.g;
    r.h = r.o * tanhf(r.c);
    return r;
}
// dh: grad w.r.t. the (pre-dropout) hidden output; dc_in: grad carried from the later step.
// Returns gate pre-activation grads and the carry for the earlier step.
struct LstmGrad {
    float di, df, dg, do_, dc_prev;
};
__device__ __forceinline__ LstmGrad lstm_point_bwd(float i, float f, float g, float o, float c, float cprev, float dh,
                                                   float dc_in) {
    LstmGrad r;
    const float tc = tanhf(c);
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

// zs[rl*BP + b] = sum_k Wsm[rl*K + k] * hs[b*K + k]  (+ sum_l MW[row(rl)*mw_stride + b*L + l] * as[b*L + l] if HAS_MW)
// for rl < R (R <= 32... any R <= 8*warps), b < B.  All threads of the CTA must call (contains __syncthreads).
// K % 4 == 0, Wsm/hs 16-byte aligned rows.  `part` scratch: kRecWarps*32 floats.
template <bool HAS_MW>
__device__ __forceinline__ void cta_matvec_fwd(const float* __restrict__ Wsm, int R, int K, const float* __restrict__ hs,
                                               const float* __restrict__ MW, int mw_stride, int mw_gs, int mw_u0, int L,
                                               const float* __restrict__ as, int B, float* part, float* zs, int BP) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int RG = (R + 7) >> 3;
    const int KS = RG > 0 ? kRecWarps / RG : 0;
    const int rg = RG > 0 ? w % RG : 0, ks = RG > 0 ? w / RG : 0;
    const int KCh = (K + 127) >> 7;
    const int KCl = HAS_MW ? ((L + 31) >> 5) : 0;
    for (int bt = 0; bt < B; bt += 4) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        if (ks < KS) {
            for (int j = ks; j < KCh + KCl; j += KS) {
                if (j < KCh) {
                    const int k = (j << 7) + (lane << 2);
                    if (k < K) {
                        float4 h4[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            h4[b] = (bt + b < B) ? *reinterpret_cast<const float4*>(hs + (size_t)(bt + b) * K + k)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const int rl = rg * 8 + rr;
                            if (rl < R) {
                                const float4 w4 = *reinterpret_cast<const float4*>(Wsm + (size_t)rl * K + k);
#pragma unroll
                                for (int b = 0; b < 4; ++b)
                                    acc[rr * 4 + b] += w4.x * h4[b].x + w4.y * h4[b].y + w4.z * h4[b].z + w4.w * h4[b].w;
                            }
                        }
                    }
                } else if (HAS_MW) {
                    const int l = ((j - KCh) << 5) + lane;
                    if (l < L) {
                        float a4[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) a4[b] = (bt + b < B) ? as[(bt + b) * L + l] : 0.f;
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const int rl = rg * 8 + rr;
                            if (rl < R) {
                                // resident slice: local row order; global fallback: gate-major rows of the full matrix
                                const int mrow = mw_gs ? ((rl & 3) * mw_gs + mw_u0 + (rl >> 2)) : rl;
#pragma unroll
                                for (int b = 0; b < 4; ++b)
                                    if (bt + b < B) acc[rr * 4 + b] += MW[(size_t)mrow * mw_stride + (bt + b) * L + l] * a4[b];
                            }
                        }
                    }
                }
            }
        }
        const float tot = warp_transpose_reduce32(acc);
        if (ks < KS) part[(ks * RG + rg) * 32 + lane] = tot;
        __syncthreads();
        if ((int)threadIdx.x < R * 4) {
            const int rl = threadIdx.x >> 2, b = threadIdx.x & 3;
            const int rg2 = rl >> 3, idx = (rl & 7) * 4 + b;
            float s = 0.f;
            for (int q = 0; q < KS; ++q) s += part[(q * RG + rg2) * 32 + idx];
            if (bt + b < B) zs[rl * BP + bt + b] = s;
        }
        __syncthreads();
    }
}

// Backward recurrent term for the owned units:
//   out[ul*BP + b] = sum_r WT[ul*R4 + r] * dz[b*R4 + r],  ul < U (<= 8), b < B, r < R4 (= 4H, multiple of 4)
// dz lives in global memory, written by other CTAs in the previous step (read through L2).
// If npairs > 0 also computes pair_out[i] = sum_r MWp[(pair0+i)*R4 + r] * dz[b_i*R4 + r] with b_i = (pair0+i)/L.
// All threads must call.  `part` scratch: kRecWarps*32 floats, `red` scratch 33 floats.
__device__ __forceinline__ void cta_matvec_bwd(const float* __restrict__ WT, int U, int R4, const float* dz, int B,
                                               float* part, float* out, int BP, const float* __restrict__ MWp,
                                               int pair0, int npairs, int L, float* pair_out, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nchunk = R4 >> 2;
    for (int bt = 0; bt < B; bt += 4) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        for (int ch = threadIdx.x; ch < nchunk; ch += kRecThreads) {
            const int r = ch << 2;
            float4 d4[4];
#pragma unroll
            for (int b = 0; b < 4; ++b)
                d4[b] = (bt + b < B) ? ld_cg4(dz + (size_t)(bt + b) * R4 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int ul = 0; ul < kUMax; ++ul) {
                if (ul < U) {
                    const float4 w4 = *reinterpret_cast<const float4*>(WT + (size_t)ul * R4 + r);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        acc[ul * 4 + b] += w4.x * d4[b].x + w4.y * d4[b].y + w4.z * d4[b].z + w4.w * d4[b].w;
                }
            }
        }
        const float tot = warp_transpose_reduce32(acc);
        part[w * 32 + lane] = tot;
        __syncthreads();
        if ((int)threadIdx.x < U * 4) {
            const int ul = threadIdx.x >> 2, b = threadIdx.x & 3;
            float s = 0.f;
            for (int q = 0; q < kRecWarps; ++q) s += part[q * 32 + ul * 4 + b];
            if (bt + b < B) out[ul * BP + bt + b] = s;
        }
        __syncthreads();
    }
    for (int i = 0; i < npairs; ++i) {
        const int p = pair0 + i, b = p / L;
        float s = 0.f;
        for (int ch = threadIdx.x; ch < nchunk; ch += kRecThreads) {
            const int r = ch << 2;
            const float4 d = ld_cg4(dz + (size_t)b * R4 + r);
            const float4 m = __ldg(reinterpret_cast<const float4*>(MWp + (size_t)p * R4 + r));
            s += m.x * d.x + m.y * d.y + m.z * d.z + m.w * d.w;
        }
        s = block_sum(s, red);
        if (threadIdx.x == 0) pair_out[i] = s;
    }
    __syncthreads();
}

}  // namespace msa
