// Grouped persistent recurrences with tensor-core gate products.
//
// The single-task kernels (lstm_rec.cu, attn_chain.cu) hand h(t) / dz(t) between the 148 co-resident CTAs once per time step
// for the B rows of ONE pass; at B = 4 a step is almost pure hand-off latency.  All tasks of a meta-batch take their first inner
// step from the same weights theta_0 (maml.py:38-54, reptile.py:38-56), so their train-split passes can share the resident
// weight slices AND the hand-offs: the kernels here run one recurrence for G tasks x B rows = R <= 32 rows per launch.
// With R >= 8 rows the per-step gate product of a CTA -- [<= 32 gate rows x H] . [H x R] forward,
// [R x 4H] . [4H x <= 8 units] backward (decoder.py:253-265 and its autograd transpose) -- is a real tile for the tensor cores:
// mma.sync.m16n8k16 on bf16 operands, fp32 accumulation, every fp32 operand split into bf16 hi + bf16 lo and the product
// evaluated as hi.hi + hi.lo + lo.hi (relative error ~1e-5: the weights and activations keep 16 mantissa bits; the dropped
// lo.lo term is 2^-18).  Per CTA the weight slice is converted once into ready-made A (forward) / B (backward) fragments that stay
// in shared memory for all T steps; the other operand arrives from the other CTAs through L2 and goes straight from the
// canary-polled loads (common.cuh) into fragment registers -- it is never staged in shared memory.
//
// Row indexing of a group: row r = g * B + b is batch row b of task g; every float array of task g is the array of task 0
// shifted by g * tstride floats (the tasks' workspaces are identical slices of one allocation, pass.cu).
#include <cuda_bf16.h>

#include "rec_common.cuh"
#include "kernels.h"

namespace msa {

namespace {

#ifndef MSA_PF_FWD
#define MSA_PF_FWD 2
#endif
#ifndef MSA_PF_BWD
#define MSA_PF_BWD 2
#endif
constexpr int kMT = 512;              // threads per CTA
constexpr int kMW = kMT / 32;         // warps = K slices of the gate product

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                         const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// fp32 pair -> packed bf16 hi pair and packed bf16 lo pair (x = hi + lo up to 2^-17 relative)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// D += A.B with both operands given as (hi, lo) bf16 pairs: hi.hi + hi.lo + lo.hi, small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const uint4& ahi, const uint4& alo, const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
    mma_bf16(d, alo.x, alo.y, alo.z, alo.w, bhi[0], bhi[1]);
    mma_bf16(d, ahi.x, ahi.y, ahi.z, ahi.w, blo[0], blo[1]);
    mma_bf16(d, ahi.x, ahi.y, ahi.z, ahi.w, bhi[0], bhi[1]);
}

// Fragment column permutation: within a k16 block, lane j (= lane & 3) owns the four PHYSICAL columns 4j .. 4j+3 and feeds them
// as the logical k indices {2j, 2j+1, 2j+8, 2j+9} -- one 128-bit load per row and k16 block on the streaming side; the resident
// side is packed with the same mapping once.

struct Grp {
    int G, Bt, R;
    int64_t tstride;
    __device__ __forceinline__ int task(int r) const { return r / Bt; }
    __device__ __forceinline__ int brow(int r) const { return r - (r / Bt) * Bt; }
};

// =====================================================================================================================
// LSTM recurrence, forward.  CTA c of a direction owns hidden units [u0, u1) (<= 8): local gate row rl = gate * 8 + ul
// (two m16 tiles: {i, f} and {g, o}), so that after the product one thread holds all four gates of a cell.
// smem: A fragments [warp][k16 step][m tile][hi|lo][lane] uint4, partial tiles [warp][n tile][m tile][reg][lane].
// =====================================================================================================================
template <int NT>
__global__ void __launch_bounds__(kMT, 1) k_lstm_fwd_mma(LstmRecParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = p.H, T = p.T, H4 = 4 * H;
    const Grp gr{p.G, p.B, p.G * p.B, p.tstride};
    const int R = gr.R;
    const int ncta_dir = gridDim.x / p.ndir;
    const int dir = blockIdx.x / ncta_dir;
    if (dir >= p.ndir) return;
    const int ci = blockIdx.x % ncta_dir;
    const int u0 = part_lo(ci, H, ncta_dir), u1 = part_lo(ci + 1, H, ncta_dir), U = u1 - u0;
    if (U == 0) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lg = lane >> 2, lj = lane & 3;
    const int KS = (H + 255) / 256;                       // k16 steps per warp; warp w owns columns [w*KS*16, (w+1)*KS*16)
    uint4* Afrag = reinterpret_cast<uint4*>(smem_raw);   // [kMW][KS][2][2][32]
    float* part = reinterpret_cast<float*>(Afrag + (size_t)kMW * KS * 2 * 2 * 32);      // [kMW][NT][2][4][32]

    const float* whh = p.whh + (size_t)dir * p.whh_dir_stride;
    const int64_t dir_off_h = (int64_t)dir * T * gr.Bt * H, dir_off_z = (int64_t)dir * T * gr.Bt * H4;

    // ---- one-time: the weight slice as A fragments ----
    for (int idx = threadIdx.x; idx < kMW * KS * 2 * 32; idx += kMT) {
        const int ln = idx & 31, mt = (idx >> 5) & 1, s = (idx >> 6) % KS, ww = (idx >> 6) / KS;
        const int g = ln >> 2, j = ln & 3;
        const int col = (ww * KS + s) * 16 + 4 * j;
        float v[2][4];
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int rl = mt * 16 + hr * 8 + g;          // local row: gate = rl >> 3, unit = rl & 7
            const int gate = rl >> 3, ul = rl & 7;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                v[hr][c] = (ul < U && col + c < H) ? __ldg(whh + (size_t)(gate * H + u0 + ul) * H + col + c) : 0.f;
        }
        uint4 hi, lo;
        split2(v[0][0], v[0][1], hi.x, lo.x);     // a0: row g,   logical k 2j, 2j+1
        split2(v[1][0], v[1][1], hi.y, lo.y);     // a1: row g+8
        split2(v[0][2], v[0][3], hi.z, lo.z);     // a2: row g,   logical k 2j+8, 2j+9
        split2(v[1][2], v[1][3], hi.w, lo.w);     // a3: row g+8
        const size_t o = (((size_t)ww * KS + s) * 2 + mt) * 2 * 32;
        Afrag[o + ln] = hi;
        Afrag[o + 32 + ln] = lo;
    }

    // ---- point-wise role: thread c < NT*64 owns cell (unit ul, row r) for the whole launch ----
    const int c_nt = threadIdx.x >> 6, c_par = (threadIdx.x >> 5) & 1, c_ln = threadIdx.x & 31;
    const int c_ul = c_ln >> 2, c_r = c_nt * 8 + (c_ln & 3) * 2 + c_par;
    const bool pw = (int)threadIdx.x < NT * 64 && c_ul < U && c_r < R;
    const int c_g = pw ? gr.task(c_r) : 0, c_b = pw ? gr.brow(c_r) : 0, c_u = u0 + c_ul;
    const float* zin = p.zin + c_g * gr.tstride + dir_off_z;
    float* hout_c = p.hout + c_g * gr.tstride + dir_off_h;
    float* cout_c = p.cout + c_g * gr.tstride + dir_off_h;
    float* gates_c = p.gates + c_g * gr.tstride + dir_off_z;
    const uint8_t* mask_c = pw ? (p.G > 1 ? p.mask_g[c_g] : p.mask) : nullptr;
    const int64_t* len_p = pw ? (p.G > 1 ? p.lengths_g[c_g] : p.lengths) : nullptr;
    const int len = len_p ? (int)len_p[c_b] : T;
    float zi[4] = {0.f, 0.f, 0.f, 0.f}, cstate = 0.f;
    unsigned char mk = 1;
    auto fetch = [&](int t) {
        const size_t zb = ((size_t)t * gr.Bt + c_b) * H4;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) zi[gate] = __ldg(zin + zb + (size_t)gate * H + c_u);
        if (mask_c) mk = mask_c[((size_t)t * gr.Bt + c_b) * H + c_u];
    };
    if (pw) fetch(dir == 0 ? 0 : T - 1);

    // ---- streaming role: lane (lg, lj) of warp w loads h[row nt*8 + lg][its 4 columns] per k16 step ----
    const float* hrow[NT];
    bool rowok[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int r = nt * 8 + lg;
        rowok[nt] = r < R;
        const int rr = rowok[nt] ? r : 0;
        hrow[nt] = p.hout + gr.task(rr) * gr.tstride + dir_off_h + (size_t)gr.brow(rr) * H;
    }
    SpinGuard sg(p.abort_word);
    __syncthreads();

    for (int s_ = 0; s_ < T; ++s_) {
        const int t = (dir == 0) ? s_ : T - 1 - s_;
        const int tp = (dir == 0) ? t - 1 : t + 1;
        float acc[NT][2][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[nt][mt][i] = 0.f;
        if (s_ > 0) {
            // sentinel of producer CTA c: h of its last unit for the last row (a hint only: every word is canary-checked below)
            {
                const int rl_ = R - 1;
                const float* hlast = p.hout + gr.task(rl_) * gr.tstride + dir_off_h + ((size_t)tp * gr.Bt + gr.brow(rl_)) * H;
                gate_wait(ncta_dir, [&](int c) { const int e = part_lo(c + 1, H, ncta_dir); return e > part_lo(c, H, ncta_dir) ? hlast + e - 1 : nullptr; }, sg);
                __syncthreads();
            }
            const size_t toff = (size_t)tp * gr.Bt * H;
            // PF k16 steps in flight per lane (a ring of register buffers with compile-time slot indices; the loads of step
            // s + PF are issued as soon as slot s has been consumed)
            constexpr int PF = MSA_PF_FWD;      // measured on B200: 2 in flight beats 4 (803 vs 736 us at R = 8): later loads find their data published
            float4 hv[PF][NT];
            auto issue = [&](int s, float4 (&dst)[NT]) {
                const int col = (w * KS + s) * 16 + 4 * lj;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    dst[nt] = (rowok[nt] && col < H) ? ld_poll4(hrow[nt] + toff + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto process = [&](int s, float4 (&cur)[NT]) {
                const int col = (w * KS + s) * 16 + 4 * lj;
                uint32_t bhi[NT][2], blo[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (rowok[nt] && col < H) {
                        sg.reset();
                        while (!ready4(cur[nt])) {
                            if (sg.bail()) break;
                            cur[nt] = ld_poll4(hrow[nt] + toff + col);
                        }
                    }
                    split2(cur[nt].x, cur[nt].y, bhi[nt][0], blo[nt][0]);
                    split2(cur[nt].z, cur[nt].w, bhi[nt][1], blo[nt][1]);
                }
                const uint4* af = Afrag + ((size_t)w * KS + s) * 2 * 2 * 32;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint4 ahi = af[(mt * 2 + 0) * 32 + lane], alo = af[(mt * 2 + 1) * 32 + lane];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma3(acc[nt][mt], ahi, alo, bhi[nt], blo[nt]);
                }
            };
#pragma unroll
            for (int i = 0; i < PF; ++i)
                if (i < KS) issue(i, hv[i]);
            for (int s0 = 0; s0 < KS; s0 += PF) {
#pragma unroll
                for (int i = 0; i < PF; ++i) {
                    const int s = s0 + i;
                    if (s < KS) {
                        process(s, hv[i]);
                        if (s + PF < KS) issue(s + PF, hv[i]);
                    }
                }
            }
        }
        // partial tiles of this warp's K slice -> shared
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) part[((((size_t)w * NT + nt) * 2 + mt) * 4 + i) * 32 + lane] = acc[nt][mt][i];
        __syncthreads();
        if (pw) {
            // gate (i, f, g, o) of cell (c_ul, c_r): m tile = gate >> 1, reg = (gate & 1) * 2 + parity, lane c_ln
            float z[4];
#pragma unroll
            for (int gate = 0; gate < 4; ++gate) {
                float sum = 0.f;
                const float* pp = part + ((((size_t)c_nt) * 2 + (gate >> 1)) * 4 + (gate & 1) * 2 + c_par) * 32 + c_ln;
#pragma unroll
                for (int ww = 0; ww < kMW; ++ww) sum += pp[(size_t)ww * NT * 2 * 4 * 32];
                z[gate] = sum + zi[gate];
            }
            const bool active = t < len;
            float ai = fast_sigmoid(z[0]), af = fast_sigmoid(z[1]), ag = fast_tanh(z[2]), ao = fast_sigmoid(z[3]);
            float cn = 0.f, hvv = 0.f;
            if (active) {
                cn = af * cstate + ai * ag;
                cstate = cn;
                hvv = ao * fast_tanh(cn);
                if (mask_c) hvv = mk ? hvv * p.drop_scale : 0.f;
            } else {
                ai = af = ag = ao = 0.f;
            }
            const size_t hb = ((size_t)t * gr.Bt + c_b) * H + c_u, zb = ((size_t)t * gr.Bt + c_b) * H4;
            st_pub(hout_c + hb, hvv);              // consumed by every CTA at the next step: first memory operation
            gates_c[zb + 0 * (size_t)H + c_u] = ai;
            gates_c[zb + 1 * (size_t)H + c_u] = af;
            gates_c[zb + 2 * (size_t)H + c_u] = ag;
            gates_c[zb + 3 * (size_t)H + c_u] = ao;
            cout_c[hb] = cn;
            if (s_ + 1 < T) fetch(dir == 0 ? t + 1 : t - 1);
        }
        // (no barrier here: `part` is rewritten only after the barrier that follows the next step's gate_wait)
    }
}

// =====================================================================================================================
// LSTM recurrence, backward.  dh_rec[r][u] = sum_j dz(t+1)[r][j] W_hh[j][u] for the owned units: M = rows (m16 tiles),
// N = 8 units, K = 4H gate rows split over the warps; the transposed weight slice is resident as B fragments.
// =====================================================================================================================
template <int MTL>
__global__ void __launch_bounds__(kMT, 1) k_lstm_bwd_mma(LstmRecBwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = p.H, T = p.T, H4 = 4 * H;
    const Grp gr{p.G, p.B, p.G * p.B, p.tstride};
    const int R = gr.R;
    const int ncta_dir = gridDim.x / p.ndir;
    const int dir = blockIdx.x / ncta_dir;
    if (dir >= p.ndir) return;
    const int ci = blockIdx.x % ncta_dir;
    const int u0 = part_lo(ci, H, ncta_dir), u1 = part_lo(ci + 1, H, ncta_dir), U = u1 - u0;
    if (U == 0) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lg = lane >> 2, lj = lane & 3;
    const int KS = (H4 + 255) / 256;                      // k16 steps per warp over the 4H gate rows
    uint2* Bfrag = reinterpret_cast<uint2*>(smem_raw);   // [kMW][KS][hi|lo][32]
    float* part = reinterpret_cast<float*>(Bfrag + (size_t)kMW * KS * 2 * 32);           // [kMW][MTL][4][32]

    const float* whh = p.whh + (size_t)dir * p.whh_dir_stride;
    const int64_t dir_off_h = (int64_t)dir * T * gr.Bt * H, dir_off_z = (int64_t)dir * T * gr.Bt * H4;

    for (int idx = threadIdx.x; idx < kMW * KS * 32; idx += kMT) {
        const int ln = idx & 31, s = (idx >> 5) % KS, ww = (idx >> 5) / KS;
        const int g = ln >> 2, j = ln & 3;                // n = unit g; k rows kb + 4j .. 4j+3
        const int kb = (ww * KS + s) * 16 + 4 * j;
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = (g < U && kb + c < H4) ? __ldg(whh + (size_t)(kb + c) * H + u0 + g) : 0.f;
        uint2 hi, lo;
        split2(v[0], v[1], hi.x, lo.x);
        split2(v[2], v[3], hi.y, lo.y);
        const size_t o = ((size_t)ww * KS + s) * 2 * 32;
        Bfrag[o + ln] = hi;
        Bfrag[o + 32 + ln] = lo;
    }

    // point-wise role: thread c < MTL*128 owns cell (row r, unit ul)
    const int c_mt = threadIdx.x >> 7, c_reg = (threadIdx.x >> 5) & 3, c_ln = threadIdx.x & 31;
    const int c_r = c_mt * 16 + (c_reg >> 1) * 8 + (c_ln >> 2), c_ul = (c_ln & 3) * 2 + (c_reg & 1);
    const bool pw = (int)threadIdx.x < MTL * 128 && c_ul < U && c_r < R;
    const int c_g = pw ? gr.task(c_r) : 0, c_b = pw ? gr.brow(c_r) : 0, c_u = u0 + c_ul;
    const float* gates_c = p.gates + c_g * gr.tstride + dir_off_z;
    const float* cst_c = p.cout + c_g * gr.tstride + dir_off_h;
    const float* dhe_c = p.dh_ext + c_g * gr.tstride + dir_off_h;
    float* dz_c = p.dz + c_g * gr.tstride + dir_off_z;
    const uint8_t* mask_c = pw ? (p.G > 1 ? p.mask_g[c_g] : p.mask) : nullptr;
    const int64_t* len_p = pw ? (p.G > 1 ? p.lengths_g[c_g] : p.lengths) : nullptr;
    const int len = len_p ? (int)len_p[c_b] : T;
    float gi[4] = {0.f, 0.f, 0.f, 0.f}, cc = 0.f, cp = 0.f, dhe = 0.f, dcarry = 0.f;
    unsigned char mk = 1;
    auto fetch = [&](int t) {
        const int tp = (dir == 0) ? t - 1 : t + 1;
        const size_t zb = ((size_t)t * gr.Bt + c_b) * H4, hb = ((size_t)t * gr.Bt + c_b) * H + c_u;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) gi[gate] = __ldg(gates_c + zb + (size_t)gate * H + c_u);
        cc = __ldg(cst_c + hb);
        cp = (tp >= 0 && tp < T) ? __ldg(cst_c + ((size_t)tp * gr.Bt + c_b) * H + c_u) : 0.f;
        dhe = __ldg(dhe_c + hb);
        if (mask_c) mk = mask_c[hb];
    };
    if (pw) fetch(dir == 0 ? T - 1 : 0);

    // streaming role: lane (lg, lj) loads dz[rows mt*16 + lg, +8][its 4 gate rows] per k16 step
    const float* zrow[MTL][2];
    bool rowok[MTL][2];
#pragma unroll
    for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int r = mt * 16 + hr * 8 + lg;
            rowok[mt][hr] = r < R;
            const int rr = rowok[mt][hr] ? r : 0;
            zrow[mt][hr] = p.dz + gr.task(rr) * gr.tstride + dir_off_z + (size_t)gr.brow(rr) * H4;
        }
    SpinGuard sg(p.abort_word);
    __syncthreads();

    for (int s_ = 0; s_ < T; ++s_) {
        const int t = (dir == 0) ? T - 1 - s_ : s_;      // reverse of the forward processing order
        const int tn = (dir == 0) ? t + 1 : t - 1;       // step processed just before
        float acc[MTL][4];
#pragma unroll
        for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
        if (s_ > 0) {
            {
                const int rl_ = R - 1;
                const float* zlast = p.dz + gr.task(rl_) * gr.tstride + dir_off_z + ((size_t)tn * gr.Bt + gr.brow(rl_)) * H4 + (size_t)3 * H;
                gate_wait(ncta_dir, [&](int c) { const int e = part_lo(c + 1, H, ncta_dir); return e > part_lo(c, H, ncta_dir) ? zlast + e - 1 : nullptr; }, sg);
                __syncthreads();
            }
            const size_t toff = (size_t)tn * gr.Bt * H4;
            constexpr int PF = MSA_PF_BWD;      // measured on B200: 2 in flight beats 4 / 8 (2639 vs 3395 us at R = 32)
            float4 zv[PF][MTL][2];
            auto issue = [&](int s, float4 (&dst)[MTL][2]) {
                const int col = (w * KS + s) * 16 + 4 * lj;
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr)
                        dst[mt][hr] = (rowok[mt][hr] && col < H4) ? ld_poll4(zrow[mt][hr] + toff + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto process = [&](int s, float4 (&cur)[MTL][2]) {
                const int col = (w * KS + s) * 16 + 4 * lj;
                const uint2* bf = Bfrag + ((size_t)w * KS + s) * 2 * 32;
                const uint2 bh = bf[lane], bl = bf[32 + lane];
                const uint32_t bhi[2] = {bh.x, bh.y}, blo[2] = {bl.x, bl.y};
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt) {
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr) {
                        if (rowok[mt][hr] && col < H4) {
                            sg.reset();
                            while (!ready4(cur[mt][hr])) {
                                if (sg.bail()) break;
                                cur[mt][hr] = ld_poll4(zrow[mt][hr] + toff + col);
                            }
                        }
                    }
                    uint4 ahi, alo;
                    split2(cur[mt][0].x, cur[mt][0].y, ahi.x, alo.x);     // a0: row g,   logical k 2j, 2j+1
                    split2(cur[mt][1].x, cur[mt][1].y, ahi.y, alo.y);     // a1: row g+8
                    split2(cur[mt][0].z, cur[mt][0].w, ahi.z, alo.z);     // a2: row g,   logical k 2j+8, 2j+9
                    split2(cur[mt][1].z, cur[mt][1].w, ahi.w, alo.w);     // a3: row g+8
                    mma3(acc[mt], ahi, alo, bhi, blo);
                }
            };
#pragma unroll
            for (int i = 0; i < PF; ++i)
                if (i < KS) issue(i, zv[i]);
            for (int s0 = 0; s0 < KS; s0 += PF) {
#pragma unroll
                for (int i = 0; i < PF; ++i) {
                    const int s = s0 + i;
                    if (s < KS) {
                        process(s, zv[i]);
                        if (s + PF < KS) issue(s + PF, zv[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) part[(((size_t)w * MTL + mt) * 4 + i) * 32 + lane] = acc[mt][i];
        __syncthreads();
        if (pw) {
            float drec = 0.f;
            const float* pp = part + ((size_t)c_mt * 4 + c_reg) * 32 + c_ln;
#pragma unroll
            for (int ww = 0; ww < kMW; ++ww) drec += pp[(size_t)ww * MTL * 4 * 32];
            const size_t zb = ((size_t)t * gr.Bt + c_b) * H4;
            const bool active = t < len;
            LstmGrad g = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (active) {
                float dh = dhe + drec;
                if (mask_c) dh = mk ? dh * p.drop_scale : 0.f;
                g = lstm_point_bwd(gi[0], gi[1], gi[2], gi[3], cc, cp, dh, dcarry);
                dcarry = g.dc_prev;
            }
            st_pub(dz_c + zb + 0 * (size_t)H + c_u, g.di);
            st_pub(dz_c + zb + 1 * (size_t)H + c_u, g.df);
            st_pub(dz_c + zb + 2 * (size_t)H + c_u, g.dg);
            st_pub(dz_c + zb + 3 * (size_t)H + c_u, g.do_);
            if (s_ + 1 < T) fetch(dir == 0 ? t - 1 : t + 1);
        }
    }
}

size_t lstm_fwd_mma_smem(int H, int NT) {
    const int KS = (H + 255) / 256;
    return (size_t)kMW * KS * 2 * 2 * 32 * sizeof(uint4) + (size_t)kMW * NT * 2 * 4 * 32 * sizeof(float);
}
size_t lstm_bwd_mma_smem(int H, int MTL) {
    const int KS = (4 * H + 255) / 256;
    return (size_t)kMW * KS * 2 * 32 * sizeof(uint2) + (size_t)kMW * MTL * 4 * 32 * sizeof(float);
}

bool lstm_mma_ok(int R, int H, int ndir, int sm_count, size_t smem_limit) {
    if (R < 1 || R > 32 || H % 4 != 0) return false;
    const int ncta_dir = sm_count / ndir;
    if (ncta_dir < 1 || (H + ncta_dir - 1) / ncta_dir > 8) return false;
    return lstm_fwd_mma_smem(H, (R + 7) / 8) <= smem_limit && lstm_bwd_mma_smem(H, (R + 15) / 16) <= smem_limit;
}

template <class Kern, class Params>
int coop_launch(Kern kern, const Params& p, int sm_count, size_t smem, cudaStream_t st) {
    MSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Params pp = p;
    void* args[] = {&pp};
    MSA_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(sm_count), dim3(kMT), args, smem, st));
    count_launch();
    return 0;
}

}  // namespace

int launch_lstm_rec_fwd_mma(const LstmRecParams& p0, int sm_count, size_t smem_limit, cudaStream_t st) {
    LstmRecParams p = p0;
    if (p.G < 1) p.G = 1;
    const int R = p.G * p.B;
    MSA_CHECK(lstm_mma_ok(R, p.H, p.ndir, sm_count, smem_limit), MSA_E_UNSUPPORTED, "lstm_rec_fwd_mma: %d rows x H=%d not supported", R, p.H);
    const int NT = (R + 7) / 8;
    for (int g = 0; g < p.G; ++g)      // canaries of the hand-off array of every task (common.cuh)
        MSA_TRY(k_fill_canary(p.hout + g * p.tstride, (int64_t)((size_t)p.ndir * p.T * p.B * p.H), st));
    const size_t smem = lstm_fwd_mma_smem(p.H, NT);
    switch (NT) {
        case 1: return coop_launch(k_lstm_fwd_mma<1>, p, sm_count, smem, st);
        case 2: return coop_launch(k_lstm_fwd_mma<2>, p, sm_count, smem, st);
        case 3: return coop_launch(k_lstm_fwd_mma<3>, p, sm_count, smem, st);
        default: return coop_launch(k_lstm_fwd_mma<4>, p, sm_count, smem, st);
    }
}

int launch_lstm_rec_bwd_mma(const LstmRecBwdParams& p0, int sm_count, size_t smem_limit, cudaStream_t st) {
    LstmRecBwdParams p = p0;
    if (p.G < 1) p.G = 1;
    const int R = p.G * p.B;
    MSA_CHECK(lstm_mma_ok(R, p.H, p.ndir, sm_count, smem_limit), MSA_E_UNSUPPORTED, "lstm_rec_bwd_mma: %d rows x H=%d not supported", R, p.H);
    const int MTL = (R + 15) / 16;
    for (int g = 0; g < p.G; ++g)
        MSA_TRY(k_fill_canary(p.dz + g * p.tstride, (int64_t)((size_t)p.ndir * p.T * p.B * 4 * p.H), st));
    const size_t smem = lstm_bwd_mma_smem(p.H, MTL);
    if (MTL == 1) return coop_launch(k_lstm_bwd_mma<1>, p, sm_count, smem, st);
    return coop_launch(k_lstm_bwd_mma<2>, p, sm_count, smem, st);
}

// pass.cu asks per chain family: the LSTM recurrences (encoder BiLSTM, decoder RNN) and the attention chain
bool chain_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit) {
    (void)T; (void)L;
    const int R = G * B;
    return lstm_mma_ok(R, cfg.dec_rnn_dim, 1, sm_count, smem_limit) && lstm_mma_ok(R, cfg.enc_dim / 2, 2, sm_count, smem_limit);
}
bool attn_chain_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit) {
    (void)cfg; (void)G; (void)B; (void)T; (void)L; (void)sm_count; (void)smem_limit;
    return false;
}
int launch_attn_chain_fwd_mma(const AttnChainParams&, int, size_t, cudaStream_t) { set_error("attn_chain_fwd_mma: not built"); return MSA_E_UNSUPPORTED; }
int launch_attn_chain_bwd_mma(const AttnChainBwdParams&, int, size_t, cudaStream_t) { set_error("attn_chain_bwd_mma: not built"); return MSA_E_UNSUPPORTED; }

}  // namespace msa
