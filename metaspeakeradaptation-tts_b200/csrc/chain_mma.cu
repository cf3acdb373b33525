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

#ifndef MSA_LSTM_GATE
#define MSA_LSTM_GATE 1
#endif
#ifndef MSA_SENT_LAST
#define MSA_SENT_LAST 0
#endif
#ifndef MSA_SLICE_ROT
#define MSA_SLICE_ROT 1
#endif
#ifndef MSA_PF_FWD
#define MSA_PF_FWD 2
#endif
#ifndef MSA_PF_BWD1
#define MSA_PF_BWD1 2
#endif
#ifndef MSA_PF_BWD2
#define MSA_PF_BWD2 2
#endif
#define MSA_PF_BWD (MTL == 1 ? MSA_PF_BWD1 : MSA_PF_BWD2)
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

// Polling slow paths, deliberately NOT inlined: the fast path of every hand-off load is one canary test; only a load that came back
// too early calls in here.  (Inlined spin loops with their time-out bookkeeping at every poll site had grown the attention kernels to
// 130-170 KB of SASS -- far beyond the instruction caches, every step re-fetched its code from L2.)
__device__ __noinline__ float4 poll4_slow(const float* p, unsigned int* abort_word) {
    float4 v = ld_poll4(p);
    if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) return v;
    unsigned int n = 0;
    while (!ready4(v)) {
        if ((++n & 1023u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) break;
            if (n > (1u << 21)) { *reinterpret_cast<volatile unsigned int*>(abort_word) = 1u; break; }
        }
        v = ld_poll4(p);
    }
    return v;
}
__device__ __noinline__ float poll1_slow(const float* p, unsigned int* abort_word) {
    float v = ld_poll(p);
    if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) return v;
    unsigned int n = 0;
    while (is_canary(v)) {
        if ((++n & 1023u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) break;
            if (n > (1u << 21)) { *reinterpret_cast<volatile unsigned int*>(abort_word) = 1u; break; }
        }
        v = ld_poll(p);
    }
    return v;
}
__device__ __forceinline__ float4 pollq4(const float* p, unsigned int* abort_word) {
    float4 v = ld_poll4(p);
    if (!ready4(v)) v = poll4_slow(p, abort_word);
    return v;
}
__device__ __forceinline__ float pollq1(const float* p, unsigned int* abort_word) {
    float v = ld_poll(p);
    if (is_canary(v)) v = poll1_slow(p, abort_word);
    return v;
}
// first n threads wait for one sentinel word each (rec_common.cuh::gate_wait with the out-of-line poller)
template <class AddrFn>
__device__ __forceinline__ void gate_waitq(int n, AddrFn addr, unsigned int* abort_word) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float* a = addr(i);
        if (a) (void)pollq1(a, abort_word);
    }
}

// Streaming of ready-made weight fragments (per-task-weight variants): 16-byte asynchronous copies into a per-warp ring; every lane
// copies and later reads only its own 16-byte slots, so the copies need no cross-lane synchronisation.
#ifndef MSA_RING_FWD
#define MSA_RING_FWD 4
#endif
constexpr int kRingF = MSA_RING_FWD;  // forward: ring stages of 2 KB per warp (4 stages: 128 KB per SM in flight between product phases)
// backward: a ring stage of a warp holds one k16 step of all G tasks -- G x 512 bytes of weight fragments and G x Bt x 64 bytes of
// dz(t+1) -- and both travel by cp.async: the depth of the stream no longer depends on registers (two polled k16 steps in registers
// were what bound the dz gather); five stages at two tasks, three at three or four (shared memory)
__host__ __device__ constexpr int ring_depth_bwd(int G) { return G <= 2 ? 5 : 3; }
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16_sa(uint32_t smem_addr, const void* g) {      // shared-window address computed by the caller
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
    // K slice of this warp, rotated by the CTA index: the 148 CTAs stream the SAME h / dz rows, without the rotation they would all
    // ask the same L2 slices for the same lines at the same moment
    const int wsl = MSA_SLICE_ROT ? (w + blockIdx.x) % kMW : w;
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
    const bool is_sent = MSA_SENT_LAST && pw && c_ul == U - 1 && c_r == R - 1;
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
            if (MSA_LSTM_GATE) {
                const int rl_ = R - 1;
                const float* hlast = p.hout + gr.task(rl_) * gr.tstride + dir_off_h + ((size_t)tp * gr.Bt + gr.brow(rl_)) * H;
                gate_waitq(ncta_dir, [&](int c) { const int e = part_lo(c + 1, H, ncta_dir); return e > part_lo(c, H, ncta_dir) ? hlast + e - 1 : nullptr; }, p.abort_word);
                __syncthreads();
            }
            const size_t toff = (size_t)tp * gr.Bt * H;
            // PF k16 steps in flight per lane (a ring of register buffers with compile-time slot indices; the loads of step
            // s + PF are issued as soon as slot s has been consumed)
            constexpr int PF = MSA_PF_FWD;      // measured on B200: 2 in flight beats 4 (803 vs 736 us at R = 8): later loads find their data published
            float4 hv[PF][NT];
            auto issue = [&](int s, float4 (&dst)[NT]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    dst[nt] = (rowok[nt] && col < H) ? ld_poll4(hrow[nt] + toff + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto process = [&](int s, float4 (&cur)[NT]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
                uint32_t bhi[NT][2], blo[NT][2];
                unsigned int cmax = 0u;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) cmax = umax_acc(cmax, cur[nt]);
                if (cmax == kCanary) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        if (rowok[nt] && col < H && !ready4(cur[nt])) cur[nt] = poll4_slow(hrow[nt] + toff + col, p.abort_word);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    split2(cur[nt].x, cur[nt].y, bhi[nt][0], blo[nt][0]);
                    split2(cur[nt].z, cur[nt].w, bhi[nt][1], blo[nt][1]);
                }
                const uint4* af = Afrag + ((size_t)wsl * KS + s) * 2 * 2 * 32;
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
        if ((int)threadIdx.x < NT * 64) {
          float hvv = 0.f;
          size_t hb = 0;
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
            float cn = 0.f;
            if (active) {
                cn = af * cstate + ai * ag;
                cstate = cn;
                hvv = ao * fast_tanh(cn);
                if (mask_c) hvv = mk ? hvv * p.drop_scale : 0.f;
            } else {
                ai = af = ag = ao = 0.f;
            }
            hb = ((size_t)t * gr.Bt + c_b) * H + c_u;
            const size_t zb = ((size_t)t * gr.Bt + c_b) * H4;
            if (!is_sent) st_pub(hout_c + hb, hvv);              // consumed by every CTA at the next step: first memory operation
            gates_c[zb + 0 * (size_t)H + c_u] = ai;
            gates_c[zb + 1 * (size_t)H + c_u] = af;
            gates_c[zb + 2 * (size_t)H + c_u] = ag;
            gates_c[zb + 3 * (size_t)H + c_u] = ao;
            cout_c[hb] = cn;
            if (s_ + 1 < T) fetch(dir == 0 ? t + 1 : t - 1);
          }
          // the sentinel word (the one the consumers' gate waits for) leaves after every other cell of this CTA has been published
          if (MSA_SENT_LAST) asm volatile("bar.sync 1, %0;" ::"n"(NT * 64));
          if (is_sent) st_pub(hout_c + hb, hvv);
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
    // K slice of this warp, rotated by the CTA index: the 148 CTAs stream the SAME h / dz rows, without the rotation they would all
    // ask the same L2 slices for the same lines at the same moment
    const int wsl = MSA_SLICE_ROT ? (w + blockIdx.x) % kMW : w;
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
    const bool is_sent = MSA_SENT_LAST && pw && c_ul == U - 1 && c_r == R - 1;
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
                gate_waitq(ncta_dir, [&](int c) { const int e = part_lo(c + 1, H, ncta_dir); return e > part_lo(c, H, ncta_dir) ? zlast + e - 1 : nullptr; }, p.abort_word);
                __syncthreads();
            }
            const size_t toff = (size_t)tn * gr.Bt * H4;
            constexpr int PF = MSA_PF_BWD;      // measured on B200: 2 in flight beats 4 / 8 (2639 vs 3395 us at R = 32)
            float4 zv[PF][MTL][2];
            auto issue = [&](int s, float4 (&dst)[MTL][2]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr)
                        dst[mt][hr] = (rowok[mt][hr] && col < H4) ? ld_poll4(zrow[mt][hr] + toff + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto process = [&](int s, float4 (&cur)[MTL][2]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
                const uint2* bf = Bfrag + ((size_t)wsl * KS + s) * 2 * 32;
                const uint2 bh = bf[lane], bl = bf[32 + lane];
                const uint32_t bhi[2] = {bh.x, bh.y}, blo[2] = {bl.x, bl.y};
                unsigned int cmax = 0u;
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                    for (int hr = 0; hr < 2; ++hr) cmax = umax_acc(cmax, cur[mt][hr]);
                if (cmax == kCanary) {
#pragma unroll
                    for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                        for (int hr = 0; hr < 2; ++hr)
                            if (rowok[mt][hr] && col < H4 && !ready4(cur[mt][hr])) cur[mt][hr] = poll4_slow(zrow[mt][hr] + toff + col, p.abort_word);
                }
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt) {
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
        if ((int)threadIdx.x < MTL * 128) {
          float dzo = 0.f;
          size_t zo = 0;
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
            dzo = g.do_;
            zo = zb + 3 * (size_t)H + c_u;
            if (!is_sent) st_pub(dz_c + zo, dzo);
            if (s_ + 1 < T) fetch(dir == 0 ? t - 1 : t + 1);
          }
          if (MSA_SENT_LAST) asm volatile("bar.sync 1, %0;" ::"n"(MTL * 128));      // the sentinel word leaves last (see the forward kernel)
          if (is_sent) st_pub(dz_c + zo, dzo);
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

// =====================================================================================================================
// Attention chain, forward (same recurrence as attn_chain.cu::k_attn_chain_fwd without the forward-attention options):
//     z_a(t) = XW[t] + MW . a(t-1) + W_hh . h_a'(t-1);  h_a'(t) = dropout(LSTMCell);  q(t) = W_q . h_a'(t)
//     e(t) = v . tanh(q + loc(a(t-1), cum(t-1)) + PM) + b_v;  a(t) = normalise(e(t))            (decoder.py:253-258, forward_attn.py:121-131)
// for R = G*B rows.  CTA roles as there (unit owner / position owner / query owner); three hand-offs per step (h, q, e).
// The W_hh product and the query projection are ONE tensor-core tile: the CTA's gate rows sit in local rows gate*8 + ul
// (ul < U <= 7) and the W_q row of the owned attention dim in the free slot ul = 7 of gate 0, so q(t)[r] falls out of the same
// mma as "gate i of pseudo-unit 7".  The context term MW . a(t-1) streams the CTA's MW rows from L2 (they are 1 MB per batch row
// and cannot be resident for a whole group).
// =====================================================================================================================
// A fragment idx = ((K slice ww * KS + k16 step s) * 2 + m tile mt) * 32 + lane of the CTA that owns units [u0, u0 + U) and (has_q)
// attention dim d0: local row rl = mt * 16 + hr * 8 + g is gate rl >> 3 of unit slot rl & 7; slot 7 of gate 0 is the W_q row
__device__ __forceinline__ void attn_fwd_frag(const float* __restrict__ whh, const float* __restrict__ wq, int Ha, int KS, int u0, int U,
                                              bool has_q, int d0, int idx, uint4& hi, uint4& lo) {
    const int ln = idx & 31, mt = (idx >> 5) & 1, s = (idx >> 6) % KS, ww = (idx >> 6) / KS;
    const int g = ln >> 2, j = ln & 3;
    const int col = (ww * KS + s) * 16 + 4 * j;
    float v[2][4];
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
        const int rl = mt * 16 + hr * 8 + g;
        const int gate = rl >> 3, ul = rl & 7;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float x = 0.f;
            if (col + c < Ha) {
                if (ul < U) x = __ldg(whh + (size_t)(gate * Ha + u0 + ul) * Ha + col + c);
                else if (has_q && ul == 7 && gate == 0) x = __ldg(wq + (size_t)d0 * Ha + col + c);
            }
            v[hr][c] = x;
        }
    }
    split2(v[0][0], v[0][1], hi.x, lo.x);
    split2(v[1][0], v[1][1], hi.y, lo.y);
    split2(v[0][2], v[0][3], hi.z, lo.z);
    split2(v[1][2], v[1][3], hi.w, lo.w);
}
// the same fragments of G tasks written to global memory in the order the PT kernel streams them: [task][cta][ww][s][mt][hi|lo][lane]
struct FragSrc {
    const float* whh[kPtGroupMax];
    const float* wq[kPtGroupMax];
};
__global__ void __launch_bounds__(kMT) k_attn_frag_fwd(FragSrc src, int Ha, int A, int64_t stride, uint4* out) {
    const int task = blockIdx.y, cta = blockIdx.x, ncta = gridDim.x;
    const float* whh = src.whh[task];
    const float* wq = src.wq[task];
    const int KS = (Ha + 255) / 256;
    const int u0 = part_lo(cta, Ha, ncta), U = part_lo(cta + 1, Ha, ncta) - u0;
    const int d0 = part_lo(cta, A, ncta);
    const bool has_q = part_lo(cta + 1, A, ncta) > d0;
    uint4* o = out + (size_t)task * stride + (size_t)cta * kMW * KS * 128;
    for (int idx = threadIdx.x; idx < kMW * KS * 2 * 32; idx += kMT) {
        uint4 hi, lo;
        attn_fwd_frag(whh, wq, Ha, KS, u0, U, has_q, d0, idx, hi, lo);
        const size_t b = (size_t)(idx >> 5) * 2 * 32;
        o[b + (idx & 31)] = hi;
        o[b + 32 + (idx & 31)] = lo;
    }
}

struct AttnFwdMmaLay {
    size_t afrag, part, as_, ah, wldT, wloc, vs, pm, pre, cf, zm, rowoff, posoff, posrl, total;     // byte offsets
    int KS, LP, LH, CKP, NPmax, NRown, RP;
};
__host__ __device__ inline AttnFwdMmaLay attn_fwd_mma_layout(int R, int L, int Ha, int A, int F, int Kl, int ncta, int NT, int pt = 0) {
    AttnFwdMmaLay s;
    s.KS = (Ha + 255) / 256;
    s.LP = round_up_i(L, 4);
    s.LH = round_up_i(L + Kl - 1, 4);
    s.CKP = ((2 * Kl + 7) & ~7) + 8;              // filter stride = 8 banks (mod 32): the 4 filters x 8 taps of a warp hit 32 banks
    if ((s.CKP & 31) != 8) s.CKP = ((s.CKP + 31) & ~31) + 8;
    s.NPmax = (R * L + ncta - 1) / ncta;
    s.NRown = (s.NPmax + L - 2) / L + 1;          // batch rows a run of NPmax consecutive positions can touch
    s.RP = round_up_i(R, 4);
    const int NTP = NT < 2 ? NT : 2;              // partial tiles are reduced two n tiles at a time
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    // resident A fragments, or (per-task weights) the ring of streamed fragments: kRingF stages of one (k16 step, task) chunk per warp
    s.afrag = take(pt ? (size_t)kMW * kRingF * 2 * 2 * 32 * sizeof(uint4) : (size_t)kMW * s.KS * 2 * 2 * 32 * sizeof(uint4));
    const int NW = pt ? 2 : 1;                    // weight sets of the small attention weights: the tasks of the owned positions
    // partial tiles; the gathered energies [R*L] and the gathered queries [NRown][A] reuse the region later in the step
    size_t pbytes = (size_t)kMW * NTP * 2 * 4 * 32 * sizeof(float);
    const size_t gbytes = ((size_t)R * L + (size_t)s.NRown * A) * sizeof(float) + 32;
    if (gbytes > pbytes) pbytes = gbytes;
    s.part = take(pbytes);
    s.as_ = take((size_t)R * s.LP * sizeof(float));
    s.ah = take((size_t)2 * s.NRown * s.LH * sizeof(float));
    s.wldT = take((size_t)NW * F * A * sizeof(float));
    s.wloc = take((size_t)NW * 2 * Kl * F * sizeof(float));      // tap-major [2 * Kl][F]: lanes = filters, conflict-free
    s.vs = take((size_t)NW * (A + 4) * sizeof(float));           // v, then b_v at [A]
    s.pm = take((size_t)s.NPmax * A * sizeof(float));
    s.pre = take((size_t)s.NPmax * A * sizeof(float));
    s.cf = take((size_t)s.NPmax * F * sizeof(float));
    s.zm = take((size_t)32 * s.RP * sizeof(float));
    s.rowoff = take((size_t)64 * sizeof(long long));            // row r -> task(r) * tstride; [32 + r] -> ... + brow(r) * L (floats)
    s.posoff = take((size_t)s.NPmax * sizeof(long long));      // owned position -> task * tstride (floats)
    s.posrl = take((size_t)s.NPmax * 4 * sizeof(int));         // owned position -> (row r, l, index inside the task's [B*L] arrays, weight set)
    s.total = o;
    return s;
}

// PT (per-task weights): n tile nt holds the Bt <= 8 rows of task nt (NT = G <= 4), the A fragments of every (k16 step, task) are
// streamed through the per-warp ring, the small attention weights are staged for the (at most two) tasks of the owned positions.
template <int NT, bool PT>
__global__ void __launch_bounds__(kMT, 1) k_attn_fwd_mma(AttnChainParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float zn_s[kBMax];
    const int T = p.T, L = p.L, Ha = p.Ha, A = p.A, F = p.F, Kl = p.Kl, H4 = 4 * Ha;
    const Grp gr{p.G, p.B, p.G * p.B, p.tstride};
    const int R = gr.R, Bt = gr.Bt, BtL = Bt * L, RL = R * L, pl = (Kl - 1) / 2;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lg = lane >> 2, lj = lane & 3;
    // K slice of this warp, rotated by the CTA index: the 148 CTAs stream the SAME h / dz rows, without the rotation they would all
    // ask the same L2 slices for the same lines at the same moment
    const int wsl = MSA_SLICE_ROT ? (w + blockIdx.x) % kMW : w;
    const AttnFwdMmaLay lay = attn_fwd_mma_layout(R, L, Ha, A, F, Kl, ncta, NT, PT ? 1 : 0);
    const int KS = lay.KS, LP = lay.LP, LH = lay.LH, RP = lay.RP;
    constexpr int NTP = NT < 2 ? NT : 2;
    uint4* Afrag = reinterpret_cast<uint4*>(smem_raw + lay.afrag);
    uint4* ring = Afrag + (size_t)w * kRingF * 128;      // PT: this warp's kRingF stages of [m tile][hi|lo][32] uint4
    float* part = reinterpret_cast<float*>(smem_raw + lay.part);
    float* es = part;                                  // [R*L]        gathered energies (after the tile reduction)
    float* q_s = part + ((RL + 3) & ~3);               // [NRown][A]   gathered queries of the owned rows
    float* as_ = reinterpret_cast<float*>(smem_raw + lay.as_);     // [R][LP]   a(t-1)
    float* ah = reinterpret_cast<float*>(smem_raw + lay.ah);       // [2][NRown][LH] a(t-1), cum(t-1) of the owned rows, zero halo
    float* wldT = reinterpret_cast<float*>(smem_raw + lay.wldT);   // [F][A]
    float* wloc_s = reinterpret_cast<float*>(smem_raw + lay.wloc); // [weight set][2 * Kl][F] (channel-major taps, filter innermost)
    float* vs = reinterpret_cast<float*>(smem_raw + lay.vs);       // [weight set][A + 4]: v, b_v
    float* pm_s = reinterpret_cast<float*>(smem_raw + lay.pm);     // [np][A]
    float* pre_s = reinterpret_cast<float*>(smem_raw + lay.pre);   // [np][A]  loc + pm, then v * tanh(.)
    float* cf_s = reinterpret_cast<float*>(smem_raw + lay.cf);     // [np][F]
    float* zm = reinterpret_cast<float*>(smem_raw + lay.zm);       // [32][RP] context term of the local gate rows
    long long* rowoff = reinterpret_cast<long long*>(smem_raw + lay.rowoff);
    long long* posoff = reinterpret_cast<long long*>(smem_raw + lay.posoff);
    int* posrl = reinterpret_cast<int*>(smem_raw + lay.posrl);

    const int u0 = part_lo(cta, Ha, ncta), u1 = part_lo(cta + 1, Ha, ncta), U = u1 - u0;
    const int p0 = part_lo(cta, RL, ncta), p1 = part_lo(cta + 1, RL, ncta), np = p1 - p0;
    const int d0 = part_lo(cta, A, ncta), d1 = part_lo(cta + 1, A, ncta);
    const bool has_q = d1 > d0;                        // at most one attention dim per CTA (launcher checks A <= ncta, U <= 7)
    const int r_lo = np > 0 ? p0 / L : 0, r_hi = np > 0 ? (p1 - 1) / L : -1, nrown = r_hi - r_lo + 1;
    // position pp = r*L + l of the group -> offset inside a [.. per task ..][Bt*L] array of task g = r / Bt
    auto pos_task = [&](int pp) { return (pp / L) / Bt; };

    // ---- one-time staging ----
    if (!PT) {
        for (int idx = threadIdx.x; idx < kMW * KS * 2 * 32; idx += kMT) {
            uint4 hi, lo;
            attn_fwd_frag(p.whh, p.wq, Ha, KS, u0, U, has_q, d0, idx, hi, lo);
            const size_t o = (size_t)(idx >> 5) * 2 * 32;
            Afrag[o + (idx & 31)] = hi;
            Afrag[o + 32 + (idx & 31)] = lo;
        }
    }
    // small attention weights: one set, or (PT) one per task of the owned positions (at most two: the launcher checks NPmax <= L)
    const int g_lo = r_lo / Bt, nset = PT ? (np > 0 ? r_hi / Bt - g_lo + 1 : 0) : 1;
    const int WLS = 2 * Kl * F, WDS = F * A, VSS = A + 4;      // strides of a weight set
    for (int ws = 0; ws < nset; ++ws) {
        const float* wloc_g = PT ? p.wloc_g[g_lo + ws] : p.wloc;
        const float* wld_g = PT ? p.wld_g[g_lo + ws] : p.wld;
        const float* v_g = PT ? p.v_g[g_lo + ws] : p.v;
        const float* bv_g = PT ? p.bv_g[g_lo + ws] : p.bv;
        for (int idx = threadIdx.x; idx < F * 2 * Kl; idx += kMT) {
            const int f = idx / (2 * Kl), ck = idx % (2 * Kl);
            wloc_s[ws * WLS + ck * F + f] = __ldg(wloc_g + idx);
        }
        for (int idx = threadIdx.x; idx < A * F; idx += kMT) {
            const int d = idx / F, f = idx % F;
            wldT[ws * WDS + f * A + d] = __ldg(wld_g + idx);
        }
        for (int idx = threadIdx.x; idx < A; idx += kMT) vs[ws * VSS + idx] = __ldg(v_g + idx);
        if (threadIdx.x == 0) vs[ws * VSS + A] = __ldg(bv_g);
    }
    for (int idx = threadIdx.x; idx < np * A; idx += kMT) {
        const int pi = idx / A, d = idx - pi * A, pp = p0 + pi, g = pos_task(pp);
        pm_s[idx] = __ldg(p.pm + g * gr.tstride + (size_t)(pp - g * BtL) * A + d);
    }
    for (int idx = threadIdx.x; idx < R * LP; idx += kMT) as_[idx] = 0.f;
    for (int idx = threadIdx.x; idx < 2 * lay.NRown * LH; idx += kMT) ah[idx] = 0.f;
    for (int idx = threadIdx.x; idx < 32 * RP; idx += kMT) zm[idx] = 0.f;
    for (int i = threadIdx.x; i < np; i += kMT) {       // cum fed to the conv at t = 0
        const int pp = p0 + i, g = pos_task(pp);
        p.cum[g * gr.tstride + (pp - g * BtL)] = 0.f;
    }
    for (int r = threadIdx.x; r < R; r += kMT) {
        rowoff[r] = (long long)gr.task(r) * gr.tstride;
        rowoff[32 + r] = rowoff[r] + (long long)gr.brow(r) * L;
    }
    for (int i = threadIdx.x; i < np; i += kMT) {
        const int pp = p0 + i, r = pp / L, g = r / Bt;
        posoff[i] = (long long)g * gr.tstride;
        posrl[4 * i] = r;
        posrl[4 * i + 1] = pp - r * L;
        posrl[4 * i + 2] = pp - g * BtL;
        posrl[4 * i + 3] = PT ? g - g_lo : 0;
    }

    // point-wise role: thread c < NT*64 owns cell (slot ul, row r); slot 7 of a query-owning CTA is the query "cell"
    const int c_nt = threadIdx.x >> 6, c_par = (threadIdx.x >> 5) & 1, c_ln = threadIdx.x & 31;
    const int c_ul = c_ln >> 2, c_col = (c_ln & 3) * 2 + c_par;      // column of the n tile
    const int c_r = PT ? c_nt * Bt + c_col : c_nt * 8 + c_col;
    const bool c_in = (int)threadIdx.x < NT * 64 && (PT ? (c_col < Bt && c_nt < p.G) : c_r < R);
    const bool pw = c_in && c_ul < U, qcell = c_in && has_q && c_ul == 7;
    const int c_g = c_in ? gr.task(c_r) : 0, c_b = c_in ? gr.brow(c_r) : 0, c_u = u0 + c_ul;
    const float* xw_c = p.xw + c_g * gr.tstride;
    float* ha_c = p.ha + c_g * gr.tstride;
    float* ca_c = p.ca + c_g * gr.tstride;
    float* ga_c = p.ga + c_g * gr.tstride;
    float* q_c = p.q + c_g * gr.tstride;
    const uint8_t* mask_c = pw ? (p.G > 1 ? p.mask_g[c_g] : p.mask) : nullptr;
    float xz[4] = {0.f, 0.f, 0.f, 0.f}, zacc[4] = {0.f, 0.f, 0.f, 0.f}, cstate = 0.f;
    unsigned char mk = 1;
    auto fetch = [&](int t) {
        const size_t zb = ((size_t)t * Bt + c_b) * H4;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) xz[gate] = __ldg(xw_c + zb + (size_t)gate * Ha + c_u);
        if (mask_c) mk = mask_c[((size_t)t * Bt + c_b) * Ha + c_u];
    };
    if (pw) fetch(0);

    // streaming role of the tile product: lane (lg, lj) of warp w loads h[row nt*8 + lg][4 columns] per k16 step
    const float* hrow[NT];
    bool rowok[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int r = PT ? nt * Bt + lg : nt * 8 + lg;
        rowok[nt] = PT ? (lg < Bt && nt < p.G) : r < R;
        const int rr = rowok[nt] ? r : 0;
        hrow[nt] = p.ha + gr.task(rr) * gr.tstride + (size_t)gr.brow(rr) * Ha;
    }
    // PT: the ring of streamed A fragments; chunk (k16 step s, task nt) in the order the product consumes them, period KS * NT
    const uint4* fsrc = PT ? p.wfrag + ((size_t)cta * kMW + wsl) * KS * 128 + lane : nullptr;
    int rs = 0, rnt = 0, rslot = 0;      // next chunk to issue and the ring slot it goes to
    auto ring_issue = [&]() {
        const uint4* src = fsrc + (size_t)rnt * p.wfrag_stride + (size_t)rs * 128;
        uint4* dst = ring + rslot * 128 + lane;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) cp_async16(dst + qd * 32, src + qd * 32);
        cp_async_commit();
        if (++rnt == NT) { rnt = 0; if (++rs == KS) rs = 0; }
        if (++rslot == kRingF) rslot = 0;
    };
    if (PT) {
#pragma unroll
        for (int i = 0; i < kRingF; ++i) ring_issue();
    }
    const bool vecL = (L & 3) == 0;
    ChainProf<true> prof;      // per-phase cycles of thread 0 when a buffer is passed (profiles/), one predictable branch otherwise
    prof.start(p.prof, nullptr, 0);
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        // ---- P1: context term zm[rl][r] = sum_l MW[row rl][r][l] * a(t-1)[r][l] (zero at t = 0: zm starts zeroed) ----
        if (t > 0 && U > 0) {
            // one dot product per half-warp; warp w owns the local rows w, w + 16, ... (gate-major list of the 4*U real rows), its two
            // half-warps alternate over the batch rows; the MW loads of eight dots are issued before the first one is used (the slice
            // streams from L2: 4*U*R*L floats per step)
            const int hw = lane >> 4, hl = lane & 15;
            constexpr int PB = 8;
            for (int j = w; j < 4 * U; j += kMW) {
                const int gate = j / U, ul = j - gate * U, rl = gate * 8 + ul;
                const float* mbase = p.mw_rm + (size_t)(gate * Ha + u0 + ul) * BtL;
                for (int r0 = hw; r0 < R; r0 += 2 * PB) {
                    float acc[PB];
                    if (vecL && L <= 64) {
                        float4 mv[PB];
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int r = r0 + 2 * k;
                            mv[k] = (r < R && hl * 4 < L) ? __ldcg(reinterpret_cast<const float4*>(mbase + rowoff[32 + r]) + hl)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int r = r0 + 2 * k;
                            acc[k] = (r < R && hl * 4 < L) ? dot4(mv[k], *reinterpret_cast<const float4*>(as_ + r * LP + hl * 4)) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int r = r0 + 2 * k;
                            acc[k] = 0.f;
                            if (r < R) {
                                const float* mrow = mbase + rowoff[32 + r];
                                for (int l = hl; l < L; l += 16) acc[k] += __ldcg(mrow + l) * as_[r * LP + l];
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < PB; ++k) {
                        const int r = r0 + 2 * k;
                        float a = acc[k];
                        a += __shfl_xor_sync(0xffffffffu, a, 1);
                        a += __shfl_xor_sync(0xffffffffu, a, 2);
                        a += __shfl_xor_sync(0xffffffffu, a, 4);
                        a += __shfl_xor_sync(0xffffffffu, a, 8);
                        if (hl == 0 && r < R) zm[rl * RP + r] = a;
                    }
                }
            }
        }
        prof.mark(0, t);
        __syncthreads();
        // ---- P2: attention LSTMCell of the owned cells; publishes h_a'(t) ----
        if (pw) {
            float z[4];
#pragma unroll
            for (int gate = 0; gate < 4; ++gate) z[gate] = zacc[gate] + xz[gate] + zm[(gate * 8 + c_ul) * RP + c_r];
            const float ai = fast_sigmoid(z[0]), af = fast_sigmoid(z[1]), ag = fast_tanh(z[2]), ao = fast_sigmoid(z[3]);
            const float cn = af * cstate + ai * ag;
            cstate = cn;
            float hv = ao * fast_tanh(cn);
            if (mask_c) hv = mk ? hv * p.drop_scale : 0.f;
            const size_t hb = ((size_t)t * Bt + c_b) * Ha + c_u, zb = ((size_t)t * Bt + c_b) * H4;
            st_pub(ha_c + hb, hv);
            ca_c[hb] = cn;
            ga_c[zb + 0 * (size_t)Ha + c_u] = ai;
            ga_c[zb + 1 * (size_t)Ha + c_u] = af;
            ga_c[zb + 2 * (size_t)Ha + c_u] = ag;
            ga_c[zb + 3 * (size_t)Ha + c_u] = ao;
            if (t + 1 < T) fetch(t + 1);
        }
        prof.mark(1, t);
        // ---- P3 (shadow of the h hand-off): location features of the owned positions (forward_attn.py:121-127) ----
        // one thread per (position, filter): a warp = the F filters of one position (F = 32), so the alignment / cumulative windows are
        // broadcast reads and the tap-major weights conflict-free; two independent accumulators (previous alignment, cumulative weights)
        for (int it = threadIdx.x; it < np * F; it += kMT) {
            const int pi = it / F, f = it - pi * F;
            const int ro = posrl[4 * pi] - r_lo, l = posrl[4 * pi + 1];
            const float* a0 = ah + (size_t)ro * LH + l;
            const float* a1 = ah + (size_t)(lay.NRown + ro) * LH + l;
            const float* wl = wloc_s + posrl[4 * pi + 3] * WLS + f;
            float c0 = 0.f, c1 = 0.f;
#pragma unroll 4
            for (int k = 0; k < Kl; ++k) {
                c0 += wl[k * F] * a0[k];
                c1 += wl[(Kl + k) * F] * a1[k];
            }
            const float cf = c0 + c1;
            cf_s[it] = cf;
            p.convf[posoff[pi] + ((size_t)t * BtL + posrl[4 * pi + 2]) * F + f] = cf;
        }
        __syncthreads();
        if ((A & 3) == 0) {       // four attention dims per thread: one 128-bit weight load per filter
            for (int it = threadIdx.x; it < np * (A >> 2); it += kMT) {
                const int pi = it / (A >> 2), d = (it - pi * (A >> 2)) * 4;
                float4 acc = *reinterpret_cast<const float4*>(pm_s + pi * A + d);
                const float* wd = wldT + posrl[4 * pi + 3] * WDS;
                for (int f = 0; f < F; ++f) {
                    const float4 wv = *reinterpret_cast<const float4*>(wd + f * A + d);
                    const float c = cf_s[pi * F + f];
                    acc.x += wv.x * c; acc.y += wv.y * c; acc.z += wv.z * c; acc.w += wv.w * c;
                }
                *reinterpret_cast<float4*>(pre_s + pi * A + d) = acc;
            }
        } else {
            for (int it = threadIdx.x; it < np * A; it += kMT) {
                const int pi = it / A, d = it - pi * A;
                float l0 = pm_s[it];
                const float* wd = wldT + posrl[4 * pi + 3] * WDS;
                for (int f = 0; f < F; ++f) l0 += wd[f * A + d] * cf_s[pi * F + f];
                pre_s[it] = l0;
            }
        }
        prof.mark(2, t);
        // ---- hand-off 1: h_a'(t) of every row; tile product [gate rows + query row] x h_a'(t) ----
        {
            const int rl_ = R - 1;
            const float* hlast = p.ha + gr.task(rl_) * gr.tstride + ((size_t)t * Bt + gr.brow(rl_)) * Ha;
            gate_waitq(ncta, [&](int c) { const int e = part_lo(c + 1, Ha, ncta); return e > part_lo(c, Ha, ncta) ? hlast + e - 1 : nullptr; }, p.abort_word);
            __syncthreads();
        }
        prof.mark(3, t);
        {
            float acc[NT][2][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[nt][mt][i] = 0.f;
            const size_t toff = (size_t)t * Bt * Ha;
            constexpr int PF = MSA_PF_FWD;
            float4 hv[PF][NT];
            auto issue = [&](int s, float4 (&dst)[NT]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    dst[nt] = (rowok[nt] && col < Ha) ? ld_poll4(hrow[nt] + toff + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto process = [&](int s, float4 (&cur)[NT]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
                uint32_t bhi[NT][2], blo[NT][2];
                unsigned int cmax = 0u;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) cmax = umax_acc(cmax, cur[nt]);
                if (cmax == kCanary) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        if (rowok[nt] && col < Ha && !ready4(cur[nt])) cur[nt] = poll4_slow(hrow[nt] + toff + col, p.abort_word);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    split2(cur[nt].x, cur[nt].y, bhi[nt][0], blo[nt][0]);
                    split2(cur[nt].z, cur[nt].w, bhi[nt][1], blo[nt][1]);
                }
                if (PT) {      // the fragments of (s, task nt) arrive through the ring, in this order
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        cp_async_wait<kRingF - 1>();
                        const uint4* af = ring + rslot * 128 + lane;
                        const uint4 ahi0 = af[0], alo0 = af[32], ahi1 = af[64], alo1 = af[96];
                        mma3(acc[nt][0], ahi0, alo0, bhi[nt], blo[nt]);
                        mma3(acc[nt][1], ahi1, alo1, bhi[nt], blo[nt]);
                        ring_issue();      // refills the slot just read with the chunk kRingF ahead
                    }
                } else {
                    const uint4* af = Afrag + ((size_t)wsl * KS + s) * 2 * 2 * 32;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint4 ahi = af[(mt * 2 + 0) * 32 + lane], alo = af[(mt * 2 + 1) * 32 + lane];
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma3(acc[nt][mt], ahi, alo, bhi[nt], blo[nt]);
                    }
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
            // cross-warp reduction of the K slices, two n tiles per round; the cell threads keep the sums in registers
#pragma unroll
            for (int rd = 0; rd < (NT + 1) / 2; ++rd) {
                if (rd > 0) __syncthreads();
#pragma unroll
                for (int k = 0; k < NTP; ++k) {
                    const int nt = rd * 2 + k;
                    if (nt < NT) {
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int i = 0; i < 4; ++i) part[((((size_t)w * NTP + k) * 2 + mt) * 4 + i) * 32 + lane] = acc[nt][mt][i];
                    }
                }
                __syncthreads();
                if (c_in && (c_nt >> 1) == rd && (pw || qcell)) {
                    const int k = c_nt & 1;
#pragma unroll
                    for (int gate = 0; gate < 4; ++gate) {
                        float sum = 0.f;
                        const float* pp = part + ((((size_t)k) * 2 + (gate >> 1)) * 4 + (gate & 1) * 2 + c_par) * 32 + c_ln;
#pragma unroll
                        for (int ww = 0; ww < kMW; ++ww) sum += pp[(size_t)ww * NTP * 2 * 4 * 32];
                        zacc[gate] = sum;
                    }
                    if (qcell) st_pub(q_c + ((size_t)t * Bt + c_b) * A + d0, zacc[0]);      // q(t)[r][d0]
                }
            }
        }
        prof.mark(4, t);
        __syncthreads();      // the partial-tile region becomes the gather buffer
        // ---- hand-off 2: queries of the owned rows ----
        if ((A & 3) == 0) {
            for (int i = threadIdx.x; i < nrown * (A >> 2); i += kMT) {
                const int ro = i / (A >> 2), d4 = i - ro * (A >> 2), r = r_lo + ro;
                reinterpret_cast<float4*>(q_s)[i] = pollq4(p.q + gr.task(r) * gr.tstride + ((size_t)t * Bt + gr.brow(r)) * A + d4 * 4, p.abort_word);
            }
        } else {
            for (int i = threadIdx.x; i < nrown * A; i += kMT) {
                const int ro = i / A, d = i - ro * A, r = r_lo + ro;
                q_s[i] = pollq1(p.q + gr.task(r) * gr.tstride + ((size_t)t * Bt + gr.brow(r)) * A + d, p.abort_word);
            }
        }
        __syncthreads();
        prof.mark(5, t);
        // ---- P5: energies of the owned positions (forward_attn.py:128-131); publishes e(t) ----
        for (int it = threadIdx.x; it < np * A; it += kMT) {
            const int pi = it / A, d = it - pi * A;
            const float sv = fast_tanh(q_s[(posrl[4 * pi] - r_lo) * A + d] + pre_s[it]);
            p.s[posoff[pi] + ((size_t)t * BtL + posrl[4 * pi + 2]) * A + d] = sv;
            pre_s[it] = vs[posrl[4 * pi + 3] * VSS + d] * sv;
        }
        __syncthreads();
        for (int pi = w; pi < np; pi += kMW) {
            float e = 0.f;
            for (int d = lane; d < A; d += 32) e += pre_s[pi * A + d];
            e = warp_sum(e);
            if (lane == 0) st_pub(p.e + posoff[pi] + (size_t)t * BtL + posrl[4 * pi + 2], e + vs[posrl[4 * pi + 3] * VSS + A]);
        }
        prof.mark(6, t);
        // ---- hand-off 3: all energies of step t; a(t) = normalise(e(t)); cum += a(t) (forward_attn.py:200-210) ----
        if ((BtL & 3) == 0) {      // 128-bit polls: one L2 round trip for the whole gather
            for (int i4 = threadIdx.x; i4 < (RL >> 2); i4 += kMT) {
                const int i = i4 * 4, g = i / BtL;
                reinterpret_cast<float4*>(es)[i4] = pollq4(p.e + g * gr.tstride + (size_t)t * BtL + (i - g * BtL), p.abort_word);
            }
        } else {
            for (int i = threadIdx.x; i < RL; i += kMT) {
                const int g = pos_task(i);
                es[i] = pollq1(p.e + g * gr.tstride + (size_t)t * BtL + (i - g * BtL), p.abort_word);
            }
        }
        __syncthreads();
        for (int r = w; r < R; r += kMW) {
            float m = 0.f;
            if (p.norm == 0) {
                m = -INFINITY;
                for (int l = lane; l < L; l += 32) m = fmaxf(m, es[r * L + l]);
                m = warp_max(m);
            }
            float sum = 0.f;
            for (int l = lane; l < L; l += 32) {
                const float x = p.norm == 0 ? __expf(es[r * L + l] - m) : fast_sigmoid(es[r * L + l]);
                es[r * L + l] = x;
                sum += x;
            }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            const bool own = r >= r_lo && r <= r_hi;
            for (int l = lane; l < L; l += 32) {
                const float a = es[r * L + l] * inv;
                as_[r * LP + l] = a;
                if (own) {
                    ah[(0 * lay.NRown + (r - r_lo)) * LH + pl + l] = a;
                    ah[(1 * lay.NRown + (r - r_lo)) * LH + pl + l] += a;
                }
            }
            if (lane == 0) zn_s[r] = sum;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < np; i += kMT) {
            const int r = posrl[4 * i], l = posrl[4 * i + 1];
            const size_t o = posoff[i] + (size_t)posrl[4 * i + 2];
            p.align[o + (size_t)t * BtL] = as_[r * LP + l];
            if (t + 1 < T) p.cum[o + (size_t)(t + 1) * BtL] = ah[(1 * lay.NRown + (r - r_lo)) * LH + pl + l];
        }
        if (cta == 0 && (int)threadIdx.x < R) {
            const int r = threadIdx.x;
            p.znorm[gr.task(r) * gr.tstride + (size_t)t * Bt + gr.brow(r)] = zn_s[r];
        }
        prof.mark(7, t);
    }
    if (PT) cp_async_wait<0>();      // the ring runs kRingF chunks ahead of the last step
}

static bool attn_fwd_mma_ok(const msa_config& cfg, int R, int L, int sm_count, size_t smem_limit) {
    const int Ha = cfg.attn_rnn_dim, A = cfg.attn_dim;
    if (cfg.forward_attn || R < 1 || R > 32 || Ha % 4 != 0 || cfg.loc_filters > 32) return false;
    if (A > sm_count) return false;                                  // one query dim per CTA
    if ((Ha + sm_count - 1) / sm_count > 7) return false;          // slot 7 of the tile is the query row
    return attn_fwd_mma_layout(R, L, Ha, A, cfg.loc_filters, cfg.loc_kernel, sm_count, (R + 7) / 8).total + 256 <= smem_limit;
}

// =====================================================================================================================
// Attention chain, backward (reverse-time counterpart of k_attn_fwd_mma; same arithmetic as attn_chain.cu::k_attn_chain_bwd
// without the forward-attention options) for R = G*B rows.  Per step t, descending:
//   A  dz_a(t+1) arrives: context-path term of the owned positions pout = MWp . dz_a(t+1) (one warp per position, MWp rows
//      streamed from L2); publishes d a(t) of the owned positions
//   C  recurrent term dz_a(t+1) . W_hh for the owned units: bf16x3 tensor-core tile, accumulators stay in registers
//   D  d a(t) of every position arrives; normalisation backward -> d e(t); dq of the owned attention dim; dS, d(conv features)
//      of the owned positions; publishes dq(t), dconvf(t)
//   F  dq(t) arrives: dq(t) . W_q for the owned units goes into the SAME accumulators (eight more k16 steps), one cross-warp
//      reduction, LSTM point-wise backward; publishes dz_a(t)
//   H  dconvf(t) of the +-pad neighbours arrives: location-conv backward -> d a(t-1) ("previous alignment" channel) and the
//      running d cum
// =====================================================================================================================
struct AttnBwdMmaLay {
    size_t bfrag, bq, part, das, als, wld, wloc, vs, dss, dcw, small, total;      // byte offsets
    int KS, KQ, NPmax, NRown, CKP, WIN, FP;
};
__host__ __device__ inline AttnBwdMmaLay attn_bwd_mma_layout(int R, int L, int Ha, int A, int F, int Kl, int ncta, int MTL, int pt = 0) {
    AttnBwdMmaLay s;
    s.KS = (4 * Ha + 255) / 256;
    s.KQ = (A + 15) / 16;
    s.CKP = 2 * Kl + 1;
    s.FP = F + 1;
    s.NPmax = (R * L + ncta - 1) / ncta;
    s.NRown = (s.NPmax + L - 2) / L + 1;
    s.WIN = s.NPmax + Kl - 1 < L ? s.NPmax + Kl - 1 : L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    // resident B fragments, or (per-task weights) the ring: stages of one k16 step x MTL tasks per warp (fragments + dz); W_q^T per task
    s.bfrag = take(pt ? (size_t)kMW * ring_depth_bwd(MTL) * MTL * (32 + (R / MTL) * 4) * sizeof(uint4) : (size_t)kMW * s.KS * 2 * 32 * sizeof(uint2));
    s.bq = take((size_t)(pt ? MTL : 1) * s.KQ * 2 * 32 * sizeof(uint2));
    const int NW = pt ? 2 : 1;
    // partial tiles [kMW][MTL][4][32]; earlier in the step the region holds tq [R*L] and the dconvf partials [4][NPmax][F]
    size_t pb = (size_t)kMW * MTL * 4 * 32 * sizeof(float);
    const size_t alt = ((size_t)((R * L + 3) & ~3) + (size_t)4 * s.NPmax * F) * sizeof(float);
    if (alt > pb) pb = alt;
    s.part = take(pb);
    s.das = take((size_t)R * L * sizeof(float));
    s.als = take((size_t)R * L * sizeof(float));
    s.wld = take((size_t)NW * A * F * sizeof(float));
    s.wloc = take((size_t)NW * F * s.CKP * sizeof(float));
    s.vs = take((size_t)(NW * A + 8) * sizeof(float));           // v per weight set, then v[d0] of every task
    s.dss = take((size_t)s.NPmax * A * sizeof(float));
    s.dcw = take((size_t)s.NRown * s.WIN * s.FP * sizeof(float));
    s.small = take((size_t)(3 * s.NPmax + 64) * sizeof(float));
    s.total = o;
    return s;
}

// W_hh^T fragment idx = (K slice ww * KS + k16 step s) * 32 + lane for the CTA that owns units [u0, u0 + U): n = unit lane >> 2,
// k rows kb + 4j .. 4j + 3 (gate rows of W_hh)
__device__ __forceinline__ void attn_bwd_frag(const float* __restrict__ whh, int Ha, int KS, int u0, int U, int idx, uint2& hi, uint2& lo) {
    const int ln = idx & 31, s = (idx >> 5) % KS, ww = (idx >> 5) / KS;
    const int g = ln >> 2, j = ln & 3;
    const int kb = (ww * KS + s) * 16 + 4 * j;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = (g < U && kb + c < 4 * Ha) ? __ldg(whh + (size_t)(kb + c) * Ha + u0 + g) : 0.f;
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
}
// the fragments of G tasks in the order the PT kernel streams them: [task][cta][ww][s][lane] {hi.x, hi.y, lo.x, lo.y}
__global__ void __launch_bounds__(kMT) k_attn_frag_bwd(FragSrc src, int Ha, int64_t stride, uint4* out) {
    const int task = blockIdx.y, cta = blockIdx.x, ncta = gridDim.x;
    const int KS = (4 * Ha + 255) / 256;
    const int u0 = part_lo(cta, Ha, ncta), U = part_lo(cta + 1, Ha, ncta) - u0;
    uint4* o = out + (size_t)task * stride + (size_t)cta * kMW * KS * 32;
    for (int idx = threadIdx.x; idx < kMW * KS * 32; idx += kMT) {
        uint2 hi, lo;
        attn_bwd_frag(src.whh[task], Ha, KS, u0, U, idx, hi, lo);
        o[idx] = make_uint4(hi.x, hi.y, lo.x, lo.y);
    }
}

// PT (per-task weights): m tile mt holds the Bt <= 8 rows of task mt (MTL = G <= 4; fragment rows 8..15 are zero), the B fragments of
// every (k16 step, task) are streamed through the per-warp ring, W_q^T and the small attention weights are staged per task.
template <int MTL, bool PT>
__global__ void __launch_bounds__(kMT, 1) k_attn_bwd_mma(AttnChainBwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float zn_s[kBMax];
    const int T = p.T, L = p.L, Ha = p.Ha, A = p.A, F = p.F, Kl = p.Kl, H4 = 4 * Ha;
    const Grp gr{p.G, p.B, p.G * p.B, p.tstride};
    const int R = gr.R, Bt = gr.Bt, BtL = Bt * L, RL = R * L, pl = (Kl - 1) / 2;
    const int ncta = gridDim.x, cta = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lg = lane >> 2, lj = lane & 3;
    // K slice of this warp, rotated by the CTA index: the 148 CTAs stream the SAME h / dz rows, without the rotation they would all
    // ask the same L2 slices for the same lines at the same moment
    const int wsl = MSA_SLICE_ROT ? (w + blockIdx.x) % kMW : w;
    const AttnBwdMmaLay lay = attn_bwd_mma_layout(R, L, Ha, A, F, Kl, ncta, MTL, PT ? 1 : 0);
    const int KS = lay.KS, KQ = lay.KQ, CKP = lay.CKP, WIN = lay.WIN, FP = lay.FP;
    constexpr int HR = PT ? 1 : 2;                                     // row halves of an m tile that can hold real rows
    uint2* Bfrag = reinterpret_cast<uint2*>(smem_raw + lay.bfrag);    // [kMW][KS][hi|lo][32]  W_hh^T slice
    constexpr int RD = ring_depth_bwd(MTL);                            // PT: ring stages per warp
    const int ZS = Bt * 4;                                              // PT: dz slots (float4) per task and stage: lanes with lg < Bt
    const int STG = MTL * (32 + ZS);                                    // PT: uint4 words per stage: [MTL][32] fragments, [MTL][ZS] dz
    uint4* ring = reinterpret_cast<uint4*>(smem_raw + lay.bfrag) + (size_t)w * RD * STG;
    uint2* Bq = reinterpret_cast<uint2*>(smem_raw + lay.bq);          // [KQ][hi|lo][32]       W_q^T slice
    float* part = reinterpret_cast<float*>(smem_raw + lay.part);
    float* tq_s = part;                                                // [R*L]      d e * (1 - s^2) for the owned attention dim
    float* das = reinterpret_cast<float*>(smem_raw + lay.das);        // [R*L]      d a(t), then d e(t)
    float* als = reinterpret_cast<float*>(smem_raw + lay.als);        // [R*L]      a(t)
    float* wld_s = reinterpret_cast<float*>(smem_raw + lay.wld);      // [A][F]
    float* wloc_s = reinterpret_cast<float*>(smem_raw + lay.wloc);    // [F][CKP]
    float* vs = reinterpret_cast<float*>(smem_raw + lay.vs);
    float* ds_s = reinterpret_cast<float*>(smem_raw + lay.dss);       // [np][A]
    float* dcw = reinterpret_cast<float*>(smem_raw + lay.dcw);        // [NRown][WIN][FP] gathered d(conv features) window
    float* gcum_s = reinterpret_cast<float*>(smem_raw + lay.small);   // [np] running d cum
    float* dprev_s = gcum_s + lay.NPmax;                               // [np] d a(t-1) through the "previous alignment" channel
    float* pout_s = dprev_s + lay.NPmax;                               // [np]

    const int u0 = part_lo(cta, Ha, ncta), u1 = part_lo(cta + 1, Ha, ncta), U = u1 - u0;
    const int p0 = part_lo(cta, RL, ncta), p1 = part_lo(cta + 1, RL, ncta), np = p1 - p0;
    const int d0 = part_lo(cta, A, ncta), d1 = part_lo(cta + 1, A, ncta);
    const bool has_q = d1 > d0;
    const int r_lo = np > 0 ? p0 / L : 0, r_hi = np > 0 ? (p1 - 1) / L : -1, nrown = r_hi - r_lo + 1;
    auto pos_task = [&](int pp) { return (pp / L) / Bt; };
    // window of output positions whose conv features feed the owned positions of own row ro: [wlo(ro), whi(ro)]
    auto own_lmin = [&](int ro) { return ro == 0 ? p0 - r_lo * L : 0; };
    auto own_lmax = [&](int ro) { return ro == nrown - 1 ? (p1 - 1) - r_hi * L : L - 1; };
    auto wlo = [&](int ro) { const int a = own_lmin(ro) - pl; return a > 0 ? a : 0; };
    auto whi = [&](int ro) { const int a = own_lmax(ro) + pl; return a < L - 1 ? a : L - 1; };

    // ---- one-time staging ----
    if (!PT) {
        for (int idx = threadIdx.x; idx < kMW * KS * 32; idx += kMT) {
            uint2 hi, lo;
            attn_bwd_frag(p.whh, Ha, KS, u0, U, idx, hi, lo);
            const size_t o = (size_t)(idx >> 5) * 2 * 32;
            Bfrag[o + (idx & 31)] = hi;
            Bfrag[o + 32 + (idx & 31)] = lo;
        }
    }
    for (int tk = 0; tk < (PT ? p.G : 1); ++tk) {      // W_q^T slice(s): [task][KQ][hi|lo][32]
        const float* wq = PT ? p.wq_g[tk] : p.wq;
        for (int idx = threadIdx.x; idx < KQ * 32; idx += kMT) {
            const int ln = idx & 31, s = idx >> 5;
            const int g = ln >> 2, j = ln & 3;
            const int kb = s * 16 + 4 * j;
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = (g < U && kb + c < A) ? __ldg(wq + (size_t)(kb + c) * Ha + u0 + g) : 0.f;
            uint2 hi, lo;
            split2(v[0], v[1], hi.x, lo.x);
            split2(v[2], v[3], hi.y, lo.y);
            Bq[((size_t)tk * KQ + s) * 2 * 32 + ln] = hi;
            Bq[((size_t)tk * KQ + s) * 2 * 32 + 32 + ln] = lo;
        }
    }
    // small attention weights: one set, or (PT) one per task of the owned positions (at most two: the launcher checks NPmax <= L)
    const int g_lo = r_lo / Bt, nset = PT ? (np > 0 ? r_hi / Bt - g_lo + 1 : 0) : 1;
    const int WLS = F * CKP, WDS = A * F;
    float* vq_s = vs + (PT ? 2 : 1) * A;               // v[d0] of every task (the dq term of the query-owning CTA)
    for (int ws = 0; ws < nset; ++ws) {
        const float* wloc_g = PT ? p.wloc_g[g_lo + ws] : p.wloc;
        const float* wld_g = PT ? p.wld_g[g_lo + ws] : p.wld;
        const float* v_g = PT ? p.v_g[g_lo + ws] : p.v;
        for (int idx = threadIdx.x; idx < A * F; idx += kMT) wld_s[ws * WDS + idx] = __ldg(wld_g + idx);
        for (int idx = threadIdx.x; idx < F * 2 * Kl; idx += kMT) {
            const int f = idx / (2 * Kl), ck = idx % (2 * Kl);
            wloc_s[ws * WLS + f * CKP + ck] = __ldg(wloc_g + idx);
        }
        for (int idx = threadIdx.x; idx < A; idx += kMT) vs[ws * A + idx] = __ldg(v_g + idx);
    }
    if (has_q && (int)threadIdx.x < p.G) vq_s[threadIdx.x] = __ldg((PT ? p.v_g[threadIdx.x] : p.v) + d0);
    auto pos_set = [&](int pp) { return PT ? (pp / L) / Bt - g_lo : 0; };      // weight set of an owned position
    for (int idx = threadIdx.x; idx < np; idx += kMT) { gcum_s[idx] = 0.f; dprev_s[idx] = 0.f; pout_s[idx] = 0.f; }

    // point-wise role: thread c < MTL*128 owns cell (row r, unit ul)
    const int c_mt = threadIdx.x >> 7, c_reg = (threadIdx.x >> 5) & 3, c_ln = threadIdx.x & 31;
    const int c_row = (c_reg >> 1) * 8 + (c_ln >> 2);                  // row of the m tile
    const int c_r = PT ? c_mt * Bt + c_row : c_mt * 16 + c_row, c_ul = (c_ln & 3) * 2 + (c_reg & 1);
    const bool pw = (int)threadIdx.x < MTL * 128 && c_ul < U && (PT ? (c_row < Bt && c_mt < p.G) : c_r < R);
    const int c_g = pw ? gr.task(c_r) : 0, c_b = pw ? gr.brow(c_r) : 0, c_u = u0 + c_ul;
    const float* ga_c = p.ga + c_g * gr.tstride;
    const float* ca_c = p.ca + c_g * gr.tstride;
    const float* dhe_c = p.dha_ext + c_g * gr.tstride;
    float* dza_c = p.dza + c_g * gr.tstride;
    const uint8_t* mask_c = pw ? (p.G > 1 ? p.mask_g[c_g] : p.mask) : nullptr;
    float gi[4] = {0.f, 0.f, 0.f, 0.f}, cc = 0.f, cp = 0.f, dhe = 0.f, dcarry = 0.f;
    unsigned char mk = 1;
    // forward stash of the attention part, fetched one step ahead: s of the owned positions (item = tid + k*kMT < np*A),
    // the d0 column of s for every position (tid + k*kMT < R*L), d a_ext of the owned positions handled by this warp
    constexpr int KV = 4;
    float sv_own[KV], sv_q[KV];
    auto s_addr = [&](int t, int pp, int d) {
        const int g = pos_task(pp);
        return p.s + g * gr.tstride + ((size_t)t * BtL + (pp - g * BtL)) * A + d;
    };
    auto fetch = [&](int t) {
        if (pw) {
            const size_t zb = ((size_t)t * Bt + c_b) * H4, hb = ((size_t)t * Bt + c_b) * Ha + c_u;
#pragma unroll
            for (int gate = 0; gate < 4; ++gate) gi[gate] = __ldg(ga_c + zb + (size_t)gate * Ha + c_u);
            cc = __ldg(ca_c + hb);
            cp = t > 0 ? __ldg(ca_c + hb - (size_t)Bt * Ha) : 0.f;
            dhe = __ldg(dhe_c + hb);
            if (mask_c) mk = mask_c[hb];
        }
#pragma unroll
        for (int k = 0; k < KV; ++k) {
            const int it = threadIdx.x + k * kMT;
            if (it < np * A) sv_own[k] = __ldg(s_addr(t, p0 + it / A, it % A));
            if (has_q && it < RL) sv_q[k] = __ldg(s_addr(t, it, d0));
        }
    };
    fetch(T - 1);

    // streaming role of the recurrent tile: lane (lg, lj) loads dz[rows mt*16 + lg, +8][4 gate rows] per k16 step
    const float* zrow[MTL][HR];
    bool rowok[MTL][HR];
#pragma unroll
    for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
        for (int hr = 0; hr < HR; ++hr) {
            const int r = PT ? mt * Bt + lg : mt * 16 + hr * 8 + lg;
            rowok[mt][hr] = PT ? (lg < Bt && mt < p.G) : r < R;
            const int rr = rowok[mt][hr] ? r : 0;
            zrow[mt][hr] = p.dza + gr.task(rr) * gr.tstride + (size_t)gr.brow(rr) * H4;
        }
    // PT: the ring of streamed B fragments and dz; stage = k16 step s of all MTL tasks.  One commit group per stage fill.
    const uint4* fsrc = PT ? p.wfrag + ((size_t)cta * kMW + wsl) * KS * 32 + lane : nullptr;
    int rslot = 0;                      // slot of the next stage to consume (= the oldest; refilled right after it was read)
    // per-lane constants of the ring traffic, hoisted out of the step loop (the product loop is bound by instruction issue: shared
    // addresses as 32-bit window offsets, one base pointer per task and operand)
    const uint32_t ring_sa = PT ? (uint32_t)__cvta_generic_to_shared(ring) : 0u;
    const uint32_t stg_bytes = (uint32_t)STG * 16u, w_off = (uint32_t)lane * 16u, z_off = (uint32_t)(MTL * 32 + lg * 4 + lj) * 16u;
    const int colbase = wsl * KS * 16 + 4 * lj;
    const bool lane_z = PT && lg < Bt;                 // rowok[mt][0] of every task (MTL == G)
    const uint4* fs[MTL];
    const float* zb[MTL];
#pragma unroll
    for (int mt = 0; mt < MTL; ++mt) {
        fs[mt] = PT ? fsrc + (size_t)mt * p.wfrag_stride : nullptr;
        zb[mt] = zrow[mt][0] + colbase;
    }
    auto ring_fill_dz = [&](int slot, int s, size_t toff) {      // dz(t+1) of k16 step s into its slots of the stage
        if (lane_z && colbase + s * 16 < H4) {
            const uint32_t sa = ring_sa + (uint32_t)slot * stg_bytes + z_off;
#pragma unroll
            for (int mt = 0; mt < MTL; ++mt) cp_async16_sa(sa + (uint32_t)(mt * ZS) * 16u, zb[mt] + toff + s * 16);
        }
    };
    auto ring_fill = [&](int slot, int s, bool with_dz, size_t toff) {      // weights of k16 step s (+ dz(t+1) of that step)
        const uint32_t sa = ring_sa + (uint32_t)slot * stg_bytes + w_off;
#pragma unroll
        for (int mt = 0; mt < MTL; ++mt) cp_async16_sa(sa + (uint32_t)mt * 512u, fs[mt] + (size_t)s * 32);
        if (with_dz) ring_fill_dz(slot, s, toff);
        cp_async_commit();
    };
    const bool ring_deep = KS >= RD;
    if (PT && ring_deep) {
#pragma unroll
        for (int i = 0; i < RD; ++i) ring_fill(i, i, false, 0);
    }
    ChainProf<true> prof;
    prof.start(p.prof, nullptr, 0);
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        float acc[MTL][4];
#pragma unroll
        for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
        // ---- A: dz_a(t+1) of every row ----
        if (t < T - 1) {
            const int rl_ = R - 1;
            const float* zlast = p.dza + gr.task(rl_) * gr.tstride + ((size_t)(t + 1) * Bt + gr.brow(rl_)) * H4 + (size_t)3 * Ha;
            gate_waitq(ncta, [&](int c) { const int e = part_lo(c + 1, Ha, ncta); return e > part_lo(c, Ha, ncta) ? zlast + e - 1 : nullptr; }, p.abort_word);
            __syncthreads();
        }
        prof.mark(0, T - 1 - t);
        // context-path recurrent term of the owned positions (one warp per position) and their d a(t)
        for (int i = w; i < np; i += kMW) {
            const int pp = p0 + i, r = pp / L, g = r / Bt;
            float a = 0.f;
            if (t < T - 1) {
                const float4* m4 = reinterpret_cast<const float4*>(p.mw_pm + g * gr.tstride + (size_t)(pp - g * BtL) * H4);
                const float* zr = p.dza + g * gr.tstride + ((size_t)(t + 1) * Bt + (r - g * Bt)) * H4;
                constexpr int NB = 4;
                for (int c0 = lane; c0 < (H4 >> 2); c0 += 32 * NB) {
                    float4 mv[NB], zv[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        const int c = c0 + k * 32;
                        if (c < (H4 >> 2)) { mv[k] = __ldcg(m4 + c); zv[k] = ld_poll4(zr + (size_t)c * 4); }
                    }
                    unsigned int cmax = 0u;
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        if (c0 + k * 32 < (H4 >> 2)) cmax = umax_acc(cmax, zv[k]);
                    if (cmax == kCanary) {
#pragma unroll
                        for (int k = 0; k < NB; ++k)
                            if (c0 + k * 32 < (H4 >> 2) && !ready4(zv[k])) zv[k] = poll4_slow(zr + (size_t)(c0 + k * 32) * 4, p.abort_word);
                    }
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        if (c0 + k * 32 < (H4 >> 2)) a += dot4(mv[k], zv[k]);
                }
                a = warp_sum(a);
            }
            if (lane == 0) {
                const size_t o = g * gr.tstride + (size_t)t * BtL + (pp - g * BtL);
                pout_s[i] = a;
                st_pub(p.dat + o, __ldg(p.da_ext + o) + a + gcum_s[i] + dprev_s[i]);
            }
        }
        prof.mark(1, T - 1 - t);
        // ---- C (shadow of the d a hand-off): recurrent tile dz_a(t+1) . W_hh for the owned units ----
        if (PT && t < T - 1) {
            const size_t toff = (size_t)(t + 1) * Bt * H4;
            auto consume = [&](int slot, int s) {
                const uint4* stg = ring + (size_t)slot * STG;
                const bool zok = lane_z && colbase + s * 16 < H4;
                const float4* zs = reinterpret_cast<const float4*>(stg + MTL * 32 + lg * 4 + lj);
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt) {
                    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (zok) {
                        cur = zs[mt * ZS];
                        if (umax_acc(0u, cur) == kCanary) cur = poll4_slow(zb[mt] + toff + s * 16, p.abort_word);      // copied before it was published
                    }
                    const uint4 b = stg[mt * 32 + lane];
                    const uint32_t bhi[2] = {b.x, b.y}, blo[2] = {b.z, b.w};
                    uint4 ahi = make_uint4(0u, 0u, 0u, 0u), alo = make_uint4(0u, 0u, 0u, 0u);
                    split2(cur.x, cur.y, ahi.x, alo.x);
                    split2(cur.z, cur.w, ahi.z, alo.z);
                    mma3(acc[mt], ahi, alo, bhi, blo);
                }
            };
            if (ring_deep) {
                // both operands arrive through the per-warp cp.async ring: the slots hold the fragments of the first RD k16 steps
                // already (prefetched during the rest of the previous step); their dz parts are requested now, after the dz(t+1) gate
                {
                    int sl = rslot;
#pragma unroll
                    for (int i = 0; i < RD; ++i) {
                        ring_fill_dz(sl, i, toff);
                        cp_async_commit();
                        if (++sl == RD) sl = 0;
                    }
                }
                for (int s = 0; s < KS; ++s) {
                    cp_async_wait<RD - 1>();
                    consume(rslot, s);
                    // refill the slot just read: k16 step s + RD of this time step (fragments + dz), or the fragments of the first k16
                    // steps of the NEXT time step (their dz does not exist yet)
                    if (s + RD < KS) ring_fill(rslot, s + RD, true, toff);
                    else ring_fill(rslot, s + RD - KS, false, 0);
                    if (++rslot == RD) rslot = 0;
                }
            } else {      // fewer k16 steps per warp than ring stages (small models): one stage at a time
                for (int s = 0; s < KS; ++s) {
                    ring_fill(0, s, true, toff);
                    cp_async_wait<0>();
                    consume(0, s);
                }
            }
        } else if (t < T - 1) {
            const size_t toff = (size_t)(t + 1) * Bt * H4;
            constexpr int PF = MSA_PF_BWD;
            float4 zv[PF][MTL][HR];
            auto issue = [&](int s, float4 (&dst)[MTL][HR]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                    for (int hr = 0; hr < HR; ++hr)
                        dst[mt][hr] = (rowok[mt][hr] && col < H4) ? ld_poll4(zrow[mt][hr] + toff + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            auto process = [&](int s, float4 (&cur)[MTL][HR]) {
                const int col = (wsl * KS + s) * 16 + 4 * lj;
                const uint2* bf = Bfrag + ((size_t)wsl * KS + s) * 2 * 32;
                const uint2 bh = bf[lane], bl = bf[32 + lane];
                const uint32_t bhi[2] = {bh.x, bh.y}, blo[2] = {bl.x, bl.y};
                unsigned int cmax = 0u;
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                    for (int hr = 0; hr < HR; ++hr) cmax = umax_acc(cmax, cur[mt][hr]);
                if (cmax == kCanary) {
#pragma unroll
                    for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
                        for (int hr = 0; hr < HR; ++hr)
                            if (rowok[mt][hr] && col < H4 && !ready4(cur[mt][hr])) cur[mt][hr] = poll4_slow(zrow[mt][hr] + toff + col, p.abort_word);
                }
#pragma unroll
                for (int mt = 0; mt < MTL; ++mt) {
                    uint4 ahi, alo;
                    split2(cur[mt][0].x, cur[mt][0].y, ahi.x, alo.x);
                    split2(cur[mt][HR - 1].x, cur[mt][HR - 1].y, ahi.y, alo.y);
                    split2(cur[mt][0].z, cur[mt][0].w, ahi.z, alo.z);
                    split2(cur[mt][HR - 1].z, cur[mt][HR - 1].w, ahi.w, alo.w);
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
        prof.mark(2, T - 1 - t);
        // ---- D: d a(t) of every position, a(t), normaliser ----
        if ((BtL & 3) == 0) {
            for (int i4 = threadIdx.x; i4 < (RL >> 2); i4 += kMT) {
                const int i = i4 * 4, g = i / BtL;
                const size_t o = g * gr.tstride + (size_t)t * BtL + (i - g * BtL);
                reinterpret_cast<float4*>(als)[i4] = __ldg(reinterpret_cast<const float4*>(p.align + o));
                reinterpret_cast<float4*>(das)[i4] = pollq4(p.dat + o, p.abort_word);
            }
        } else {
            for (int i = threadIdx.x; i < RL; i += kMT) {
                const int g = pos_task(i);
                const size_t o = g * gr.tstride + (size_t)t * BtL + (i - g * BtL);
                als[i] = __ldg(p.align + o);
                das[i] = pollq1(p.dat + o, p.abort_word);
            }
        }
        if ((int)threadIdx.x < R) zn_s[threadIdx.x] = __ldg(p.znorm + gr.task(threadIdx.x) * gr.tstride + (size_t)t * Bt + gr.brow(threadIdx.x));
        __syncthreads();
        prof.mark(3, T - 1 - t);
        // normalisation backward (softmax, or sigmoid / sum): d e(t) in place of d a(t)
        for (int r = w; r < R; r += kMW) {
            float sd = 0.f;
            for (int l = lane; l < L; l += 32) sd += als[r * L + l] * das[r * L + l];
            sd = warp_sum(sd);
            for (int l = lane; l < L; l += 32) {
                const float a = als[r * L + l];
                float de = a * (das[r * L + l] - sd);
                if (p.norm == 1) de = de * (1.f - a * zn_s[r]);
                das[r * L + l] = de;
            }
        }
        __syncthreads();
        // terms of dq for the owned attention dim; d e and dS of the owned positions
        if (has_q) {
#pragma unroll
            for (int k = 0; k < KV; ++k) {
                const int it = threadIdx.x + k * kMT;
                if (it < RL) tq_s[it] = das[it] * (1.f - sv_q[k] * sv_q[k]);
            }
            for (int it = threadIdx.x + KV * kMT; it < RL; it += kMT) {
                const float sv = __ldg(s_addr(t, it, d0));
                tq_s[it] = das[it] * (1.f - sv * sv);
            }
        }
        for (int i = threadIdx.x; i < np; i += kMT) {
            const int pp = p0 + i, g = pos_task(pp);
            p.de[g * gr.tstride + (size_t)t * BtL + (pp - g * BtL)] = das[pp];
        }
        {
            auto do_item = [&](int it, float sv) {
                const int pi = it / A, d = it - pi * A, pp = p0 + pi, g = pos_task(pp);
                const float dS = das[pp] * vs[pos_set(pp) * A + d] * (1.f - sv * sv);
                p.ds[g * gr.tstride + ((size_t)t * BtL + (pp - g * BtL)) * A + d] = dS;
                ds_s[it] = dS;
            };
#pragma unroll
            for (int k = 0; k < KV; ++k) {
                const int it = threadIdx.x + k * kMT;
                if (it < np * A) do_item(it, sv_own[k]);
            }
            for (int it = threadIdx.x + KV * kMT; it < np * A; it += kMT) do_item(it, __ldg(s_addr(t, p0 + it / A, it % A)));
        }
        __syncthreads();
        prof.mark(4, T - 1 - t);
        // dq(t)[r][d0] = v[d0] * sum_l tq[r][l]  (one warp per row); publishes dq
        if (has_q)
            for (int r = w; r < R; r += kMW) {
                float a = 0.f;
                for (int l = lane; l < L; l += 32) a += tq_s[r * L + l];
                a = warp_sum(a);
                if (lane == 0) st_pub(p.dq + gr.task(r) * gr.tstride + ((size_t)t * Bt + gr.brow(r)) * A + d0, vq_s[PT ? gr.task(r) : 0] * a);
            }
        // d(conv features)[pi][f] = sum_d dS[pi][d] wld[d][f]: one thread per (position, filter) -- a warp is the F filters of one
        // position (F = 32): dS is a broadcast read, the weights are conflict-free; four independent accumulators, fixed order
        for (int it = threadIdx.x; it < np * F; it += kMT) {
            const int pi = it / F, f = it - pi * F, pp = p0 + pi, g = pos_task(pp);
            const float* dsr = ds_s + (size_t)pi * A;
            const float* wd = wld_s + pos_set(pp) * WDS + f;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            int d = 0;
            for (; d + 3 < A; d += 4) {
                a0 += dsr[d] * wd[(size_t)d * F];
                a1 += dsr[d + 1] * wd[(size_t)(d + 1) * F];
                a2 += dsr[d + 2] * wd[(size_t)(d + 2) * F];
                a3 += dsr[d + 3] * wd[(size_t)(d + 3) * F];
            }
            for (; d < A; ++d) a0 += dsr[d] * wd[(size_t)d * F];
            st_pub(p.dconvf + g * gr.tstride + ((size_t)t * BtL + (pp - g * BtL)) * F + f, (a0 + a1) + (a2 + a3));
        }
        prof.mark(5, T - 1 - t);
        __syncthreads();      // tq is done: the region becomes the partial-tile buffer
        // ---- F: dq(t) arrives: dq(t) . W_q for the owned units into the same accumulators, then one reduction ----
        for (int s = w; s < KQ; s += kMW) {
            const int col = s * 16 + 4 * lj;
            const size_t toff = (size_t)t * Bt * A;
#pragma unroll
            for (int mt = 0; mt < MTL; ++mt) {
                const uint2* bqf = Bq + ((size_t)(PT ? mt : 0) * KQ + s) * 2 * 32;
                const uint2 bh = bqf[lane], bl = bqf[32 + lane];
                const uint32_t bhi[2] = {bh.x, bh.y}, blo[2] = {bl.x, bl.y};
                float4 cur[HR];
#pragma unroll
                for (int hr = 0; hr < HR; ++hr) {
                    const int r = PT ? mt * Bt + lg : mt * 16 + hr * 8 + lg;
                    const bool ok = PT ? (lg < Bt && mt < p.G) : r < R;
                    const int rr = ok ? r : 0;
                    const float* qrow = p.dq + gr.task(rr) * gr.tstride + (size_t)gr.brow(rr) * A;
                    cur[hr] = (ok && col < A) ? pollq4(qrow + toff + col, p.abort_word) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                uint4 ahi = make_uint4(0u, 0u, 0u, 0u), alo = make_uint4(0u, 0u, 0u, 0u);
                split2(cur[0].x, cur[0].y, ahi.x, alo.x);
                split2(cur[0].z, cur[0].w, ahi.z, alo.z);
                if (!PT) {
                    split2(cur[HR - 1].x, cur[HR - 1].y, ahi.y, alo.y);
                    split2(cur[HR - 1].z, cur[HR - 1].w, ahi.w, alo.w);
                }
                mma3(acc[mt], ahi, alo, bhi, blo);
            }
        }
#pragma unroll
        for (int mt = 0; mt < MTL; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) part[(((size_t)w * MTL + mt) * 4 + i) * 32 + lane] = acc[mt][i];
        __syncthreads();
        prof.mark(6, T - 1 - t);
        if (pw) {
            float drec = 0.f;
            const float* pp_ = part + ((size_t)c_mt * 4 + c_reg) * 32 + c_ln;
#pragma unroll
            for (int ww = 0; ww < kMW; ++ww) drec += pp_[(size_t)ww * MTL * 4 * 32];
            const size_t zb = ((size_t)t * Bt + c_b) * H4;
            float dh = dhe + drec;
            if (mask_c) dh = mk ? dh * p.drop_scale : 0.f;
            const LstmGrad g = lstm_point_bwd(gi[0], gi[1], gi[2], gi[3], cc, cp, dh, dcarry);
            dcarry = g.dc_prev;
            st_pub(dza_c + zb + 0 * (size_t)Ha + c_u, g.di);
            st_pub(dza_c + zb + 1 * (size_t)Ha + c_u, g.df);
            st_pub(dza_c + zb + 2 * (size_t)Ha + c_u, g.dg);
            st_pub(dza_c + zb + 3 * (size_t)Ha + c_u, g.do_);
        }
        if (t > 0) fetch(t - 1);
        // ---- H: d(conv features) of the +-pad neighbours; location-conv backward for the owned positions ----
        for (int ro = 0; ro < nrown; ++ro) {
            const int r = r_lo + ro, g = r / Bt, lo0 = wlo(ro), nwin = whi(ro) - lo0 + 1;
            const float* src = p.dconvf + g * gr.tstride + ((size_t)t * BtL + (size_t)(r - g * Bt) * L + lo0) * F;
            for (int it = threadIdx.x; it < nwin * F; it += kMT) {
                const int wl = it / F, f = it - wl * F;
                dcw[((size_t)ro * WIN + wl) * FP + f] = pollq1(src + it, p.abort_word);
            }
        }
        __syncthreads();
        prof.mark(7, T - 1 - t);
        // one warp per owned position (input position of the conv), lanes over the taps: output position lo = l - k + pad
        for (int i = w; i < np; i += kMW) {
            const int pp = p0 + i, r = pp / L, l = pp - r * L, ro = r - r_lo, lo0 = wlo(ro);
            const float* wl_ = wloc_s + pos_set(pp) * WLS;
            float a0 = 0.f, a1 = 0.f;
            for (int k = lane; k < Kl; k += 32) {
                const int lo = l - k + pl;
                if (lo >= 0 && lo < L) {
                    const float* dv = dcw + ((size_t)ro * WIN + (lo - lo0)) * FP;
                    for (int f = 0; f < F; ++f) {
                        a0 += wl_[f * CKP + k] * dv[f];
                        a1 += wl_[f * CKP + Kl + k] * dv[f];
                    }
                }
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
            if (lane == 0) {
                dprev_s[i] = a0;
                gcum_s[i] += a1;
            }
        }
        __syncthreads();
    }
    if (PT) cp_async_wait<0>();      // the ring runs RD stages ahead of the last step
}

static bool attn_bwd_mma_ok(const msa_config& cfg, int R, int L, int sm_count, size_t smem_limit) {
    const int Ha = cfg.attn_rnn_dim, A = cfg.attn_dim;
    if (cfg.forward_attn || R < 1 || R > 32 || Ha % 4 != 0 || A % 4 != 0 || cfg.loc_filters > 32) return false;
    if (A > sm_count || (Ha + sm_count - 1) / sm_count > 8) return false;
    return attn_bwd_mma_layout(R, L, Ha, A, cfg.loc_filters, cfg.loc_kernel, sm_count, (R + 15) / 16).total + 256 <= smem_limit;
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
    (void)T;
    return attn_fwd_mma_ok(cfg, G * B, L, sm_count, smem_limit);
}
// per-task weights: what one launch can carry
static bool attn_pt_ok(const msa_config& cfg, int G, int B, int L, int sm_count, size_t smem_limit) {
    const int Ha = cfg.attn_rnn_dim, A = cfg.attn_dim, R = G * B;
    if (G < 2 || G > kPtGroupMax || B < 1 || B > 8) return false;
    if (!attn_fwd_mma_ok(cfg, R, L, sm_count, smem_limit) || !attn_bwd_mma_ok(cfg, R, L, sm_count, smem_limit)) return false;
    if ((R * L + sm_count - 1) / sm_count > L) return false;      // the owned positions of a CTA touch at most two tasks
    return attn_fwd_mma_layout(R, L, Ha, A, cfg.loc_filters, cfg.loc_kernel, sm_count, G, 1).total + 256 <= smem_limit &&
           attn_bwd_mma_layout(R, L, Ha, A, cfg.loc_filters, cfg.loc_kernel, sm_count, G, 1).total + 256 <= smem_limit;
}
bool attn_chain_pt_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit) {
    (void)T;
    return attn_pt_ok(cfg, G, B, L, sm_count, smem_limit);
}
size_t attn_chain_pt_frag_bytes(const msa_config& cfg, int sm_count) {
    const int Ha = cfg.attn_rnn_dim;
    const size_t fwd = (size_t)sm_count * kMW * ((Ha + 255) / 256) * 128 * sizeof(uint4);
    const size_t bwd = (size_t)sm_count * kMW * ((4 * Ha + 255) / 256) * 32 * sizeof(uint4);
    return fwd > bwd ? fwd : bwd;
}

int launch_attn_chain_fwd_mma(const AttnChainParams& p0, int sm_count, size_t smem_limit, cudaStream_t st) {
    AttnChainParams p = p0;
    if (p.G < 1) p.G = 1;
    const int R = p.G * p.B, NT = p.pt ? p.G : (R + 7) / 8;
    MSA_CHECK(!p.fa && R <= 32 && p.A <= sm_count && (p.Ha + sm_count - 1) / sm_count <= 7, MSA_E_UNSUPPORTED,
              "attn_chain_fwd_mma: configuration outside the grouped kernel (rows %d, attention dim %d)", R, p.A);
    const AttnFwdMmaLay lay = attn_fwd_mma_layout(R, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, NT, p.pt);
    const size_t smem = lay.total;
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "attn_chain_fwd_mma: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    if (p.pt) {
        MSA_CHECK(p.G >= 2 && p.G <= kPtGroupMax && p.B <= 8 && lay.NPmax <= p.L && p.wfrag, MSA_E_UNSUPPORTED,
                  "attn_chain_fwd_mma: per-task weights need 2..%d tasks of <= 8 rows (G=%d, B=%d)", kPtGroupMax, p.G, p.B);
        FragSrc src{};
        for (int g = 0; g < p.G; ++g) { src.whh[g] = p.whh_g[g]; src.wq[g] = p.wq_g[g]; }
        k_attn_frag_fwd<<<dim3(sm_count, p.G), kMT, 0, st>>>(src, p.Ha, p.A, p.wfrag_stride, const_cast<uint4*>(p.wfrag));
        MSA_CUDA(cudaGetLastError());
        count_launch();
    }
    const size_t TB = (size_t)p.T * p.B;
    for (int g = 0; g < p.G; ++g) {       // canaries of the three hand-off arrays of every task (common.cuh)
        MSA_TRY(k_fill_canary(p.ha + g * p.tstride, (int64_t)(TB * p.Ha), st));
        MSA_TRY(k_fill_canary(p.q + g * p.tstride, (int64_t)(TB * p.A), st));
        MSA_TRY(k_fill_canary(p.e + g * p.tstride, (int64_t)(TB * p.L), st));
    }
    if (p.pt) {
        switch (NT) {
            case 2: return coop_launch(k_attn_fwd_mma<2, true>, p, sm_count, smem, st);
            case 3: return coop_launch(k_attn_fwd_mma<3, true>, p, sm_count, smem, st);
            default: return coop_launch(k_attn_fwd_mma<4, true>, p, sm_count, smem, st);
        }
    }
    switch (NT) {
        case 1: return coop_launch(k_attn_fwd_mma<1, false>, p, sm_count, smem, st);
        case 2: return coop_launch(k_attn_fwd_mma<2, false>, p, sm_count, smem, st);
        case 3: return coop_launch(k_attn_fwd_mma<3, false>, p, sm_count, smem, st);
        default: return coop_launch(k_attn_fwd_mma<4, false>, p, sm_count, smem, st);
    }
}
bool attn_chain_bwd_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit) {
    (void)T;
    return attn_bwd_mma_ok(cfg, G * B, L, sm_count, smem_limit);
}
int launch_attn_chain_bwd_mma(const AttnChainBwdParams& p0, int sm_count, size_t smem_limit, cudaStream_t st) {
    AttnChainBwdParams p = p0;
    if (p.G < 1) p.G = 1;
    const int R = p.G * p.B, MTL = p.pt ? p.G : (R + 15) / 16;
    MSA_CHECK(!p.fa && R <= 32 && p.A <= sm_count && p.A % 4 == 0 && (p.Ha + sm_count - 1) / sm_count <= 8, MSA_E_UNSUPPORTED,
              "attn_chain_bwd_mma: configuration outside the grouped kernel (rows %d, attention dim %d)", R, p.A);
    const AttnBwdMmaLay lay = attn_bwd_mma_layout(R, p.L, p.Ha, p.A, p.F, p.Kl, sm_count, MTL, p.pt);
    const size_t smem = lay.total;
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "attn_chain_bwd_mma: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    if (p.pt) {
        MSA_CHECK(p.G >= 2 && p.G <= kPtGroupMax && p.B <= 8 && lay.NPmax <= p.L && p.wfrag, MSA_E_UNSUPPORTED,
                  "attn_chain_bwd_mma: per-task weights need 2..%d tasks of <= 8 rows (G=%d, B=%d)", kPtGroupMax, p.G, p.B);
        FragSrc src{};
        for (int g = 0; g < p.G; ++g) src.whh[g] = p.whh_g[g];
        k_attn_frag_bwd<<<dim3(sm_count, p.G), kMT, 0, st>>>(src, p.Ha, p.wfrag_stride, const_cast<uint4*>(p.wfrag));
        MSA_CUDA(cudaGetLastError());
        count_launch();
    }
    const size_t TB = (size_t)p.T * p.B;
    for (int g = 0; g < p.G; ++g) {       // canaries of the four hand-off arrays of every task (common.cuh)
        MSA_TRY(k_fill_canary(p.dza + g * p.tstride, (int64_t)(TB * 4 * p.Ha), st));
        MSA_TRY(k_fill_canary(p.dq + g * p.tstride, (int64_t)(TB * p.A), st));
        MSA_TRY(k_fill_canary(p.dat + g * p.tstride, (int64_t)(TB * p.L), st));
        MSA_TRY(k_fill_canary(p.dconvf + g * p.tstride, (int64_t)(TB * p.L * p.F), st));
    }
    if (p.pt) {
        switch (MTL) {
            case 2: return coop_launch(k_attn_bwd_mma<2, true>, p, sm_count, smem, st);
            case 3: return coop_launch(k_attn_bwd_mma<3, true>, p, sm_count, smem, st);
            default: return coop_launch(k_attn_bwd_mma<4, true>, p, sm_count, smem, st);
        }
    }
    if (MTL == 1) return coop_launch(k_attn_bwd_mma<1, false>, p, sm_count, smem, st);
    return coop_launch(k_attn_bwd_mma<2, false>, p, sm_count, smem, st);
}

}  // namespace msa
