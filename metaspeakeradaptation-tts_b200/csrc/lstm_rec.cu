// Persistent LSTM recurrence (forward and backward over all time steps in ONE cooperative launch).
//
// Used for (a) the encoder BiLSTM over packed lengths (modules_tacotron2nv/encoder.py:30-32,42-50;
// both directions run concurrently on disjoint halves of the grid) and (b) the decoder-RNN chain
// h_d(t) = LSTMCell(W_ih.[h_a(t); ctx(t)] + W_hh.h_d(t-1)) with recurrent dropout
// (modules_tacotron2nv/decoder.py:135-137,262-265), whose input projection is hoisted out of the
// loop as one GEMM because it does not depend on h_d.
// W_hh is sliced by hidden unit across the CTAs and stays resident in shared memory in fp32.
#include "rec_common.cuh"
#include "kernels.h"

namespace msa {

// Per-step hand-off: h(t) (forward) / dz(t) (backward) are published with plain stores into the stash arrays the later
// GEMMs read anyway, and consumed by canary polling (common.cuh) -- no grid barrier, no atomics.  The streaming
// per-step inputs of the point-wise update (zin, gates, c, dh_ext, mask) are fetched one step ahead so that their
// DRAM/L2 latency overlaps the wait for the other CTAs.
struct LstmFwdSmem {
    size_t wsm, hs, part, total;
    int KP, BP;
};
__host__ __device__ inline LstmFwdSmem lstm_fwd_layout(int B, int H) {
    LstmFwdSmem s;
    s.KP = round_up_i(H, 128);
    s.BP = (B + 3) & ~3;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.wsm = take((size_t)4 * kUMax * s.KP);
    s.hs = take((size_t)s.BP * s.KP);
    s.part = take((size_t)kBTiles * kRecWarps * 32);
    s.total = o;
    return s;
}

template <bool kProf, int NT>
__global__ void __launch_bounds__(NT, 1) k_lstm_rec_fwd(LstmRecParams p) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, B = p.B, T = p.T, H4 = 4 * H;
    const int ncta_dir = gridDim.x / p.ndir;
    const int dir = blockIdx.x / ncta_dir;
    const bool idle = dir >= p.ndir;
    const int ci = blockIdx.x % ncta_dir;
    const int u0 = idle ? 0 : part_lo(ci, H, ncta_dir), u1 = idle ? 0 : part_lo(ci + 1, H, ncta_dir);
    const int U = u1 - u0, R = 4 * U, RG = (R + 7) >> 3;
    if (U == 0) return;   // nothing to own: this CTA neither produces nor consumes

    const LstmFwdSmem lay = lstm_fwd_layout(B, H);
    const int KP = lay.KP, KP4 = KP >> 2;
    float* Wsm = smem + lay.wsm;                     // [RG*8][KP]  local row rl = ul*4 + gate, zero-padded
    float* hs = smem + lay.hs;                       // [BP][KP]    zero-padded
    float* part = smem + lay.part;

    const float* whh = p.whh + (size_t)dir * p.whh_dir_stride;
    const float* zin = p.zin + (size_t)dir * T * B * H4;
    float* hout = p.hout + (size_t)dir * T * B * H;
    float* cout = p.cout + (size_t)dir * T * B * H;
    float* gates = p.gates + (size_t)dir * T * B * H4;

    for (int idx = threadIdx.x; idx < RG * 8 * KP4; idx += NT) {
        const int rl = idx / KP4, k4 = idx % KP4;
        const int g = rl & 3, ul = rl >> 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < R && k4 < (H >> 2)) v = __ldg(reinterpret_cast<const float4*>(whh + (size_t)(g * H + u0 + ul) * H) + k4);
        reinterpret_cast<float4*>(Wsm)[idx] = v;
    }
    for (int idx = threadIdx.x; idx < lay.BP * KP; idx += NT) hs[idx] = 0.f;

    // point-wise role: 4 lanes per cell (ul, b), lane g evaluates gate g; the cell state lives in lane 0's register
    const bool pw = (int)threadIdx.x < 4 * U * B;
    const int g = threadIdx.x & 3, cell = threadIdx.x >> 2;
    const int ul = pw ? cell / B : 0, b = pw ? cell % B : 0, u = u0 + ul;
    const int len = (pw && p.lengths) ? (int)p.lengths[b] : T;
    float zi = 0.f, cstate = 0.f;
    unsigned char mk = 1;
    auto fetch = [&](int t) {
        zi = __ldg(zin + ((size_t)t * B + b) * H4 + (size_t)g * H + u);
        if (p.mask && g == 0) mk = p.mask[((size_t)t * B + b) * H + u];
    };
    if (pw) fetch(dir == 0 ? 0 : T - 1);
    SpinGuard sg(p.abort_word);
    ChainProf<kProf> prof;
    prof.start(p.prof, p.trace, p.trace_t0);
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? s : T - 1 - s;
        const int tp = (dir == 0) ? t - 1 : t + 1;
        if (s > 0 && (p.flags & kFlagGate)) {
            // sentinel of producer CTA c: h of its last cell (unit u1(c)-1, row B-1), the last word it publishes
            const float* hrow = hout + ((size_t)tp * B + (B - 1)) * H;
            gate_wait(ncta_dir, [&](int c) { const int e = part_lo(c + 1, H, ncta_dir); return e > 0 ? hrow + e - 1 : nullptr; }, sg);
            __syncthreads();
        }
        prof.mark(0, s);
        // h(t-1) goes from global memory (L2) straight into the mat-vec registers, canary-checked; at s == 0 it is zero
        if (s > 0) cta_matvec_fwd<NT, true>(Wsm, RG, KP, hout + (size_t)tp * B * H, B, part, 0, KP >> 7, H, &sg);
        else for (int i = threadIdx.x; i < kBTiles * kRecWarps * 32; i += NT) part[i] = 0.f;
        __syncthreads();
        prof.mark(1, s);
        if (pw) {
            const size_t zb = ((size_t)t * B + b) * H4;
            const bool active = t < len;
            const float z = lstm_gate_sum<NT>(part, RG, ul * 4 + g, b) + zi;
            float act = g == 2 ? fast_tanh(z) : fast_sigmoid(z);
            if (!active) act = 0.f;
            const unsigned int gm = 0xFu << (threadIdx.x & 28);                 // the 4 lanes of this cell
            const float ai = __shfl_sync(gm, act, 0, 4), af = __shfl_sync(gm, act, 1, 4);
            const float ag = __shfl_sync(gm, act, 2, 4), ao = __shfl_sync(gm, act, 3, 4);
            float cn = 0.f;
            if (g == 0) {
                float hv = 0.f;
                if (active) {
                    cn = af * cstate + ai * ag;
                    cstate = cn;
                    hv = ao * fast_tanh(cn);
                    if (p.mask) hv = mk ? hv * p.drop_scale : 0.f;
                }
                st_pub(hout + ((size_t)t * B + b) * H + u, hv);      // consumed by every CTA at the next step: first memory op
                if (p.flags & 8) __threadfence();
            }
            gates[zb + (size_t)g * H + u] = act;                      // stash for the backward pass: off the critical path
            if (g == 0) cout[((size_t)t * B + b) * H + u] = cn;
            if (s + 1 < T) fetch(dir == 0 ? t + 1 : t - 1);
        }
        prof.mark(2, s);
    }
}

template <bool kProf, int NT>
__global__ void __launch_bounds__(NT, 1) k_lstm_rec_bwd(LstmRecBwdParams p) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, B = p.B, T = p.T, H4 = 4 * H;
    const int BP = (B + 3) & ~3;
    const int ncta_dir = gridDim.x / p.ndir;
    const int dir = blockIdx.x / ncta_dir;
    const bool idle = dir >= p.ndir;
    const int ci = blockIdx.x % ncta_dir;
    const int u0 = idle ? 0 : part_lo(ci, H, ncta_dir), u1 = idle ? 0 : part_lo(ci + 1, H, ncta_dir);
    const int U = u1 - u0;
    if (U == 0) return;

    float* WT = smem;                                // [kUMax][4H]   WT[ul][r] = W_hh[r][u0+ul], zero rows for ul >= U
    float* part = WT + (size_t)kUMax * H4;           // [warps*32]
    float* dhs = part + NW * 32;              // [kUMax][BP]

    const float* whh = p.whh + (size_t)dir * p.whh_dir_stride;
    const float* gates = p.gates + (size_t)dir * T * B * H4;
    const float* cst = p.cout + (size_t)dir * T * B * H;
    const float* dh_ext = p.dh_ext + (size_t)dir * T * B * H;
    float* dz = p.dz + (size_t)dir * T * B * H4;

    for (int idx = threadIdx.x; idx < kUMax * H4; idx += NT) {
        const int ul = idx / H4, r = idx % H4;
        WT[idx] = ul < U ? __ldg(whh + (size_t)r * H + u0 + ul) : 0.f;
    }
    for (int idx = threadIdx.x; idx < kUMax * BP; idx += NT) dhs[idx] = 0.f;

    const bool pw = (int)threadIdx.x < U * B;
    const int ul = pw ? threadIdx.x / B : 0, b = pw ? threadIdx.x % B : 0, u = u0 + ul;
    const int len = (pw && p.lengths) ? (int)p.lengths[b] : T;
    float gi[4] = {0.f, 0.f, 0.f, 0.f}, cc = 0.f, cp = 0.f, dhe = 0.f, dcarry = 0.f;
    unsigned char mk = 1;
    auto fetch = [&](int t) {   // stash of forward step t + the external gradient of h(t)
        const int tp = (dir == 0) ? t - 1 : t + 1;
        const size_t zb = ((size_t)t * B + b) * H4, hb = ((size_t)t * B + b) * H + u;
#pragma unroll
        for (int g = 0; g < 4; ++g) gi[g] = __ldg(gates + zb + (size_t)g * H + u);
        cc = __ldg(cst + hb);
        cp = (tp >= 0 && tp < T) ? __ldg(cst + ((size_t)tp * B + b) * H + u) : 0.f;
        dhe = __ldg(dh_ext + hb);
        if (p.mask) mk = p.mask[hb];
    };
    if (pw) fetch(dir == 0 ? T - 1 : 0);
    SpinGuard sg(p.abort_word);
    ChainProf<kProf> prof;
    prof.start(p.prof, p.trace, p.trace_t0);
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? T - 1 - s : s;     // reverse of the forward processing order
        const int tn = (dir == 0) ? t + 1 : t - 1;    // step processed just before (forward-order successor)
        if (s > 0 && (p.flags & kFlagGate)) {
            // sentinel of producer CTA c: the output-gate gradient of its last cell, the last word it publishes
            const float* zrow = dz + ((size_t)tn * B + (B - 1)) * H4 + (size_t)3 * H;
            gate_wait(ncta_dir, [&](int c) { const int e = part_lo(c + 1, H, ncta_dir); return e > 0 ? zrow + e - 1 : nullptr; }, sg);
            __syncthreads();
        }
        if (s > 0) cta_matvec_bwd<NT>(WT, H4, dz + (size_t)tn * B * H4, B, part, dhs, BP, sg);
        prof.mark(0, s);
        if (pw) {
            const size_t zb = ((size_t)t * B + b) * H4;
            const bool active = t < len;
            LstmGrad g = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (active) {
                float dh = dhe + (s > 0 ? dhs[ul * BP + b] : 0.f);
                if (p.mask) dh = mk ? dh * p.drop_scale : 0.f;
                g = lstm_point_bwd(gi[0], gi[1], gi[2], gi[3], cc, cp, dh, dcarry);
                dcarry = g.dc_prev;
            }
            st_pub(dz + zb + 0 * H + u, g.di);
            st_pub(dz + zb + 1 * H + u, g.df);
            st_pub(dz + zb + 2 * H + u, g.dg);
            st_pub(dz + zb + 3 * H + u, g.do_);
            if (s + 1 < T) fetch(dir == 0 ? t - 1 : t + 1);
        }
        prof.mark(1, s);
    }
}

static int rec_check(int B, int H, int ndir, int sm_count) {
    MSA_CHECK(H % 4 == 0, MSA_E_UNSUPPORTED, "lstm_rec: hidden size %d must be a multiple of 4", H);
    MSA_CHECK(B >= 1 && B <= kBMax, MSA_E_UNSUPPORTED, "lstm_rec: batch %d outside [1,%d]", B, kBMax);
    const int ncta_dir = sm_count / ndir;
    const int umax = ncta_dir >= 1 ? (H + ncta_dir - 1) / ncta_dir : kUMax + 1;
    MSA_CHECK(umax <= kUMax, MSA_E_UNSUPPORTED, "lstm_rec: hidden size %d needs more than %d units per CTA on %d CTAs", H, kUMax, ncta_dir);
    MSA_CHECK(4 * umax * B <= kRecThreads, MSA_E_UNSUPPORTED, "lstm_rec: %d units x %d batch rows per CTA exceed the point-wise thread budget", umax, B);
    return 0;
}

bool lstm_rec_single_ok(int B, int H, int ndir, int sm_count, size_t smem_limit) {
    if (H % 4 != 0 || B < 1 || B > kBMax) return false;
    const int ncta_dir = sm_count / ndir;
    if (ncta_dir < 1) return false;
    const int umax = (H + ncta_dir - 1) / ncta_dir;
    if (umax > kUMax || 4 * umax * B > kRecThreads) return false;
    return sizeof(float) * lstm_fwd_layout(B, H).total <= smem_limit &&
           sizeof(float) * ((size_t)kUMax * 4 * H + kRecWarps * 32 + kUMax * ((B + 3) & ~3)) <= smem_limit;
}
size_t lstm_rec_fwd_smem(int B, int H) { return sizeof(float) * lstm_fwd_layout(B, H).total; }
size_t lstm_rec_bwd_smem(int B, int H) {
    const int BP = (B + 3) & ~3;
    return sizeof(float) * ((size_t)kUMax * 4 * H + kRecWarps * 32 + kUMax * BP);
}

int launch_lstm_rec_fwd(const LstmRecParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(rec_check(p.B, p.H, p.ndir, sm_count));
    const size_t smem = lstm_rec_fwd_smem(p.B, p.H);
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "lstm_rec_fwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    MSA_CUDA(cudaFuncSetAttribute(p.prof ? k_lstm_rec_fwd<true, kRecThreads> : k_lstm_rec_fwd<false, kRecThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSA_TRY(k_fill_canary(p.hout, (int64_t)((size_t)p.ndir * p.T * p.B * p.H), st));   // canaries (common.cuh)
    LstmRecParams pp = p;
    void* args[] = {&pp};
    MSA_CUDA(cudaLaunchCooperativeKernel(pp.prof ? (void*)k_lstm_rec_fwd<true, kRecThreads> : (void*)k_lstm_rec_fwd<false, kRecThreads>, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

int launch_lstm_rec_bwd(const LstmRecBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(rec_check(p.B, p.H, p.ndir, sm_count));
    const size_t smem = lstm_rec_bwd_smem(p.B, p.H);
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "lstm_rec_bwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    MSA_CUDA(cudaFuncSetAttribute(p.prof ? k_lstm_rec_bwd<true, kRecThreads> : k_lstm_rec_bwd<false, kRecThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSA_TRY(k_fill_canary(p.dz, (int64_t)((size_t)p.ndir * p.T * p.B * 4 * p.H), st));  // canaries (common.cuh)
    LstmRecBwdParams pp = p;
    void* args[] = {&pp};
    MSA_CUDA(cudaLaunchCooperativeKernel(pp.prof ? (void*)k_lstm_rec_bwd<true, kRecThreads> : (void*)k_lstm_rec_bwd<false, kRecThreads>, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

}  // namespace msa
