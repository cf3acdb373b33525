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

__global__ void __launch_bounds__(kRecThreads, 1) k_lstm_rec_fwd(LstmRecParams p) {
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, B = p.B, T = p.T, H4 = 4 * H;
    const int BP = (B + 3) & ~3;
    const int ncta_dir = gridDim.x / p.ndir;
    const int dir = blockIdx.x / ncta_dir;
    const bool idle = dir >= p.ndir;
    const int ci = blockIdx.x % ncta_dir;
    const int u0 = idle ? 0 : part_lo(ci, H, ncta_dir), u1 = idle ? 0 : part_lo(ci + 1, H, ncta_dir);
    const int U = u1 - u0, R = 4 * U;

    float* Wsm = smem;                               // [4*kUMax][H]
    float* hs = Wsm + (size_t)4 * kUMax * H;         // [BP][H]
    float* part = hs + (size_t)BP * H;               // [warps*32]
    float* zs = part + kRecWarps * 32;               // [32][BP]
    float* cs = zs + 32 * BP;                        // [kUMax][BP]

    const int dd = idle ? 0 : dir;
    const float* whh = p.whh + (size_t)dd * p.whh_dir_stride;
    const float* zin = p.zin + (size_t)dd * T * B * H4;
    float* hout = p.hout + (size_t)dd * T * B * H;
    float* cout = p.cout + (size_t)dd * T * B * H;
    float* gates = p.gates + (size_t)dd * T * B * H4;

    for (int idx = threadIdx.x; idx < R * (H >> 2); idx += kRecThreads) {
        const int rl = idx / (H >> 2), k4 = idx % (H >> 2);
        const int g = rl & 3, ul = rl >> 2;
        reinterpret_cast<float4*>(Wsm)[(size_t)rl * (H >> 2) + k4] =
            __ldg(reinterpret_cast<const float4*>(whh + (size_t)(g * H + u0 + ul) * H) + k4);
    }
    for (int idx = threadIdx.x; idx < kUMax * BP; idx += kRecThreads) cs[idx] = 0.f;
    GridBarrier gb;
    gb.init(p.barrier);
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? s : T - 1 - s;
        const int tp = (dir == 0) ? t - 1 : t + 1;
        if (U > 0) {
            for (int idx = threadIdx.x; idx < B * (H >> 2); idx += kRecThreads) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s > 0) v = ld_cg4(hout + (size_t)tp * B * H + (size_t)idx * 4);
                reinterpret_cast<float4*>(hs)[idx] = v;
            }
            __syncthreads();
            cta_matvec_fwd<false>(Wsm, R, H, hs, nullptr, 0, 0, 0, 0, nullptr, B, part, zs, BP);
            if ((int)threadIdx.x < U * B) {
                const int ul = threadIdx.x / B, b = threadIdx.x % B, u = u0 + ul;
                const size_t zb = ((size_t)t * B + b) * H4;
                const bool active = p.lengths ? (t < (int)p.lengths[b]) : true;
                float hv = 0.f;
                LstmPoint r = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (active) {
                    r = lstm_point_fwd(zs[(ul * 4 + 0) * BP + b] + __ldg(zin + zb + 0 * H + u),
                                       zs[(ul * 4 + 1) * BP + b] + __ldg(zin + zb + 1 * H + u),
                                       zs[(ul * 4 + 2) * BP + b] + __ldg(zin + zb + 2 * H + u),
                                       zs[(ul * 4 + 3) * BP + b] + __ldg(zin + zb + 3 * H + u), cs[ul * BP + b]);
                    cs[ul * BP + b] = r.c;
                    hv = r.h;
                    if (p.mask) hv = p.mask[((size_t)t * B + b) * H + u] ? hv * p.drop_scale : 0.f;
                }
                gates[zb + 0 * H + u] = r.i;
                gates[zb + 1 * H + u] = r.f;
                gates[zb + 2 * H + u] = r.g;
                gates[zb + 3 * H + u] = r.o;
                cout[((size_t)t * B + b) * H + u] = r.c;
                hout[((size_t)t * B + b) * H + u] = hv;
            }
        }
        if (s + 1 < T) gb.sync();
    }
}

__global__ void __launch_bounds__(kRecThreads, 1) k_lstm_rec_bwd(LstmRecBwdParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[33];
    const int H = p.H, B = p.B, T = p.T, H4 = 4 * H;
    const int BP = (B + 3) & ~3;
    const int ncta_dir = gridDim.x / p.ndir;
    const int dir = blockIdx.x / ncta_dir;
    const bool idle = dir >= p.ndir;
    const int ci = blockIdx.x % ncta_dir;
    const int u0 = idle ? 0 : part_lo(ci, H, ncta_dir), u1 = idle ? 0 : part_lo(ci + 1, H, ncta_dir);
    const int U = u1 - u0;

    float* WT = smem;                                // [kUMax][4H]   WT[ul][r] = W_hh[r][u0+ul]
    float* part = WT + (size_t)kUMax * H4;           // [warps*32]
    float* dhs = part + kRecWarps * 32;              // [kUMax][BP]
    float* dcs = dhs + kUMax * BP;                   // [kUMax][BP]

    const int dd = idle ? 0 : dir;
    const float* whh = p.whh + (size_t)dd * p.whh_dir_stride;
    const float* gates = p.gates + (size_t)dd * T * B * H4;
    const float* cst = p.cout + (size_t)dd * T * B * H;
    const float* dh_ext = p.dh_ext + (size_t)dd * T * B * H;
    float* dz = p.dz + (size_t)dd * T * B * H4;

    for (int idx = threadIdx.x; idx < U * H4; idx += kRecThreads) {
        const int ul = idx / H4, r = idx % H4;
        WT[idx] = __ldg(whh + (size_t)r * H + u0 + ul);
    }
    for (int idx = threadIdx.x; idx < kUMax * BP; idx += kRecThreads) { dcs[idx] = 0.f; dhs[idx] = 0.f; }
    GridBarrier gb;
    gb.init(p.barrier);
    __syncthreads();

    for (int s = 0; s < T; ++s) {
        const int t = (dir == 0) ? T - 1 - s : s;     // reverse of the forward processing order
        const int tn = (dir == 0) ? t + 1 : t - 1;    // step processed just before (forward-order successor)
        const int tp = (dir == 0) ? t - 1 : t + 1;    // forward-order predecessor (source of c_prev)
        if (U > 0) {
            if (s > 0) {
                cta_matvec_bwd(WT, U, H4, dz + (size_t)tn * B * H4, B, part, dhs, BP, nullptr, 0, 0, 1, nullptr, red);
            }
            if ((int)threadIdx.x < U * B) {
                const int ul = threadIdx.x / B, b = threadIdx.x % B, u = u0 + ul;
                const size_t zb = ((size_t)t * B + b) * H4;
                const bool active = p.lengths ? (t < (int)p.lengths[b]) : true;
                LstmGrad g = {0.f, 0.f, 0.f, 0.f, 0.f};
                if (active) {
                    float dh = __ldg(dh_ext + ((size_t)t * B + b) * H + u) + (s > 0 ? dhs[ul * BP + b] : 0.f);
                    if (p.mask) dh = p.mask[((size_t)t * B + b) * H + u] ? dh * p.drop_scale : 0.f;
                    const bool has_prev = (tp >= 0 && tp < T);
                    const float cprev = has_prev ? cst[((size_t)tp * B + b) * H + u] : 0.f;
                    g = lstm_point_bwd(__ldg(gates + zb + 0 * H + u), __ldg(gates + zb + 1 * H + u),
                                       __ldg(gates + zb + 2 * H + u), __ldg(gates + zb + 3 * H + u),
                                       cst[((size_t)t * B + b) * H + u], cprev, dh, dcs[ul * BP + b]);
                    dcs[ul * BP + b] = g.dc_prev;
                }
                dz[zb + 0 * H + u] = g.di;
                dz[zb + 1 * H + u] = g.df;
                dz[zb + 2 * H + u] = g.dg;
                dz[zb + 3 * H + u] = g.do_;
            }
        }
        if (s + 1 < T) gb.sync();
    }
}

static int rec_check(int B, int H, int ndir, int sm_count) {
    MSA_CHECK(H % 4 == 0, MSA_E_UNSUPPORTED, "lstm_rec: hidden size %d must be a multiple of 4", H);
    MSA_CHECK(B >= 1 && B <= 16, MSA_E_UNSUPPORTED, "lstm_rec: batch %d outside [1,16]", B);
    const int ncta_dir = sm_count / ndir;
    MSA_CHECK(ncta_dir >= 1 && (H + ncta_dir - 1) / ncta_dir <= kUMax, MSA_E_UNSUPPORTED,
              "lstm_rec: hidden size %d needs more than %d units per CTA on %d CTAs", H, kUMax, ncta_dir);
    return 0;
}

size_t lstm_rec_fwd_smem(int B, int H) {
    const int BP = (B + 3) & ~3;
    return sizeof(float) * ((size_t)4 * kUMax * H + (size_t)BP * H + kRecWarps * 32 + 32 * BP + kUMax * BP);
}
size_t lstm_rec_bwd_smem(int B, int H) {
    const int BP = (B + 3) & ~3;
    return sizeof(float) * ((size_t)kUMax * 4 * H + kRecWarps * 32 + 2 * kUMax * BP);
}

int launch_lstm_rec_fwd(const LstmRecParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(rec_check(p.B, p.H, p.ndir, sm_count));
    const size_t smem = lstm_rec_fwd_smem(p.B, p.H);
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "lstm_rec_fwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    MSA_CUDA(cudaFuncSetAttribute(k_lstm_rec_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSA_CUDA(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), st));
    LstmRecParams pp = p;
    void* args[] = {&pp};
    MSA_CUDA(cudaLaunchCooperativeKernel((void*)k_lstm_rec_fwd, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

int launch_lstm_rec_bwd(const LstmRecBwdParams& p, int sm_count, size_t smem_limit, cudaStream_t st) {
    MSA_TRY(rec_check(p.B, p.H, p.ndir, sm_count));
    const size_t smem = lstm_rec_bwd_smem(p.B, p.H);
    MSA_CHECK(smem <= smem_limit, MSA_E_UNSUPPORTED, "lstm_rec_bwd: needs %zu bytes of shared memory (> %zu)", smem, smem_limit);
    MSA_CUDA(cudaFuncSetAttribute(k_lstm_rec_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSA_CUDA(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), st));
    LstmRecBwdParams pp = p;
    void* args[] = {&pp};
    MSA_CUDA(cudaLaunchCooperativeKernel((void*)k_lstm_rec_bwd, dim3(sm_count), dim3(kRecThreads), args, smem, st));
    count_launch();
    return 0;
}

}  // namespace msa
