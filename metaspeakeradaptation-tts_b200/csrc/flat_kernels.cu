// Fused multi-tensor kernels over the flat fp32 parameter buffer.
//
// Replaces the per-tensor Python loops of the reference trainers:
//   higher diffopt.step SGD rule      (maml.py:54, reptile.py:56; torch.optim.SGD)
//   mix_grad / apply_grad             (utils/grad_utils.py:8-31, maml.py:94-99)
//   Reptile delta                     (reptile.py:42,73-77)
//   clip_grad_norm_ + outer step      (maml.py:101-105, reptile.py:85-89)
//   EWC Fisher / penalty / step       (continual_ewc.py:59-89,345-357)
// All are HBM-bound streaming kernels: 128-bit loads/stores, grid = a multiple of the SM
// count, grid-stride loops, deterministic two-stage reductions (no float atomics).
#include "common.cuh"

namespace msa {

constexpr int kFlatThreads = 256;
constexpr int kFlatBlocks = 148 * 8;  // 8 resident blocks of 256 threads per SM
constexpr int kPartials = kFlatBlocks;

static inline int flat_grid(int64_t n4) {
    int64_t b = (n4 + kFlatThreads - 1) / kFlatThreads;
    return (int)(b < kFlatBlocks ? (b > 0 ? b : 1) : kFlatBlocks);
}

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ void st4(float* p, int64_t i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }

#define FLAT_LOOP(n4) for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n4); i += (int64_t)gridDim.x * blockDim.x)
#define FOR4(body) { const int c = 0; body } { const int c = 1; body } { const int c = 2; body } { const int c = 3; body }
#define C4(v) (reinterpret_cast<float*>(&(v))[c])

// p_out may alias p (in-place step on the fast weights): p carries no __restrict__
__global__ void __launch_bounds__(kFlatThreads) k_sgd_step(const float* p, const float* __restrict__ g,
                                                         float* p_out, float* buf, int64_t n4, float lr, float mom,
                                                         float damp, float wd, int nesterov, int first) {
    FLAT_LOOP(n4) {
        float4 pv = ld4(p, i), gv = ld4(g, i), o;
        if (mom != 0.f) {
            float4 bv = first ? make_float4(0, 0, 0, 0) : ld4(buf, i);
            FOR4(float gg = C4(gv) + wd * C4(pv); float b = first ? gg : mom * C4(bv) + (1.f - damp) * gg; C4(bv) = b;
                 float d = nesterov ? gg + mom * b : b; C4(o) = C4(pv) - lr * d;)
            st4(buf, i, bv);
        } else {
            FOR4(float gg = C4(gv) + wd * C4(pv); C4(o) = C4(pv) - lr * gg;)
        }
        st4(p_out, i, o);
    }
}

__global__ void __launch_bounds__(kFlatThreads) k_axpy(float* acc, const float* __restrict__ g, int64_t n4, float w, int init) {
    FLAT_LOOP(n4) {
        float4 gv = ld4(g, i), a = init ? make_float4(0, 0, 0, 0) : ld4(acc, i);
        FOR4(C4(a) = C4(a) + w * C4(gv);)
        st4(acc, i, a);
    }
}

__global__ void __launch_bounds__(kFlatThreads) k_reptile_delta(float* acc, const float* __restrict__ pT,
                                                              const float* __restrict__ p0, int64_t n4, float w, int init) {
    FLAT_LOOP(n4) {
        float4 a = init ? make_float4(0, 0, 0, 0) : ld4(acc, i), t = ld4(pT, i), z = ld4(p0, i);
        FOR4(C4(a) = C4(a) + w * (-(C4(t) - C4(z)));)
        st4(acc, i, a);
    }
}

__global__ void __launch_bounds__(kFlatThreads) k_sumsq_partial(const float* __restrict__ g, int64_t n4, float* partials) {
    __shared__ float red[33];
    float s = 0.f;
    FLAT_LOOP(n4) {
        float4 v = ld4(g, i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_final_sum(const float* __restrict__ partials, int n, float* out) {
    __shared__ float red[33];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

__device__ __forceinline__ float clip_coef(const float* sumsq, float max_norm) {
    if (max_norm <= 0.f) return 1.f;
    float c = max_norm / (sqrtf(sumsq[0]) + 1e-6f);
    return c < 1.f ? c : 1.f;
}

__global__ void __launch_bounds__(kFlatThreads) k_clip_sgd(float* p, const float* __restrict__ g, float* buf,
                                                         const float* __restrict__ sumsq, int64_t n4, float lr,
                                                         float max_norm, float mom, float damp, float wd, int nesterov,
                                                         int first) {
    if (sumsq != nullptr && !isfinite(sumsq[0])) return;      // msa_abort_guard / diverged gradient: leave p and the state alone
    const float coef = clip_coef(sumsq, max_norm);
    FLAT_LOOP(n4) {
        float4 pv = ld4(p, i), gv = ld4(g, i);
        if (mom != 0.f) {
            float4 bv = first ? make_float4(0, 0, 0, 0) : ld4(buf, i);
            FOR4(float gg = coef * C4(gv) + wd * C4(pv); float b = first ? gg : mom * C4(bv) + (1.f - damp) * gg;
                 C4(bv) = b; float d = nesterov ? gg + mom * b : b; C4(pv) = C4(pv) - lr * d;)
            st4(buf, i, bv);
        } else {
            FOR4(float gg = coef * C4(gv) + wd * C4(pv); C4(pv) = C4(pv) - lr * gg;)
        }
        st4(p, i, pv);
    }
}

// p_out may alias p (in-place outer step) or be a fresh buffer (functional inner-loop step: p stays intact)
__global__ void __launch_bounds__(kFlatThreads) k_clip_adam(const float* p, float* p_out, const float* __restrict__ g, float* m, float* v,
                                                          const float* __restrict__ sumsq, int64_t n4, float lr, float b1,
                                                          float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                          float max_norm) {
    if (sumsq != nullptr && !isfinite(sumsq[0])) return;      // msa_abort_guard / diverged gradient: leave p and the state alone
    const float coef = clip_coef(sumsq, max_norm);
    const float step_size = lr / bc1;
    FLAT_LOOP(n4) {
        float4 pv = ld4(p, i), gv = ld4(g, i), mv = ld4(m, i), vv = ld4(v, i);
        FOR4(float gg = coef * C4(gv) + wd * C4(pv); float mm = b1 * C4(mv) + (1.f - b1) * gg;
             float v2 = b2 * C4(vv) + (1.f - b2) * gg * gg; C4(mv) = mm; C4(vv) = v2;
             float denom = sqrtf(v2) / bc2_sqrt + eps; C4(pv) = C4(pv) - step_size * (mm / denom);)
        st4(m, i, mv);
        st4(v, i, vv);
        st4(p_out, i, pv);
    }
}

__global__ void __launch_bounds__(kFlatThreads) k_fisher(float* f, const float* __restrict__ g, int64_t n4, float inv_n, int init) {
    FLAT_LOOP(n4) {
        float4 gv = ld4(g, i), a = init ? make_float4(0, 0, 0, 0) : ld4(f, i);
        FOR4(C4(a) = C4(a) + C4(gv) * C4(gv) * inv_n;)
        st4(f, i, a);
    }
}

// penalty partials; do_step 1: also the fused SGD step p -= lr*(g + 2*lam*F*(p-mu)); do_step 2: the penalty gradient added to g
// (g += 2*lam*F*(p-mu), p untouched) for an optimizer whose update is not plain SGD
__global__ void __launch_bounds__(kFlatThreads) k_ewc(float* p, float* g, const float* __restrict__ mu,
                                                    const float* __restrict__ f, int64_t n4, float lr, float lam,
                                                    float* partials, int do_step) {
    __shared__ float red[33];
    float s = 0.f;
    FLAT_LOOP(n4) {
        float4 pv = ld4(p, i), mv = ld4(mu, i), fv = ld4(f, i);
        float4 gv = do_step ? ld4(g, i) : make_float4(0, 0, 0, 0);
        FOR4(float d = C4(pv) - C4(mv); s += C4(fv) * d * d;
             if (do_step == 1) C4(pv) = C4(pv) - lr * (C4(gv) + 2.f * lam * C4(fv) * d);
             if (do_step == 2) C4(gv) = C4(gv) + 2.f * lam * C4(fv) * d;)
        if (do_step == 1) st4(p, i, pv);
        if (do_step == 2) st4(g, i, gv);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

static int check_flat(const void* a, int64_t n) {
    MSA_CHECK(a != nullptr, MSA_E_ARG, "flat kernel: null buffer");
    MSA_CHECK(n > 0 && n % 4 == 0, MSA_E_ARG, "flat kernel: n=%lld must be a positive multiple of 4", (long long)n);
    MSA_CHECK(((uintptr_t)a & 15) == 0, MSA_E_ARG, "flat kernel: buffer not 16-byte aligned");
    return 0;
}

}  // namespace msa

using namespace msa;

extern "C" {

int msa_flat_partials(void) { return kPartials; }

int msa_flat_sgd_step(const float* p, const float* g, float* p_out, float* momentum_buf, int64_t n, float lr, float momentum,
                      float dampening, float weight_decay, int nesterov, int first_step, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(g, n)); MSA_TRY(check_flat(p_out, n));
    if (momentum != 0.f) MSA_TRY(check_flat(momentum_buf, n));
    k_sgd_step<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(p, g, p_out, momentum_buf, n / 4, lr, momentum,
                                                                            dampening, weight_decay, nesterov, first_step);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_flat_axpy(float* acc, const float* g, int64_t n, float w, int init, void* stream) {
    MSA_TRY(check_flat(acc, n)); MSA_TRY(check_flat(g, n));
    k_axpy<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(acc, g, n / 4, w, init);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_flat_reptile_delta(float* acc, const float* p_T, const float* p_0, int64_t n, float w, int init, void* stream) {
    MSA_TRY(check_flat(acc, n)); MSA_TRY(check_flat(p_T, n)); MSA_TRY(check_flat(p_0, n));
    k_reptile_delta<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(acc, p_T, p_0, n / 4, w, init);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_flat_sumsq(const float* g, int64_t n, float* partials, float* out, void* stream) {
    MSA_TRY(check_flat(g, n));
    MSA_CHECK(partials && out, MSA_E_ARG, "msa_flat_sumsq: null scratch/out");
    int grid = flat_grid(n / 4);
    k_sumsq_partial<<<grid, kFlatThreads, 0, (cudaStream_t)stream>>>(g, n / 4, partials);
    MSA_LAUNCH_CHECK();
    k_final_sum<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, grid, out);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_flat_clip_sgd(float* p, const float* g, float* momentum_buf, const float* sumsq, int64_t n, float lr, float max_norm,
                      float momentum, float dampening, float weight_decay, int nesterov, int first_step, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(g, n));
    MSA_CHECK(max_norm <= 0.f || sumsq != nullptr, MSA_E_ARG, "msa_flat_clip_sgd: clipping needs sumsq");
    if (momentum != 0.f) MSA_TRY(check_flat(momentum_buf, n));
    k_clip_sgd<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(p, g, momentum_buf, sumsq, n / 4, lr, max_norm,
                                                                            momentum, dampening, weight_decay, nesterov, first_step);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_flat_clip_adam(float* p, const float* g, float* m, float* v, const float* sumsq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, float max_norm, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(g, n)); MSA_TRY(check_flat(m, n)); MSA_TRY(check_flat(v, n));
    MSA_CHECK(step >= 1, MSA_E_ARG, "msa_flat_clip_adam: step is 1-based");
    MSA_CHECK(max_norm <= 0.f || sumsq != nullptr, MSA_E_ARG, "msa_flat_clip_adam: clipping needs sumsq");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    k_clip_adam<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(p, p, g, m, v, sumsq, n / 4, lr, beta1, beta2, eps,
                                                                             weight_decay, (float)bc1, (float)sqrt(bc2), max_norm);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_flat_adam_step(const float* p, const float* g, float* p_out, float* m, float* v, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(g, n)); MSA_TRY(check_flat(p_out, n)); MSA_TRY(check_flat(m, n)); MSA_TRY(check_flat(v, n));
    MSA_CHECK(step >= 1, MSA_E_ARG, "msa_flat_adam_step: step is 1-based");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    k_clip_adam<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(p, p_out, g, m, v, nullptr, n / 4, lr, beta1, beta2, eps,
                                                                             weight_decay, (float)bc1, (float)sqrt(bc2), 0.f);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_ewc_fisher_accum(float* fisher, const float* g, int64_t n, float inv_n_batches, int init, void* stream) {
    MSA_TRY(check_flat(fisher, n)); MSA_TRY(check_flat(g, n));
    k_fisher<<<flat_grid(n / 4), kFlatThreads, 0, (cudaStream_t)stream>>>(fisher, g, n / 4, inv_n_batches, init);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_ewc_penalty(const float* p, const float* mu, const float* fisher, int64_t n, float* partials, float* out, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(mu, n)); MSA_TRY(check_flat(fisher, n));
    MSA_CHECK(partials && out, MSA_E_ARG, "msa_ewc_penalty: null scratch/out");
    int grid = flat_grid(n / 4);
    k_ewc<<<grid, kFlatThreads, 0, (cudaStream_t)stream>>>(const_cast<float*>(p), nullptr, mu, fisher, n / 4, 0.f, 0.f, partials, 0);
    MSA_LAUNCH_CHECK();
    k_final_sum<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, grid, out);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_ewc_sgd_step(float* p, const float* g, const float* mu, const float* fisher, int64_t n, float lr, float lam,
                     float* partials, float* penalty_out, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(g, n)); MSA_TRY(check_flat(mu, n)); MSA_TRY(check_flat(fisher, n));
    MSA_CHECK(partials && penalty_out, MSA_E_ARG, "msa_ewc_sgd_step: null scratch/out");
    int grid = flat_grid(n / 4);
    k_ewc<<<grid, kFlatThreads, 0, (cudaStream_t)stream>>>(p, const_cast<float*>(g), mu, fisher, n / 4, lr, lam, partials, 1);
    MSA_LAUNCH_CHECK();
    k_final_sum<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, grid, penalty_out);
    MSA_LAUNCH_CHECK();
    return 0;
}

int msa_ewc_penalty_grad(const float* p, float* g, const float* mu, const float* fisher, int64_t n, float lam, float* partials,
                         float* penalty_out, void* stream) {
    MSA_TRY(check_flat(p, n)); MSA_TRY(check_flat(g, n)); MSA_TRY(check_flat(mu, n)); MSA_TRY(check_flat(fisher, n));
    MSA_CHECK(partials && penalty_out, MSA_E_ARG, "msa_ewc_penalty_grad: null scratch/out");
    int grid = flat_grid(n / 4);
    k_ewc<<<grid, kFlatThreads, 0, (cudaStream_t)stream>>>(const_cast<float*>(p), g, mu, fisher, n / 4, 0.f, lam, partials, 2);
    MSA_LAUNCH_CHECK();
    k_final_sum<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, grid, penalty_out);
    MSA_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
