// Shared device/host helpers for the sm_100a kernels of the hot path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/msa_b200.h"

namespace msa {

// ---- error plumbing (no exceptions cross the C ABI) ---------------------------------
void set_error(const char* fmt, ...);
#define MSA_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            msa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)
#define MSA_CHECK(cond, code, ...)                                                         \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            msa::set_error(__VA_ARGS__);                                                   \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)
#define MSA_TRY(expr)                                                                      \
    do {                                                                                   \
        int _r = (expr);                                                                   \
        if (_r != 0) return _r;                                                            \
    } while (0)
// every kernel of this library is launched through a wrapper that ends in MSA_LAUNCH_CHECK (or count_launch for
// cooperative launches), so msa_launch_count() is the number of OUR kernels enqueued (cuBLAS calls are not counted)
void count_launch();
#define MSA_LAUNCH_CHECK()              \
    do {                                \
        msa::count_launch();            \
        MSA_CUDA(cudaGetLastError());   \
    } while (0)

constexpr int kAlign = 32;  // floats; every flat-buffer tensor starts on a 128-byte boundary
inline int64_t align_up(int64_t n, int64_t a = kAlign) { return (n + a - 1) / a * a; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// ---- device helpers ----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Sum over a thread block; result valid in every thread.  `red` = shared scratch of >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// 32 per-lane accumulators -> lane i holds the warp-wide total of accumulator i (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce32(float (&a)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            // keep the half of the values that matches this lane's bit, send the other half
            float keep = upper ? a[i + half] : a[i];
            float send = upper ? a[i] : a[i + half];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return a[0];
}

// 16 per-lane accumulators -> lanes i and i+16 hold the warp-wide total of accumulator i (16 shuffles).
__device__ __forceinline__ float warp_transpose_reduce16(float (&a)[16]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            float keep = upper ? a[i + half] : a[i];
            float send = upper ? a[i] : a[i + half];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 16);
}

// Loads of data written by OTHER thread blocks of the same (persistent) kernel: bypass L1.
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- cross-CTA hand-off inside the persistent kernels: "canary polling" -------------------------------------------
// Every per-step array that one CTA produces and other CTAs consume in the same launch (h, q, energies, dz, ...)
// is ALSO the stash the backward pass / the batched GEMMs read later, so each word is written exactly once per launch.
// The host fills those arrays with 0xFF bytes before the launch; a consumer spins on the data words themselves until
// none of them is the canary.  Aligned 32-bit words are single-copy atomic, so no flag, no fence and no atomic is
// needed: one L2 round trip from the producer's store to the consumer's load.  0xFFFFFFFF is a NaN pattern no
// arithmetic instruction produces (hardware NaNs are 0x7FFFFFFF), so finite data can never be mistaken for it.
constexpr unsigned int kCanary = 0xFFFFFFFFu;
__device__ __forceinline__ bool is_canary(float x) { return __float_as_uint(x) == kCanary; }
__device__ __forceinline__ bool ready4(const float4& v) { return !(is_canary(v.x) | is_canary(v.y) | is_canary(v.z) | is_canary(v.w)); }
// Fast-path test of several polled 128-bit loads at once: the canary is the all-ones word, the largest unsigned value, so "some word
// is still the canary" == "the unsigned maximum over the words is the canary" (two 3-input max instructions per load, ONE compare and
// branch per batch; the exact per-load test and the re-poll loop sit behind that branch)
__device__ __forceinline__ unsigned int umax_acc(unsigned int m, const float4& v) {
    m = __vimax3_u32(m, __float_as_uint(v.x), __float_as_uint(v.y));
    return __vimax3_u32(m, __float_as_uint(v.z), __float_as_uint(v.w));
}
__device__ __forceinline__ float ld_poll(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_poll4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pub(float* p, float v) {
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_pub4(float* p, const float4& v) {
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Bounded spinning: a producer that never arrives (a bug, or NaN data) must not hang the GPU.  After ~4M empty polls a
// thread raises the abort word of the handle; every spinner checks it every 1024 polls and gives up.  The word is sticky
// (no pass clears it): the trainers fold it into the gradient-norm scalar on the device (msa_abort_guard: the clip + optimizer
// kernels then skip the update) and read it back with the step's own device->host traffic (msa_abort_read_async); tests
// call msa_check_abort, which synchronises.
struct SpinGuard {
    unsigned int* abort_word;
    unsigned int n;
    bool dead;      // sticky: once this thread has seen the abort it never waits again (the launch drains in microseconds)
    __device__ __forceinline__ explicit SpinGuard(unsigned int* a) : abort_word(a), n(0), dead(false) {}
    __device__ __forceinline__ bool bail() {
        if (dead) return true;
#ifdef MSA_POLL_BACKOFF
        __nanosleep(MSA_POLL_BACKOFF);      // experiment: thin out the polling traffic
#endif
        if ((++n & 1023u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) dead = true;
            else if (n > (1u << 21)) {
                *reinterpret_cast<volatile unsigned int*>(abort_word) = 1u;
                dead = true;
            }
        }
        return dead;
    }
    __device__ __forceinline__ void reset() { n = 0; }
};
__device__ __forceinline__ float poll1(const float* p, SpinGuard& sg) {
    float v = ld_poll(p);
    sg.reset();
    while (is_canary(v)) {
        if (sg.bail()) break;
        v = ld_poll(p);
    }
    return v;
}
__device__ __forceinline__ float4 poll4(const float* p, SpinGuard& sg) {
    float4 v = ld_poll4(p);
    sg.reset();
    while (!ready4(v)) {
        if (sg.bail()) break;
        v = ld_poll4(p);
    }
    return v;
}
#endif  // __CUDACC__

}  // namespace msa
