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

// Loads of data written by OTHER thread blocks of the same (persistent) kernel: bypass L1.
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Grid-wide barrier for cooperative (co-resident) launches: monotonically increasing arrival counter.
struct GridBarrier {
    unsigned int* counter;  // zeroed before launch
    unsigned int target;    // per-thread copy of the next release value
    __device__ __forceinline__ void init(unsigned int* c) { counter = c; target = 0; }
    __device__ __forceinline__ void sync() {
        target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();  // release: publish this block's writes (cumulative over the bar.sync)
            atomicAdd(counter, 1u);
            unsigned int v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while ((int)(v - target) < 0);
            __threadfence();  // acquire side + L1 invalidation for the block's later plain loads
        }
        __syncthreads();
    }
};
#endif  // __CUDACC__

}  // namespace msa
