// How fast can a persistent kernel (148 CTAs x 16 warps) re-stream a weight set that does not fit in shared memory?
// (development tool behind the "per-task weights" variants of the grouped recurrent kernels: every step each warp reads its
// fixed slice of the per-step footprint, either with cp.async into a private shared-memory ring or with 128-bit loads into
// registers, and consumes it.)       nvcc -arch=sm_100a -O3 -o stream_bench stream_bench.cu && ./stream_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kWarps = 16, kThreads = kWarps * 32;

__device__ __forceinline__ void cp16(void* smem, const void* g) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// footprint: [cta][warp][chunk][32 lanes] uint4, chunks_per_step chunks of CH*512 bytes per warp per step
template <int NS, int CH>      // NS ring stages of CH x 512 bytes per warp
__global__ void __launch_bounds__(kThreads, 1) k_ring(const uint4* __restrict__ w, int chunks, int steps, float* sink, long long* cyc) {
    extern __shared__ uint4 ring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* my = ring + (size_t)warp * NS * CH * 32;
    const uint4* src = w + ((size_t)blockIdx.x * kWarps + warp) * chunks * CH * 32;
    const long long total = (long long)chunks * steps;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < NS - 1; ++i) {
        if (i < total) {
            const uint4* s = src + (size_t)(i % chunks) * CH * 32;
#pragma unroll
            for (int c = 0; c < CH; ++c) cp16(my + ((size_t)i * CH + c) * 32 + lane, s + c * 32 + lane);
        }
        cp_commit();
    }
    for (long long i = 0; i < total; ++i) {
        const long long j = i + NS - 1;
        if (j < total) {
            const uint4* s = src + (size_t)(j % chunks) * CH * 32;
            uint4* d = my + (size_t)(j % NS) * CH * 32;
#pragma unroll
            for (int c = 0; c < CH; ++c) cp16(d + c * 32 + lane, s + c * 32 + lane);
        }
        cp_commit();
        cp_wait<NS - 1>();
        __syncwarp();
        const uint4* d = my + (size_t)(i % NS) * CH * 32;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const uint4 v = d[c * 32 + lane];
            acc += __uint_as_float(v.x) + __uint_as_float(v.y) + __uint_as_float(v.z) + __uint_as_float(v.w);
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

// same != 0: every CTA reads the SAME footprint (the all-gather pattern of the recurrent kernels: h / dz of all rows), warp slices
// rotated by the CTA index like the kernels do
template <int PF>               // PF 128-bit loads in flight per lane
__global__ void __launch_bounds__(kThreads, 1) k_regs(const uint4* __restrict__ w, int chunks, int steps, float* sink, long long* cyc, int same = 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4* src = same ? w + (size_t)((warp + blockIdx.x) % kWarps) * chunks * 32 : w + ((size_t)blockIdx.x * kWarps + warp) * chunks * 32;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int st = 0; st < steps; ++st) {
        for (int i = 0; i < chunks; i += PF) {
            uint4 v[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) v[k] = __ldcg(src + (size_t)(i + k) * 32 + lane);
#pragma unroll
            for (int k = 0; k < PF; ++k) acc += __uint_as_float(v[k].x) + __uint_as_float(v[k].y) + __uint_as_float(v[k].z) + __uint_as_float(v[k].w);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

// The dz / h gather of the grouped recurrent kernels: every CTA reads the same R rows of W floats; warp w owns the columns
// [w * KS * 16, (w + 1) * KS * 16) and per k16 step lane (lg, lj) loads one float4 of row mt * 8 + lg at column 4 * lj -- 64-byte
// segments of 8 rows per instruction.  MODE 0: ld.relaxed.gpu (the polling load of the kernels), 1: ld.global.cg, PF loads in flight.
template <int MODE, int PF, int MT>
__global__ void __launch_bounds__(kThreads, 1) k_gather(const float* __restrict__ x, int W, int steps, float* sink, long long* cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lg = lane >> 2, lj = lane & 3;
    const int wsl = (warp + blockIdx.x) % kWarps, KS = W / (kWarps * 16);
    float acc = 0.f;
    auto ld = [&](const float* p) {
        float4 v;
        if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
        else asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
        return v;
    };
    const long long t0 = clock64();
    for (int st = 0; st < steps; ++st) {
        for (int s0 = 0; s0 < KS; s0 += PF) {
            float4 v[PF][MT];
#pragma unroll
            for (int k = 0; k < PF; ++k)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) v[k][mt] = ld(x + (size_t)(mt * 8 + lg) * W + (wsl * KS + s0 + k) * 16 + 4 * lj);
#pragma unroll
            for (int k = 0; k < PF; ++k)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) acc += v[k][mt].x + v[k][mt].y + v[k][mt].z + v[k][mt].w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int ncta = pr.multiProcessorCount;
    float* sink; long long* cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, ncta * 8));
    long long* hc = (long long*)malloc(ncta * 8);
    const size_t maxb = (size_t)160 << 20;
    uint4* w;
    CK(cudaMalloc(&w, maxb));
    CK(cudaMemset(w, 0, maxb));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int mbs[] = {17, 34, 67, 100, 134};
    for (int mb : mbs) {
        // chunks of 512 bytes per warp per step
        const int chunks = (int)(((size_t)mb << 20) / ((size_t)ncta * kWarps * 512)) & ~7;
        const double bytes_step = (double)chunks * 512 * kWarps * ncta;
        const int steps = 40;
        auto report = [&](const char* name, float ms) {
            CK(cudaMemcpy(hc, cyc, ncta * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < ncta; ++i) mx = hc[i] > mx ? hc[i] : mx;
            printf("%6.1f MB/step %-22s: %7.2f us/step  %6.2f TB/s  %5.1f B/clk/SM\n", bytes_step / 1e6, name, ms * 1e3 / steps,
                   bytes_step * steps / (ms * 1e-3) / 1e12, bytes_step / ncta * steps / (double)mx);
        };
#define RUN(name, launch) \
        launch; CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); \
        { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); report(name, ms); }
        {
            constexpr int NS = 5, CH = 4;
            const int smem = kWarps * NS * CH * 512;
            CK(cudaFuncSetAttribute(k_ring<NS, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RUN("cp.async ring 5x2KB", (k_ring<NS, CH><<<ncta, kThreads, smem>>>(w, chunks / CH, steps, sink, cyc)))
        }
        {
            constexpr int NS = 3, CH = 4;
            const int smem = kWarps * NS * CH * 512;
            CK(cudaFuncSetAttribute(k_ring<NS, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RUN("cp.async ring 3x2KB", (k_ring<NS, CH><<<ncta, kThreads, smem>>>(w, chunks / CH, steps, sink, cyc)))
        }
        {
            constexpr int NS = 4, CH = 2;
            const int smem = kWarps * NS * CH * 512;
            CK(cudaFuncSetAttribute(k_ring<NS, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RUN("cp.async ring 4x1KB", (k_ring<NS, CH><<<ncta, kThreads, smem>>>(w, chunks / CH, steps, sink, cyc)))
        }
        RUN("ldcg.128 x4 in flight", (k_regs<4><<<ncta, kThreads>>>(w, chunks, steps, sink, cyc)))
        RUN("ldcg.128 x8 in flight", (k_regs<8><<<ncta, kThreads>>>(w, chunks, steps, sink, cyc)))
    }
    // all-gather pattern: every CTA reads the same 64 KB .. 1 MB per step
    for (int kb : {64, 128, 256, 512, 1024}) {
        const int chunks = kb * 1024 / (kWarps * 512);
        const double bytes_step = (double)chunks * 512 * kWarps * ncta;      // bytes delivered to the SMs per step
        const int steps = 200;
        auto report = [&](const char* name, float ms) {
            CK(cudaMemcpy(hc, cyc, ncta * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < ncta; ++i) mx = hc[i] > mx ? hc[i] : mx;
            printf("same %4d KB per CTA and step %-22s: %7.2f us/step  %6.2f TB/s delivered  %5.1f B/clk/SM\n", kb, name, ms * 1e3 / steps,
                   bytes_step * steps / (ms * 1e-3) / 1e12, bytes_step / ncta * steps / (double)mx);
        };
        RUN("ldcg.128 x4 in flight", (k_regs<4><<<ncta, kThreads>>>(w, chunks, steps, sink, cyc, 1)))
        RUN("ldcg.128 x8 in flight", (k_regs<8><<<ncta, kThreads>>>(w, chunks, steps, sink, cyc, 1)))
    }
    // row-gather pattern of the kernels: R = 8 * MT rows of 4096 floats, same for every CTA
    {
        const int W = 4096, steps = 200;
        auto rep = [&](const char* name, int R, float ms) {
            CK(cudaMemcpy(hc, cyc, ncta * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < ncta; ++i) mx = hc[i] > mx ? hc[i] : mx;
            printf("gather %2d rows x 16 KB per CTA and step, %-28s: %7.2f us/step  %5.1f B/clk/SM\n", R, name, ms * 1e3 / steps,
                   (double)R * W * 4 * steps / (double)mx);
        };
#define RUNG(name, R, launch) \
        launch; CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); \
        { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); rep(name, R, ms); }
        const float* x = reinterpret_cast<const float*>(w);
        RUNG("ld.relaxed.gpu, 2 in flight", 8, (k_gather<0, 2, 1><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.global.cg,   2 in flight", 8, (k_gather<1, 2, 1><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.relaxed.gpu, 4 in flight", 8, (k_gather<0, 4, 1><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.global.cg,   4 in flight", 8, (k_gather<1, 4, 1><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.relaxed.gpu, 2 in flight", 32, (k_gather<0, 2, 4><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.global.cg,   2 in flight", 32, (k_gather<1, 2, 4><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.relaxed.gpu, 4 in flight", 32, (k_gather<0, 4, 4><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
        RUNG("ld.global.cg,   4 in flight", 32, (k_gather<1, 4, 4><<<ncta, kThreads>>>(x, W, steps, sink, cyc)))
    }
    return 0;
}
