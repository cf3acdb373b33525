// Micro-benchmark of the cross-CTA hand-off schemes used by the persistent recurrent kernels (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_bench hop_bench.cu && ./hop_bench
// Every "step" each of the 148 co-resident CTAs publishes W words and then needs the words of ALL CTAs
// (an all-gather of 148*W words), like h(t) in the LSTM recurrence.  Reports cycles per step for several schemes.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float ld_poll(const float* p) { float v; asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_poll4(const float* p) { float4 v; asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_pub(float* p, float v) { asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

constexpr int WMAX = 32;
constexpr int NT = 512;

// mode 0: warp-0 gate on one sentinel per CTA, then bulk copy by all threads (checked)
// mode 1: gate only (no bulk copy)
// mode 2: mass polling: all threads poll their bulk words directly
// mode 3: mode 0 + background prefetch loads from DRAM issued before the gate
// mode 4: classic grid barrier (fence + atomic counter + poll by one thread), then bulk copy
// mode 5: mode 0 but publish with atomicExch
// mode 6: mode 0 but a __threadfence() after the publish
__global__ void __launch_bounds__(NT, 1) hop(float* buf, int steps, int mode, const float* dram, unsigned int* counter, long long* out, float* sink, int W, float* rep) {
    __shared__ float sm[148 * WMAX];
    const int ncta = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    float acc = 0.f;
    unsigned int target = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        float* slot = buf + (size_t)s * ncta * W;
        // "compute": a little dependent math so that the publish is not back-to-back with the previous gather
        float x = acc + 1.0f + (float)s;
        for (int i = 0; i < 16; ++i) x = x * 0.999f + 0.5f;
        const int R = mode == 12 ? 2 : mode == 13 ? 4 : mode == 14 ? 8 : mode == 15 ? 16 : 1;
        if (mode >= 12) {
            float* rslot = rep + (size_t)s * 16 * ncta * WMAX;
            if (tid < W) for (int r = 0; r < R; ++r) st_pub(rslot + (size_t)r * ncta * W + cta * W + tid, x);
            const float* mine = rslot + (size_t)(cta % R) * ncta * W;
            const int n4 = ncta * W / 4;
            for (int i = tid; i < n4; i += NT) {
                float4 v = ld_poll4(mine + i * 4);
                while (v.x == 0.f || v.y == 0.f || v.z == 0.f || v.w == 0.f) v = ld_poll4(mine + i * 4);
                reinterpret_cast<float4*>(sm)[i] = v;
            }
            __syncthreads();
            acc += sm[(tid * 7) % (ncta * W)];
            continue;
        }
        if (mode == 16 || mode == 17) {
            float* rslot = rep + (size_t)s * 16 * ncta * WMAX;
            if (tid < 4) st_pub(mode == 16 ? rslot + tid * ncta + cta : rslot + cta * 8 + tid, x);
            for (int i = tid; i < 4 * ncta; i += NT) {
                const float* a = mode == 16 ? rslot + i : rslot + (i >> 2) * 8 + (i & 3);
                float v = ld_poll(a);
                while (v == 0.f) v = ld_poll(a);
                sm[i] = v;
            }
            __syncthreads();
            acc += sm[(tid * 7) % (4 * ncta)];
            continue;
        }
        if (mode == 7) {
            if (tid == 0 && cta == 0) st_pub(slot + 0, 1.f);
            if (tid == 0 && cta == 1) { while (ld_poll(slot + 0) == 0.f) {} st_pub(slot + W, 1.f); }
        } else if (tid < W) {
            if (mode == 5) atomicExch(reinterpret_cast<unsigned int*>(slot + cta * W + tid), __float_as_uint(x));
            else st_pub(slot + cta * W + tid, x);
            if (mode == 6) __threadfence();
        }
        if (mode == 3 && tid < 128) acc += __ldg(dram + ((size_t)s * ncta + cta) * 4096 + tid * 32);
        if (mode == 4) {
            target += ncta;
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(counter, 1u);
                unsigned int v;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while ((int)(v - target) < 0);
                __threadfence();
            }
            __syncthreads();
        } else if (mode == 8 || mode == 9) {
            if (tid < 32) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int c = tid + 32 * j; v[j] = c < ncta ? ld_poll(slot + c * W + (W - 1)) : 1.f; }
#pragma unroll
                for (int j = 0; j < 8; ++j) { const int c = tid + 32 * j; while (v[j] == 0.f) v[j] = ld_poll(slot + c * W + (W - 1)); }
            }
            __syncthreads();
        } else if (mode == 10) {
            if (tid < ncta) { const float* a = slot + tid * W + (W - 1); while (ld_poll(a) == 0.f) {} }
            __syncthreads();
        } else if (mode == 11 || mode == 7) {
        } else if (mode != 2) {
            if (tid < 32) {
                for (int c = tid; c < ncta; c += 32) {
                    const float* a = slot + c * W + (W - 1);
                    while (ld_poll(a) == 0.f) {}
                }
            }
            __syncthreads();
        }
        if (mode == 7) {      // ping-pong between CTA 0 and CTA 1: one word each way per step
            if (tid == 0 && cta < 2) {
                const float* a = slot + (1 - cta) * W;
                if (cta == 0) { while (ld_poll(a) == 0.f) {} }
            }
            __syncthreads();
        } else if (mode == 11) {
            const int n4 = ncta * W / 4;
            if (tid < 128)
            for (int i = tid; i < n4; i += 128 * 4) {
                float4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) if (i + j * 128 < n4) v[j] = ld_poll4(slot + (i + j * 128) * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) if (i + j * 128 < n4) {
                    while (v[j].x == 0.f || v[j].y == 0.f || v[j].z == 0.f || v[j].w == 0.f) v[j] = ld_poll4(slot + (i + j * 128) * 4);
                    reinterpret_cast<float4*>(sm)[i + j * 128] = v[j];
                }
            }
            __syncthreads();
            acc += sm[(tid * 7) % (ncta * W)];
        } else if (mode != 1 && mode != 8) {
            const int n4 = ncta * W / 4;
            for (int i = tid; i < n4; i += NT) {
                float4 v = ld_poll4(slot + i * 4);
                while (v.x == 0.f || v.y == 0.f || v.z == 0.f || v.w == 0.f) v = ld_poll4(slot + i * 4);
                reinterpret_cast<float4*>(sm)[i] = v;
            }
            __syncthreads();
            acc += sm[(tid * 7) % (ncta * W)];
        }
    }
    const long long t1 = clock64();
    if (tid == 0) out[cta] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int nsm = prop.multiProcessorCount, steps = 400;
    for (int ncta : {nsm, nsm / 2, nsm / 4, 18, 8, 2}) {
    printf("---- %d CTAs ----\n", ncta);
    float *buf, *dram, *sink; unsigned int* counter; long long* out;
    const size_t nbuf = (size_t)steps * ncta * WMAX;
    CK(cudaMalloc(&buf, nbuf * 4)); CK(cudaMalloc(&dram, (size_t)steps * ncta * 4096 * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&counter, 4)); float* repbuf; CK(cudaMalloc(&repbuf, nbuf * 16 * 4)); CK(cudaMalloc(&out, ncta * 8));
    CK(cudaMemset(dram, 0, (size_t)steps * ncta * 4096 * 4));
    const char* names[] = {"gate(warp0 sentinels)+bulk", "gate only", "mass polling", "gate+bulk with DRAM prefetch loads in flight", "grid barrier (fence+atomic)+bulk", "gate+bulk, publish via atomicExch", "gate+bulk, __threadfence after publish", "ping-pong CTA0<->CTA1 (round trip)", "batched gate only", "batched gate + bulk", "distributed gate (1 sentinel/thread) + bulk", "polling by 128 threads, 4 loads in flight", "mass polling, 2 replicas", "mass polling, 4 replicas", "mass polling, 8 replicas", "mass polling, 16 replicas", "4 words/CTA interleaved [b][cta] (false sharing)", "4 words/CTA in a private 32B sector"};
    for (int Wv = 8; Wv <= 32; Wv *= 4)
    for (int mode = 0; mode < 18; ++mode) {
        if (mode != 0 && mode != 2 && mode != 7) continue;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(buf, 0, nbuf * 4)); CK(cudaMemset(repbuf, 0, nbuf * 16 * 4)); CK(cudaMemset(counter, 0, 4));
            int st = steps; int md = mode;
            int Wa = Wv; void* args[] = {&buf, &st, &md, &dram, &counter, &out, &sink, &Wa, &repbuf};
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void*)hop, dim3(ncta), dim3(NT), args, 0, 0));
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            long long h[256]; CK(cudaMemcpy(h, out, ncta * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < ncta; ++i) mx = h[i] > mx ? h[i] : mx;
            if (rep == 1) printf("W=%2d mode %2d %-48s: %8.0f cycles/step  (%.2f us/step by events)\n", Wv, mode, names[mode], (double)mx / steps, ms * 1e3 / steps);
        }
    }
    }
    return 0;
}
