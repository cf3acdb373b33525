// Latency of the load/store flavours used for cross-CTA hand-offs (development tool).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int F> __device__ __forceinline__ unsigned int ld(const unsigned int* p) {
    unsigned int v;
    if (F == 0) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 2) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 3) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 4) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (F == 5) asm volatile("ld.relaxed.cta.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// dependent chain of loads: buf[i] holds the index of the next element
template <int F> __global__ void chase(const unsigned int* buf, int n, long long* out, unsigned int* sink) {
    unsigned int idx = 0;
    for (int i = 0; i < 64; ++i) idx = ld<F>(buf + idx);     // warm
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) idx = ld<F>(buf + idx);
    const long long t1 = clock64();
    out[0] = t1 - t0;
    sink[0] = idx;
}
template <int LF, int SF> __global__ void pingpong(unsigned int* a, unsigned int* b, int n, long long* out) {
    // block 0 writes a, waits b; block 1 waits a, writes b
    const long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
        if (blockIdx.x == 0) {
            if (SF == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a), "r"(i) : "memory");
            if (SF == 1) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(a), "r"(i) : "memory");
            if (SF == 2) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(a), "r"(i) : "memory");
            if (SF == 3) atomicExch(a, (unsigned int)i);
            if (SF == 4) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(a), "r"(i) : "memory");
            while (ld<LF>(b) != (unsigned int)i) {}
        } else {
            while (ld<LF>(a) != (unsigned int)i) {}
            if (SF == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(b), "r"(i) : "memory");
            if (SF == 1) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(b), "r"(i) : "memory");
            if (SF == 2) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(b), "r"(i) : "memory");
            if (SF == 3) atomicExch(b, (unsigned int)i);
            if (SF == 4) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(b), "r"(i) : "memory");
        }
    }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
}
int main() {
    CK(cudaSetDevice(0));
    const int N = 1 << 16;   // 256 KB: L2 resident
    unsigned int* h = (unsigned int*)malloc(N * 4);
    for (int i = 0; i < N; ++i) h[i] = (unsigned int)((i + 4099 * 32) % N);   // stride of many lines
    unsigned int *buf, *sink, *a, *b; long long* out;
    CK(cudaMalloc(&buf, N * 4)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&a, 256)); CK(cudaMalloc(&b, 256));
    CK(cudaMemcpy(buf, h, N * 4, cudaMemcpyHostToDevice));
    const int n = 2000; long long c;
    const char* ln[] = {"ld.global.cg", "ld.relaxed.gpu", "ld.volatile", "ld.acquire.gpu", "ld.global.ca", "ld.relaxed.cta"};
#define CHASE(F) chase<F><<<1, 1>>>(buf, n, out, sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost)); printf("dependent-load latency %-16s: %6.0f cycles\n", ln[F], (double)c / n);
    CHASE(0) CHASE(1) CHASE(2) CHASE(3) CHASE(4) CHASE(5)
    const char* sn[] = {"st.relaxed.gpu", "st.global.cg", "st.volatile", "atomicExch", "st.global.wt"};
#define PP(LF, SF) CK(cudaMemset(a, 0, 256)); CK(cudaMemset(b, 0, 256)); pingpong<LF, SF><<<2, 1>>>(a, a + 32, n, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost)); printf("ping-pong %-16s + %-16s: %6.0f cycles round trip\n", sn[SF], ln[LF], (double)c / n);
    PP(1, 0) PP(0, 0) PP(0, 1) PP(2, 2) PP(0, 3) PP(1, 3) PP(0, 4) PP(3, 0)
    return 0;
}
