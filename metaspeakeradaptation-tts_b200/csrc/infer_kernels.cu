// Layout helpers of free-running inference (Decoder.infer, modules_tacotron2nv/decoder.py:334-411); the per-step kernels
// are in infer_decode.cu.
#include "common.cuh"
#include "kernels.h"

namespace msa {

// [T'][B][M] -> [B][M][ld] (reference layout with row stride ld >= T')
__global__ void ker_tm_to_ref_ld(const float* __restrict__ x_tm, float* out, int T, int B, int M, int ld) {
    const int64_t n = (int64_t)T * B * M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), b = (int)((i / M) % B), t = (int)(i / ((int64_t)M * B));
        out[((size_t)b * M + m) * ld + t] = x_tm[i];
    }
}
// [B][T'][M] -> [B][M][ld]
__global__ void ker_bt_to_ref_ld(const float* __restrict__ x_bt, float* out, int B, int T, int M, int ld) {
    const int64_t n = (int64_t)B * T * M;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), t = (int)((i / M) % T), b = (int)(i / ((int64_t)M * T));
        out[((size_t)b * M + m) * ld + t] = x_bt[i];
    }
}

__global__ void ker_fill_ones_i32(int* p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 1;
}
int k_fill_ones_i32(int* p, int n, cudaStream_t st) {
    ker_fill_ones_i32<<<cdiv(n, 256), 256, 0, st>>>(p, n);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_tm_to_ref_ld(const float* x_tm, float* out, int T, int B, int M, int ld, cudaStream_t st) {
    ker_tm_to_ref_ld<<<cdiv((int64_t)T * B * M, 256) > 2368 ? 2368 : cdiv((int64_t)T * B * M, 256), 256, 0, st>>>(x_tm, out, T, B, M, ld);
    MSA_LAUNCH_CHECK();
    return 0;
}
int k_bt_to_ref_ld(const float* x_bt, float* out, int B, int T, int M, int ld, cudaStream_t st) {
    ker_bt_to_ref_ld<<<cdiv((int64_t)T * B * M, 256) > 2368 ? 2368 : cdiv((int64_t)T * B * M, 256), 256, 0, st>>>(x_bt, out, B, T, M, ld);
    MSA_LAUNCH_CHECK();
    return 0;
}

}  // namespace msa
