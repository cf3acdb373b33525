// LSTMCell of the free-running decoder step (decoder.py:253-255, 262-264) with the operand stream on the TMA engine.
//
// Same decomposition as ker_infer_rows (infer_decode.cu): the 4H gate rows are split by hidden unit over all SMs, every CTA
// streams its 4 x 7 rows of [W_ih | W_hh] and ALL of x = [input | h] through shared memory, multiplies on the tensor cores
// (3 x TF32 mma.sync, hi/lo split) and finishes with the point-wise cell update.  What changes is the loader: ker_infer_rows
// ingests ~20 B/clk per SM through LDGSTS whatever the ring depth (profiles/r01_infer_notes.txt), and one 1-D bulk copy per
// 512-byte row is slower still.  Here a chunk of 256 columns is 16 tensor-map boxes: per 32-column slab one 3-D box
// {32 columns, 7 units, 4 gates} of the weight matrix (viewed as [4][H][K]) and one 2-D box {32 columns, 32 batch rows} of x,
// written with the 128-byte swizzle so that ldmatrix reads them without bank conflicts; completion is counted in bytes on one
// mbarrier per ring stage, rows / columns outside the tensor are zero-filled by the copy engine.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace msa {

constexpr int kLtThreads = 512;
constexpr int kLtKC = 256;                       // columns per chunk = 8 slabs of 32
constexpr int kLtSlabs = kLtKC / 32;
constexpr int kLtNS = 3;                         // ring stages
constexpr int kLtSlabBytes = 8192;               // x box 32 rows x 128 B + W box (28 rows, padded to 32) x 128 B
constexpr int kLtStageBytes = kLtSlabs * kLtSlabBytes;
constexpr int kLtU = 7;                          // hidden units per CTA (box height)
constexpr unsigned kLtBoxBytesX = 32 * 128, kLtBoxBytesW = 4 * kLtU * 128;

__device__ __forceinline__ unsigned lt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lt_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lt_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lt_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0, spins = 0;
    const unsigned addr = lt_smem_u32(bar);
    while (!ok) {
        if (++spins > (1u << 26)) __trap();      // a byte-count mismatch must fail the launch, not hang the GPU
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void lt_tma_2d(unsigned smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(map), "r"(c0), "r"(c1), "r"(lt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lt_tma_3d(unsigned smem_dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(lt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lt_ldsm_x4(unsigned (&r)[4], unsigned smem_addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}
__device__ __forceinline__ void lt_split(unsigned v, unsigned& hi, unsigned& lo) {
    hi = v & 0xffffe000u;
    lo = __float_as_uint(__uint_as_float(v) - __uint_as_float(hi));
}
// round-to-nearest TF32 of an fp32 bit pattern (half an ulp of the 10-bit mantissa added to the magnitude, then truncated)
__device__ __forceinline__ unsigned lt_round(unsigned v) { return (v + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void lt_mma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct LstmTmaArgs {
    int B, H, K0, K1;
    const float* bias_ih; const float* bias_hh;
    float* c; float* h1; int ldh1; float* h2; int ldh2;
    const int* state;
};

// kMma: products per k8 step and accumulator tile.  3 = hi/lo split of both operands (fp32-accurate; the GEMM policies with an
// fp32 forward, msa_config.gemm_tf32 = 0 / 1); 1 = plain TF32, both operands rounded to nearest (policy 2, "TF32 everywhere");
// 2 = weights hi/lo, activations rounded (MSA_INFER_MMA=2, measured only).  Measured on B200 at the default dimensions against
// the fp32 oracle over 150 free-running steps: relative error of mel_post 1.5e-6 / 6.0e-5 / 8.9e-5 (no growth with the horizon),
// 76.6 / 73.2 / 70.5 us per step.
template <int kMma>
__global__ void __launch_bounds__(kLtThreads, 1)
ker_infer_lstm_tma(const __grid_constant__ CUtensorMap mx0, const __grid_constant__ CUtensorMap mw0,
                   const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mw1, LstmTmaArgs p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) unsigned char lt_raw[];
    __shared__ __align__(8) unsigned long long full_bar[kLtNS];
    constexpr int R = 4 * kLtU, NW = kLtThreads / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b0 = blockIdx.y * 32, nb = min(32, p.B - b0);
    const int u0 = (int)(((long long)blockIdx.x * p.H) / gridDim.x);
    const int U = (int)(((long long)(blockIdx.x + 1) * p.H) / gridDim.x) - u0;
    const unsigned ring = (lt_smem_u32(lt_raw) + 1023u) & ~1023u;      // the 128-byte swizzle pattern is anchored at 1 KB boundaries
    float* red = reinterpret_cast<float*>(lt_raw + (ring - lt_smem_u32(lt_raw)));
    if (threadIdx.x == 0) {
        for (int s = 0; s < kLtNS; ++s) lt_mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nch0 = (p.K0 + kLtKC - 1) / kLtKC, nch1 = (p.K1 + kLtKC - 1) / kLtKC, nch = nch0 + nch1;
    // warp 0: lane j < 8 copies the x box of slab j, lane 8 + j the weight box of slab j
    auto issue = [&](int c, int stage, bool weights, bool inputs, bool arm) {
        const int s = c < nch0 ? 0 : 1, k0 = (s ? c - nch0 : c) * kLtKC, K = s ? p.K1 : p.K0;
        const int nslab = min(kLtSlabs, (K - k0 + 31) / 32);
        if (arm && lane == 0) lt_mbar_expect_tx(&full_bar[stage], (unsigned)nslab * (kLtBoxBytesX + kLtBoxBytesW));
        __syncwarp();
        const int j = lane & 7;
        if (j < nslab) {
            const unsigned dst = ring + (unsigned)stage * kLtStageBytes + (unsigned)j * kLtSlabBytes;
            if (lane < 8 && inputs) lt_tma_2d(dst, s ? &mx1 : &mx0, k0 + 32 * j, b0, &full_bar[stage]);
            else if (lane >= 8 && lane < 16 && weights) lt_tma_3d(dst + 4096, s ? &mw1 : &mw0, k0 + 32 * j, u0, 0, &full_bar[stage]);
        }
    };
    // the weights do not depend on the previous kernel of the step: their first chunks are requested before the dependency wait
    if (w == 0)
        for (int c = 0; c < kLtNS && c < nch; ++c) issue(c, c, true, false, true);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.state[1]) {
        // (launch-uniform) drain the copies already in flight before leaving: shared memory must not be written after exit
        if (w == 0)
            for (int c = 0; c < kLtNS && c < nch; ++c) issue(c, c, false, true, false);
        for (int c = 0; c < kLtNS && c < nch; ++c) lt_mbar_wait(&full_bar[c], 0);
        return;
    }
    if (w == 0)
        for (int c = 0; c < kLtNS && c < nch; ++c) issue(c, c, false, true, false);

    // epilogue operands of the point-wise threads, fetched before the K loop (thread = (unit, batch row))
    float pre_b[4] = {0.f, 0.f, 0.f, 0.f}, pre_c = 0.f;
    if ((int)threadIdx.x < kLtU * 32) {
        const int ul = threadIdx.x >> 5, b = threadIdx.x & 31, u = u0 + ul;
        if (ul < U && b < nb) {
#pragma unroll
            for (int g = 0; g < 4; ++g) pre_b[g] = p.bias_ih[g * p.H + u] + p.bias_hh[g * p.H + u];
            pre_c = p.c[(size_t)(b0 + b) * p.H + u];
        }
    }
    float acc[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    // ldmatrix lane roles: matrix mi = lane >> 3, row mr = lane & 7 of that matrix.  In a slab, row r of a box lives at r*128 and
    // its 16-byte chunk c at ((c ^ (r & 7)) * 16) (128-byte swizzle); r & 7 == mr for every row this lane addresses.
    const int mi = lane >> 3, mr = lane & 7;
    const unsigned a_row = 4096u + (unsigned)((mi & 1) * 8 + mr) * 128u;      // + m*16*128; chunk = 2*j2 + (mi >> 1)
    const unsigned b_row = (unsigned)((mi >> 1) * 8 + mr) * 128u;             // + np*16*128; chunk = 2*j2 + (mi & 1)

    for (int c = 0; c < nch; ++c) {
        const int stage = c % kLtNS;
        const int s = c < nch0 ? 0 : 1, k0c = (s ? c - nch0 : c) * kLtKC;
        const int kvalid = min(kLtKC, (s ? p.K1 : p.K0) - k0c);
        lt_mbar_wait(&full_bar[stage], (unsigned)(c / kLtNS) & 1u);
        const unsigned base = ring + (unsigned)stage * kLtStageBytes;
#pragma unroll
        for (int ks = 0; ks < kLtKC / 8 / NW; ++ks) {
            const int kk = ks * NW + w;                 // k8 step of this warp within the chunk
            if (kk * 8 >= kvalid) continue;             // (columns beyond K inside a copied slab are zero-filled)
            const unsigned slab = base + (unsigned)(kk >> 2) * kLtSlabBytes;
            const int j2 = kk & 3;
            unsigned bh[4][2], bl[4][2];
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                unsigned r4[4];
                lt_ldsm_x4(r4, slab + b_row + (unsigned)np * 2048u + (unsigned)(((2 * j2 + (mi & 1)) ^ mr) * 16));
                if (kMma == 3) {
                    lt_split(r4[0], bh[2 * np][0], bl[2 * np][0]);
                    lt_split(r4[1], bh[2 * np][1], bl[2 * np][1]);
                    lt_split(r4[2], bh[2 * np + 1][0], bl[2 * np + 1][0]);
                    lt_split(r4[3], bh[2 * np + 1][1], bl[2 * np + 1][1]);
                } else {
                    bh[2 * np][0] = lt_round(r4[0]); bh[2 * np][1] = lt_round(r4[1]);
                    bh[2 * np + 1][0] = lt_round(r4[2]); bh[2 * np + 1][1] = lt_round(r4[3]);
                }
            }
            unsigned ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                unsigned r4[4];
                lt_ldsm_x4(r4, slab + a_row + (unsigned)m * 2048u + (unsigned)(((2 * j2 + (mi >> 1)) ^ mr) * 16));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (kMma >= 2) lt_split(r4[i], ah[m][i], al[m][i]);
                    else ah[m][i] = lt_round(r4[i]);
                }
            }
            if (kMma >= 2) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < 4; ++n) lt_mma(acc[m][n], al[m], bh[n][0], bh[n][1]);
            }
            if (kMma == 3) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < 4; ++n) lt_mma(acc[m][n], ah[m], bl[n][0], bl[n][1]);
            }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) lt_mma(acc[m][n], ah[m], bh[n][0], bh[n][1]);
        }
        __syncthreads();      // every warp is done with this stage
        if (w == 0 && c + kLtNS < nch) issue(c + kLtNS, stage, true, true, true);
    }
    // split-K partials -> shared (the ring is free: every copy has been consumed): red[w][rl][b], row stride 40
    constexpr int RS = 40;
    {
        const int g = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int rl = m * 16 + g + hh * 8, b = n * 8 + 2 * tq;
                    if (rl < R) *reinterpret_cast<float2*>(red + ((size_t)w * R + rl) * RS + b) = make_float2(acc[m][n][2 * hh], acc[m][n][2 * hh + 1]);
                }
    }
    __syncthreads();
    if ((int)threadIdx.x < kLtU * 32) {
        const int ul = threadIdx.x >> 5, b = threadIdx.x & 31, u = u0 + ul;
        if (ul < U && b < nb) {
            float z[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float sacc = 0.f;
#pragma unroll
                for (int q = 0; q < NW; ++q) sacc += red[((size_t)q * R + g * kLtU + ul) * RS + b];
                z[g] = sacc + pre_b[g];
            }
            const float gi = sigmoidf_(z[0]), gf = sigmoidf_(z[1]), gg = tanhf(z[2]), go = sigmoidf_(z[3]);
            const float cn = gf * pre_c + gi * gg;
            p.c[(size_t)(b0 + b) * p.H + u] = cn;
            const float hv = go * tanhf(cn);
            p.h1[(size_t)(b0 + b) * p.ldh1 + u] = hv;
            if (p.h2) p.h2[(size_t)(b0 + b) * p.ldh2 + u] = hv;
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
typedef CUresult (*LtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static LtEncodeFn lt_encode() {
    static LtEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<LtEncodeFn>(p);
    }
    return fn;
}

bool infer_lstm_tma_supported(int H, int K0, int ld0, int K1, int ld1, int sm_count) {
    if (lt_encode() == nullptr) return false;
    // at most 7 units per CTA on one wave of CTAs, 16-byte aligned rows
    const int gx = H < sm_count ? H : sm_count;
    return (H + gx - 1) / gx <= kLtU && K0 % 4 == 0 && K1 % 4 == 0 && ld0 % 4 == 0 && ld1 % 4 == 0;
}

// x: [B][ld] row-major, K columns -> 2-D map {K, B}, box {32, 32}
int infer_lstm_tma_map_x(void* map_out, const float* x, int B, int K, int ld) {
    LtEncodeFn enc = lt_encode();
    MSA_CHECK(enc != nullptr, MSA_E_NODEVICE, "infer_lstm_tma: cuTensorMapEncodeTiled is not available from this driver");
    MSA_CHECK(((uintptr_t)x & 15) == 0, MSA_E_UNSUPPORTED, "infer_lstm_tma: input not 16-byte aligned");
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)B};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSA_CHECK(r == CUDA_SUCCESS, MSA_E_ARG, "infer_lstm_tma: cuTensorMapEncodeTiled failed (%d) for x [%d x %d], ld %d", (int)r, B, K, ld);
    return 0;
}
// W: [4H][ld] row-major (gate-major rows g*H + u), K columns -> 3-D map {K, H, 4}, box {32, 7, 4}
int infer_lstm_tma_map_w(void* map_out, const float* W, int H, int K, int ld) {
    LtEncodeFn enc = lt_encode();
    MSA_CHECK(enc != nullptr, MSA_E_NODEVICE, "infer_lstm_tma: cuTensorMapEncodeTiled is not available from this driver");
    MSA_CHECK(((uintptr_t)W & 15) == 0, MSA_E_UNSUPPORTED, "infer_lstm_tma: weights not 16-byte aligned");
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)H, 4};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), (cuuint64_t)H * ld * sizeof(float)};
    const cuuint32_t box[3] = {32, (cuuint32_t)kLtU, 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(W), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSA_CHECK(r == CUDA_SUCCESS, MSA_E_ARG, "infer_lstm_tma: cuTensorMapEncodeTiled failed (%d) for W [4 x %d x %d], ld %d", (int)r, H, K, ld);
    return 0;
}

int k_infer_lstm_tma(const InferLstmTmaLaunch& a, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)kLtNS * kLtStageBytes + 1024;
    static int mma_env = -1;
    if (mma_env < 0) {
        const char* e = getenv("MSA_INFER_MMA");
        mma_env = e ? atoi(e) : 0;
        if (mma_env < 0 || mma_env > 3) mma_env = 0;
        MSA_CUDA(cudaFuncSetAttribute(ker_infer_lstm_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MSA_CUDA(cudaFuncSetAttribute(ker_infer_lstm_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MSA_CUDA(cudaFuncSetAttribute(ker_infer_lstm_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    LstmTmaArgs p{};
    p.B = a.B; p.H = a.H; p.K0 = a.K0; p.K1 = a.K1; p.bias_ih = a.bias_ih; p.bias_hh = a.bias_hh;
    p.c = a.c; p.h1 = a.h1; p.ldh1 = a.ldh1; p.h2 = a.h2; p.ldh2 = a.ldh2; p.state = a.state;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.H < sm_count ? a.H : sm_count, (a.B + 31) / 32);
    cfg.blockDim = dim3(kLtThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int mma = mma_env ? mma_env : (a.tf32 ? 1 : 3);
    const CUtensorMap &m0 = *static_cast<const CUtensorMap*>(a.map_x0), &m1 = *static_cast<const CUtensorMap*>(a.map_w0);
    const CUtensorMap &m2 = *static_cast<const CUtensorMap*>(a.map_x1), &m3 = *static_cast<const CUtensorMap*>(a.map_w1);
    if (mma == 1) MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_lstm_tma<1>, m0, m1, m2, m3, p));
    else if (mma == 2) MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_lstm_tma<2>, m0, m1, m2, m3, p));
    else MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_lstm_tma<3>, m0, m1, m2, m3, p));
    MSA_LAUNCH_CHECK();
    return 0;
}

}  // namespace msa
