// Hand-written sm_100a tensor-core GEMM for the batched (non-recurrent) contractions of the pass:
//     C[M x N] = alpha * op(A) . op(B) + beta * C          (row-major)
// Forward projections of Tacotron2NV are "x . W^T" (both operands K-contiguous = K-major tiles); the backward pass needs the
// input gradients dX = dY . W (B is [K][N], N-contiguous) and the weight gradients dW = dY^T . X (both operands are [K][.] with
// the contraction index as the ROW index) -- those operands are MN-major tiles.  tcgen05.mma kind::tf32 reads both kinds straight
// from shared memory (instruction-descriptor bits 15 / 16), so no operand is ever transposed in memory: a K-major tile is one
// TMA box {32 K-floats, 128 rows} with the 128-byte swizzle, an MN-major tile four boxes {32 MN-floats, 32 K-rows} with the
// 128-byte swizzle of 32-byte atoms (the only MN-major layout the tensor core accepts for 32-bit operands).
// fp32 accumulators in TMEM, a 3-stage mbarrier pipeline, one TMA warp + one MMA warp + four epilogue warps.
//
// Precision: a single TF32 product loses 13 mantissa bits of each operand, which the forward pass cannot afford
// (mel_post 2.5e-3 vs the 1e-3 tolerance).  Mode 0 therefore evaluates the "3xTF32" split
//     A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo,   x_hi = the 19 bits the tensor core reads, x_lo = x - x_hi (exact in fp32)
// (three MMAs per k-step on tiles of the four operand arrays; the lo arrays are produced by a streaming split kernel), which
// is fp32-accurate to ~2^-21.  Mode 1 is the plain single TF32 product (backward-pass policy).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace msa {

constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32;        // CTA tile; BK floats = one 128-byte swizzle row
constexpr int kTcStages = 3;
constexpr int kTcTileBytes = kTcBM * kTcBK * 4;             // 16 KB per operand tile
constexpr int kTcThreads = 320;                             // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-9 lo-tile converters, 2-5 epilogue
constexpr int kTcConv = kTcThreads - 64;                    // converter threads
constexpr int kTcTmemCols = 128;

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor: K-major operand tile [rows][32 floats] with the 128-byte swizzle TMA wrote
// (canonical layout ((8,n),2):((8,SBO),1) in 16-byte units: LBO = 1, SBO = 8 rows * 128 B = 64, version 1, layout SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// shared-memory matrix descriptor: MN-major operand tile = four blocks [32 K-rows][32 MN-floats = 128 B].  For 32-bit operands the
// tensor core transposes 32-byte units, so MN-major TF32 tiles must use the "128-byte swizzle with 32-byte atoms" (layout type 1,
// SWIZZLE_128B_BASE32B; TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): canonical layout
// ((T,8,m),(4,k)):((1,T,LBO),(.,SBO)), T = 4 floats per 16 bytes -- LBO = distance between the 32-float MN blocks (4 KB),
// SBO = distance between groups of 4 K-rows (512 B); one UMMA (K = 8) reads the K-rows [8j, 8j+8).
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, M = 128, N = 128; bit 15 / 16: A / B is MN-major
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate, uint32_t idesc = kTcIdesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TcSmem {
    uint64_t full[kTcStages], conv[kTcStages], empty[kTcStages], tmem_full;
    uint32_t tmem_base;
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// x_lo = x - (the 19 bits the tensor core reads), exact in fp32
__device__ __forceinline__ float tc_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// grid = (M tiles, N tiles, K splits).  Split z handles the K blocks [z * kb_per_split, ...) and, when `partial` is given, stores
// its raw accumulator tile to partial[z][M][N] (ker_splitk_reduce applies alpha / beta in a fixed order); otherwise it writes C.
// kSplit (3xTF32): only the fp32 tiles A and B come in by TMA; the four converter / epilogue warps derive the lo tiles in shared
// memory (same swizzled position, 32 KB further up) while the next stage is in flight, which halves the bytes every SM ingests.
template <bool kSplit>
__global__ void __launch_bounds__(kTcThreads, 1)
k_gemm_tf32_nt(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ C, int ldc,
               int M, int N, int K, float alpha, float beta, float* __restrict__ partial, int kb_per_split,
               const float* __restrict__ bias1, const float* __restrict__ bias2, int amn, int bmn) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned operand tiles (swizzle atom = 8 rows x 128 B), then the barriers
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int kOps = kSplit ? 4 : 2;
    TcSmem* sb = reinterpret_cast<TcSmem*>(tiles + (size_t)kTcStages * kOps * kTcTileBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * kTcBN;
    const int num_kb = (K + kTcBK - 1) / kTcBK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(num_kb, kb0 + kb_per_split) - kb0;          // >= 1 (host)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&sb->full[s], 1); mbar_init(&sb->conv[s], kTcConv); mbar_init(&sb->empty[s], 1); }
        mbar_init(&sb->tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation (whole warp), 128 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sb->tmem_base)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = sb->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kTcStages, it = i / kTcStages;
                mbar_wait(&sb->empty[s], (it & 1) ^ 1);                       // slot free (first pass: passes immediately)
                mbar_expect_tx(&sb->full[s], 2 * kTcTileBytes);
                uint8_t* st = tiles + (size_t)s * kOps * kTcTileBytes;
                if (!amn) tma_load_2d(st, &mapA, (kb0 + i) * kTcBK, m0, &sb->full[s]);
                else
                    for (int b = 0; b < 4; ++b) tma_load_2d(st + b * 4096, &mapA, m0 + 32 * b, (kb0 + i) * kTcBK, &sb->full[s]);
                if (!bmn) tma_load_2d(st + kTcTileBytes, &mapB, (kb0 + i) * kTcBK, n0, &sb->full[s]);
                else
                    for (int b = 0; b < 4; ++b) tma_load_2d(st + kTcTileBytes + b * 4096, &mapB, n0 + 32 * b, (kb0 + i) * kTcBK, &sb->full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kTcStages, it = i / kTcStages;
                mbar_wait(kSplit ? &sb->conv[s] : &sb->full[s], it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(tiles + (size_t)s * kOps * kTcTileBytes);
                const uint32_t idesc = kTcIdesc | (amn ? 1u << 15 : 0u) | (bmn ? 1u << 16 : 0u);
                const uint64_t dA = amn ? umma_desc_mn128(st) : umma_desc_k128(st);
                const uint64_t dB = bmn ? umma_desc_mn128(st + kTcTileBytes) : umma_desc_k128(st + kTcTileBytes);
                const uint64_t dAl = amn ? umma_desc_mn128(st + 2 * kTcTileBytes) : umma_desc_k128(st + 2 * kTcTileBytes);
                const uint64_t dBl = bmn ? umma_desc_mn128(st + 3 * kTcTileBytes) : umma_desc_k128(st + 3 * kTcTileBytes);
                // one UMMA = 8 K-floats: 32 bytes further in a K-major tile (2 descriptor units), 8 K-rows = 1 KB further (64 units)
                const uint64_t sa = amn ? 64 : 2, sbk = bmn ? 64 : 2;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {
                    const uint64_t ada = (uint64_t)k * sa, adb = (uint64_t)k * sbk;
                    umma_tf32(tmem_acc, dA + ada, dB + adb, (i | k) != 0, idesc);
                    if (kSplit) {
                        umma_tf32(tmem_acc, dAl + ada, dB + adb, 1u, idesc);
                        umma_tf32(tmem_acc, dA + ada, dBl + adb, 1u, idesc);
                    }
                }
                umma_commit(&sb->empty[s]);                                    // frees the stage when these MMAs retire
            }
            umma_commit(&sb->tmem_full);                                       // accumulator complete
        }
    } else {
        const int q = warp & 3;
        if (kSplit) {
            // ===== lo tiles: [A | B] (32 KB, as TMA swizzled them) -> [A_lo | B_lo] at the same positions 32 KB further up =====
            const int ct = threadIdx.x - 64;                                   // 0 .. kTcConv-1
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kTcStages, it = i / kTcStages;
                mbar_wait(&sb->full[s], it & 1);
                const float4* src = reinterpret_cast<const float4*>(tiles + (size_t)s * kOps * kTcTileBytes);
                float4* dst = reinterpret_cast<float4*>(tiles + (size_t)s * kOps * kTcTileBytes + 2 * kTcTileBytes);
                constexpr int kPer = 2 * kTcTileBytes / 16 / kTcConv;          // float4 words per thread (8)
                float4 v[kPer];
#pragma unroll
                for (int j = 0; j < kPer; ++j) v[j] = src[ct + j * kTcConv];
#pragma unroll
                for (int j = 0; j < kPer; ++j) dst[ct + j * kTcConv] = make_float4(tc_lo(v[j].x), tc_lo(v[j].y), tc_lo(v[j].z), tc_lo(v[j].w));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
                mbar_arrive(&sb->conv[s]);
            }
        }
        if (warp >= 6) goto done;                                              // converter-only warps
        // ===== epilogue: TMEM -> registers -> shared (transpose) -> global; warp q handles TMEM lanes [32q, 32q+32) =====
        // (a thread owns an accumulator ROW; going through a 32 x 33 tile in the now idle operand ring turns the stores into
        // 128-byte row segments instead of 32 scattered words per instruction)
        mbar_wait(&sb->tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* xp = reinterpret_cast<float*>(tiles) + q * (32 * 33);
        float* obase = partial != nullptr ? partial + (size_t)blockIdx.z * M * N : C;
        const int ldo = partial != nullptr ? N : ldc;
#pragma unroll 1
        for (int cc = 0; cc < kTcBN / 32; ++cc) {
            uint32_t v[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) xp[lane * 33 + c] = __uint_as_float(v[c]);
            __syncwarp();
            const int col = n0 + cc * 32 + lane;
            if (col < N) {
                const int rows = min(32, M - (m0 + q * 32));
                float* out = obase + (size_t)(m0 + q * 32) * ldo + col;
                if (partial == nullptr && beta != 0.f) {
                    // C is read for all 32 rows before the first store: one memory round trip per tile column block instead of 32
                    // dependent ones (most products of the pass accumulate onto a bias-filled C with beta = 1)
                    float old[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) old[r] = r < rows ? __ldcs(out + (size_t)r * ldo) : 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        if (r < rows) out[(size_t)r * ldo] = alpha * xp[r * 33 + lane] + beta * old[r];
                } else {
                    // bias1 / bias2: column bias instead of a bias-filled C (C is then write-only: no fill launch, no read of C)
                    const float sc = partial != nullptr ? 1.f : alpha;
                    float bs = 0.f;
                    if (partial == nullptr && bias1 != nullptr) bs = bias1[col] + (bias2 != nullptr ? bias2[col] : 0.f);
#pragma unroll 8
                    for (int r = 0; r < rows; ++r) out[(size_t)r * ldo] = sc * xp[r * 33 + lane] + bs;
                }
            }
            __syncwarp();
        }
    }
done:
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTcTmemCols) : "memory");
    }
}

// C = alpha * (partial[0] + partial[1] + ...) + beta * C, splits summed in index order (deterministic)
__global__ void ker_splitk_reduce(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C, int ldc,
                                  float alpha, float beta, const float* __restrict__ bias1, const float* __restrict__ bias2) {
    const int64_t n = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = partial[i];
        for (int z = 1; z < splits; ++z) s += partial[(size_t)z * n + i];
        const int r = (int)(i / N), c = (int)(i - (int64_t)r * N);
        float* out = C + (size_t)r * ldc + c;
        const float bs = bias1 != nullptr ? bias1[c] + (bias2 != nullptr ? bias2[c] : 0.f) : 0.f;
        *out = beta != 0.f ? alpha * s + beta * *out + bs : alpha * s + bs;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// K-major operand: memory [rows = M or N][K] (K contiguous), box {32 K, 128 rows}; MN-major operand (mn): memory [K][rows = M or N]
// (the M / N index contiguous), box {32 MN, 32 K-rows}
static int make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, bool mn = false) {
    EncodeTiledFn enc = get_encode();
    MSA_CHECK(enc != nullptr, MSA_E_NODEVICE, "gemm_tc: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)(mn ? rows : K), (cuuint64_t)(mn ? K : rows)};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)(mn ? kTcBK : kTcBM)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSA_CHECK(r == CUDA_SUCCESS, MSA_E_ARG, "gemm_tc: cuTensorMapEncodeTiled failed (%d) for a [%d x %d] operand, ld %d", (int)r, rows, K, ld);
    return 0;
}

bool gemm_tc_supported(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb, const float* C, int64_t ldc) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    (void)C; (void)ldc;
    return M >= 1 && N >= 1 && K >= 1 && (lda % 4) == 0 && (ldb % 4) == 0 && al16(A) && al16(B) && M < (1 << 30) && N < (1 << 30) &&
           K < (1 << 30);
}
// K split: under-filled grids (few output tiles, long K) are spread over ~one wave of the 148 SMs, at least 4 K blocks per split
static int tc_plan_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ((M + kTcBM - 1) / kTcBM) * ((N + kTcBN - 1) / kTcBN);
    const int64_t num_kb = (K + kTcBK - 1) / kTcBK;
    int64_t splits = std::min<int64_t>(148 / std::max<int64_t>(tiles, 1), num_kb / 4);
    if (splits < 2) return 1;
    const int64_t per = (num_kb + splits - 1) / splits;
    return (int)((num_kb + per - 1) / per);
}
size_t gemm_tc_scratch_floats(int64_t M, int64_t N, int64_t K) {
    const int splits = tc_plan_splits(M, N, K);
    return splits > 1 ? (size_t)splits * M * N + 64 : 64;
}

// mode 0: 3xTF32 (fp32-accurate); mode 1: single TF32 product.  `scratch` (gemm_tc_scratch_floats(M, N, K) floats, 16-byte
// aligned) holds the partial tiles of a K split; with scratch == nullptr the K range is not split.
int gemm_tc_nt(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
               float* C, int64_t ldc, int mode, float* scratch, cudaStream_t st, const float* bias1, const float* bias2) {
    return gemm_tc(false, true, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, mode, scratch, st, bias1, bias2);
}

// General form, row-major: ta: A is given as [K][M] (else [M][K]); tb: B is given as [N][K] (else [K][N]).
int gemm_tc(bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
            float beta, float* C, int64_t ldc, int mode, float* scratch, cudaStream_t st, const float* bias1, const float* bias2) {
    MSA_CHECK(bias1 != nullptr || bias2 == nullptr, MSA_E_ARG, "gemm_tc: bias2 without bias1");
    MSA_CHECK(bias1 == nullptr || beta == 0.f, MSA_E_ARG, "gemm_tc: a column bias replaces the accumulation onto C (beta must be 0)");
    MSA_CHECK(gemm_tc_supported(M, N, K, A, lda, B, ldb, C, ldc), MSA_E_UNSUPPORTED, "gemm_tc: operand alignment / leading dimensions");
    const int amn = ta ? 1 : 0, bmn = tb ? 0 : 1;
    CUtensorMap mA, mB;
    MSA_TRY(make_map(&mA, A, (int)M, (int)K, (int)lda, amn != 0));
    MSA_TRY(make_map(&mB, B, (int)N, (int)K, (int)ldb, bmn != 0));
    const int num_kb = (int)((K + kTcBK - 1) / kTcBK);
    const int splits = scratch != nullptr ? tc_plan_splits(M, N, K) : 1;
    const int per = (num_kb + splits - 1) / splits;
    float* partial = splits > 1 ? scratch : nullptr;
    const dim3 grid((unsigned)((M + kTcBM - 1) / kTcBM), (unsigned)((N + kTcBN - 1) / kTcBN), (unsigned)splits);
    if (mode == 0) {
        const size_t smem = (size_t)kTcStages * 4 * kTcTileBytes + sizeof(TcSmem) + 1024;
        static bool attr0 = false;
        if (!attr0) { MSA_CUDA(cudaFuncSetAttribute(k_gemm_tf32_nt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr0 = true; }
        k_gemm_tf32_nt<true><<<grid, kTcThreads, smem, st>>>(mA, mB, C, (int)ldc, (int)M, (int)N, (int)K, alpha, beta, partial, per, bias1, bias2, amn, bmn);
    } else {
        const size_t smem = (size_t)kTcStages * 2 * kTcTileBytes + sizeof(TcSmem) + 1024;
        static bool attr1 = false;
        if (!attr1) { MSA_CUDA(cudaFuncSetAttribute(k_gemm_tf32_nt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr1 = true; }
        k_gemm_tf32_nt<false><<<grid, kTcThreads, smem, st>>>(mA, mB, C, (int)ldc, (int)M, (int)N, (int)K, alpha, beta, partial, per, bias1, bias2, amn, bmn);
    }
    MSA_LAUNCH_CHECK();
    if (splits > 1) {
        const int64_t n = M * N;
        ker_splitk_reduce<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(partial, splits, (int)M, (int)N, C, (int)ldc,
                                                                                                 alpha, beta, bias1, bias2);
        MSA_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace msa

extern "C" {
size_t msa_gemm_nt_scratch_floats(int64_t M, int64_t N, int64_t K) { return msa::gemm_tc_scratch_floats(M, N, K); }
int msa_gemm_nt(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                float* C, int64_t ldc, int mode, float* scratch, void* stream) {
    return msa::gemm_tc_nt(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, mode, scratch, (cudaStream_t)stream, nullptr, nullptr);
}
int msa_gemm(int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, const float* B,
             int64_t ldb, float beta, float* C, int64_t ldc, int mode, float* scratch, void* stream) {
    return msa::gemm_tc(trans_a != 0, trans_b != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, mode, scratch, (cudaStream_t)stream,
                        nullptr, nullptr);
}
}
