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
//
// Implicit convolutions (Encoder / Postnet: conv1d "same" + BatchNorm, encoder.py:36-37, decoder.py:63-72).  The same kernel runs the
// three contractions of a k-tap convolution on channels-last activations x [B][T][Ci] WITHOUT an im2col matrix:
//   forward          y[b,t,co]  = bias[co] + sum_k sum_ci x[b,t+k-pad,ci] . Wp[k][co][ci]     K blocks = (tap, 32 input channels)
//   input gradient   dx[b,t,ci] = sum_k sum_co dy[b,t-k+pad,co] . Wp[k][co][ci]               K blocks = (tap, 32 output channels)
//   weight gradient  dW[co][ci][k] = sum_{b,t} dy[b,t,co] . x[b,t+k-pad,ci]                    K blocks = (batch row, 32 time steps)
// The activation operand is a 3-D tensor map {channels, T, B}: the tap shift is a coordinate offset and TMA's out-of-bounds
// zero fill IS the convolution's zero padding (per batch row).  Wp[k][co][ci] is the tap-major repack of the native weight
// (one small kernel per pass for all layers); the weight gradient is stored straight into the native [Co][Ci][K] layout.
// The forward epilogue also emits per-channel (count, mean, M2) of every 32-row slab of the output, so BatchNorm's batch statistics
// need no extra pass over y (merged with Chan's update in a fixed order by the normalisation kernel).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace msa {

constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32;        // CTA tile; BK floats = one 128-byte swizzle row
constexpr int kTcStages = 3;                                // 3xTF32: four 16 KB tiles per stage
constexpr int kTcStagesPlain = 6;                           // single-product modes: two tiles per stage, twice the depth (the
                                                            // round-to-nearest pass adds a hop between TMA and MMA)
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
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
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
    uint64_t full[kTcStagesPlain], conv[kTcStagesPlain], empty[kTcStagesPlain], tmem_full;
    uint32_t tmem_base;
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// x_lo = x - (the 19 bits the tensor core reads), exact in fp32
__device__ __forceinline__ float tc_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float tc_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Implicit-convolution addressing (mode 0: plain GEMM).  mode 1 (forward / input gradient): M tile = 128 time steps of ONE batch row
// (grid.x = B * tilesT), K block kb = (tap kb / cb, channel block kb % cb); A = 3-D activation map at time t0 + sign * (tap - pad),
// B = 2-D map over Wp viewed as [taps * brows][.].  mode 2 (weight gradient): grid.z = tap, K block kb = (batch row kb / cb, time
// block kb % cb), both operands MN-major 3-D activation maps, B shifted by tap - pad; the output element (row, col) of tap z lives
// at C[row * ldc + col * out_cs + z].
struct TcConv {
    int mode, taps, cb, pad, sign, Tn, tilesT, brows, out_cs;
};

// grid = (M tiles, N tiles, K splits).  Split z handles the K blocks [z * kb_per_split, ...) and, when `partial` is given, stores
// its raw accumulator tile to partial[z][M][N] (ker_splitk_reduce applies alpha / beta in a fixed order); otherwise it writes C.
// kSplit (3xTF32): only the fp32 tiles A and B come in by TMA; the four converter / epilogue warps derive the lo tiles in shared
// memory (same swizzled position, 32 KB further up) while the next stage is in flight, which halves the bytes every SM ingests.
// kMode 2 (single TF32 product with round-to-nearest operands): tcgen05.mma kind::tf32 TRUNCATES its fp32 operands to 19 bits (a
// biased rounding, twice the error of cuBLAS's TF32 path); the converter warps rewrite both tiles in place with cvt.rna.tf32.f32
// before the MMA warp reads them.
template <int kMode>
__global__ void __launch_bounds__(kTcThreads, 1)
k_gemm_tf32_nt(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ C, int ldc,
               int M, int N, int K, float alpha, float beta, float* __restrict__ partial, int kb_per_split,
               const float* __restrict__ bias1, const float* __restrict__ bias2, int amn, int bmn, TcConv cv, float* __restrict__ stats) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned operand tiles (swizzle atom = 8 rows x 128 B), then the barriers
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr bool kSplit = kMode == 0, kRn = kMode == 2 || kMode == 3, kRnBonly = kMode == 3;
    constexpr int kOps = kSplit ? 4 : 2;
    constexpr int kSt = kSplit ? kTcStages : kTcStagesPlain;
    TcSmem* sb = reinterpret_cast<TcSmem*>(tiles + (size_t)kSt * kOps * kTcTileBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * kTcBN;
    const int num_kb = (K + kTcBK - 1) / kTcBK;
    const int kb0 = cv.mode == 2 ? 0 : blockIdx.z * kb_per_split;
    const int nkb = min(num_kb, kb0 + kb_per_split) - kb0;          // >= 1 (host)
    // output rows of this tile: [orow0, orow0 + orows)
    int orow0 = m0, orows = min(kTcBM, M - m0), cbat = 0, ct0 = 0;
    if (cv.mode == 1) {
        cbat = blockIdx.x / cv.tilesT;
        ct0 = (blockIdx.x - cbat * cv.tilesT) * kTcBM;
        orow0 = cbat * cv.Tn + ct0;
        orows = min(kTcBM, cv.Tn - ct0);
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSt; ++s) { mbar_init(&sb->full[s], 1); mbar_init(&sb->conv[s], kTcConv); mbar_init(&sb->empty[s], 1); }
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
    // programmatic dependent launch: everything above (barriers, TMEM) overlapped the tail of the previous kernel of the stream; its
    // memory is visible after the wait.  Our own successor may start its prologue as soon as every CTA of this grid got here.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kSt, it = i / kSt;
                mbar_wait(&sb->empty[s], (it & 1) ^ 1);                       // slot free (first pass: passes immediately)
                mbar_expect_tx(&sb->full[s], 2 * kTcTileBytes);
                uint8_t* st = tiles + (size_t)s * kOps * kTcTileBytes;
                const int kb = kb0 + i;
                if (cv.mode == 1) {
                    const int tap = kb / cv.cb, c0 = (kb - tap * cv.cb) * kTcBK;
                    tma_load_3d(st, &mapA, c0, ct0 + cv.sign * (tap - cv.pad), cbat, &sb->full[s]);
                    if (!bmn) tma_load_2d(st + kTcTileBytes, &mapB, c0, tap * cv.brows + n0, &sb->full[s]);
                    else
                        for (int b = 0; b < 4; ++b) tma_load_2d(st + kTcTileBytes + b * 4096, &mapB, n0 + 32 * b, tap * cv.brows + c0, &sb->full[s]);
                    continue;
                }
                if (cv.mode == 2) {
                    const int bb = kb / cv.cb, t0 = (kb - bb * cv.cb) * kTcBK, tap = blockIdx.z;
                    for (int b = 0; b < 4; ++b) tma_load_3d(st + b * 4096, &mapA, m0 + 32 * b, t0, bb, &sb->full[s]);
                    for (int b = 0; b < 4; ++b) tma_load_3d(st + kTcTileBytes + b * 4096, &mapB, n0 + 32 * b, t0 + tap - cv.pad, bb, &sb->full[s]);
                    continue;
                }
                if (!amn) tma_load_2d(st, &mapA, kb * kTcBK, m0, &sb->full[s]);
                else
                    for (int b = 0; b < 4; ++b) tma_load_2d(st + b * 4096, &mapA, m0 + 32 * b, kb * kTcBK, &sb->full[s]);
                if (!bmn) tma_load_2d(st + kTcTileBytes, &mapB, kb * kTcBK, n0, &sb->full[s]);
                else
                    for (int b = 0; b < 4; ++b) tma_load_2d(st + kTcTileBytes + b * 4096, &mapB, n0 + 32 * b, kb * kTcBK, &sb->full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kSt, it = i / kSt;
                mbar_wait((kSplit || kRn) ? &sb->conv[s] : &sb->full[s], it & 1);
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
                const int s = i % kSt, it = i / kSt;
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
        if (kRn) {
            // ===== both tiles (mode 3: the B tile only -- A holds TF32-exact values already) in place: fp32 -> nearest TF32 value =====
            const int ct = threadIdx.x - 64;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % kSt, it = i / kSt;
                mbar_wait(&sb->full[s], it & 1);
                float4* t4 = reinterpret_cast<float4*>(tiles + (size_t)s * kOps * kTcTileBytes + (kRnBonly ? kTcTileBytes : 0));
                constexpr int kPer = (kRnBonly ? 1 : 2) * kTcTileBytes / 16 / kTcConv;
                float4 v[kPer];
#pragma unroll
                for (int j = 0; j < kPer; ++j) v[j] = t4[ct + j * kTcConv];
#pragma unroll
                for (int j = 0; j < kPer; ++j) t4[ct + j * kTcConv] = make_float4(tc_rn(v[j].x), tc_rn(v[j].y), tc_rn(v[j].z), tc_rn(v[j].w));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
        float* obase = partial != nullptr ? partial + (size_t)blockIdx.z * M * N : (cv.mode == 2 ? C + blockIdx.z : C);
        const int ldo = partial != nullptr ? N : ldc;
        const int ocs = partial != nullptr || cv.out_cs < 1 ? 1 : cv.out_cs;      // column stride of the output
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
                const int rows = min(32, orows - q * 32);
                float* out = obase + (size_t)(orow0 + q * 32) * ldo + (size_t)col * ocs;
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
                    if (stats != nullptr && partial == nullptr) {
                        // BatchNorm batch statistics of this 32-row slab: (count, mean, M2) per output channel, two passes over the tile
                        float sum = 0.f;
#pragma unroll 8
                        for (int r = 0; r < rows; ++r) {
                            const float v = sc * xp[r * 33 + lane] + bs;
                            out[(size_t)r * ldo] = v;
                            sum += v;
                        }
                        const float cnt = rows > 0 ? (float)rows : 0.f, mu = rows > 0 ? sum / cnt : 0.f;
                        float m2 = 0.f;
#pragma unroll 8
                        for (int r = 0; r < rows; ++r) {
                            const float dv = sc * xp[r * 33 + lane] + bs - mu;
                            m2 += dv * dv;
                        }
                        float* so = stats + (size_t)(blockIdx.x * 4 + q) * 3 * N + col;
                        so[0] = cnt;
                        so[N] = mu;
                        so[2 * (size_t)N] = m2;
                    } else {
#pragma unroll 8
                        for (int r = 0; r < rows; ++r) out[(size_t)r * ldo] = sc * xp[r * 33 + lane] + bs;
                    }
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

// The same reduction for a convolution's forward product, one block per (32 rows, 32 columns): also emits the slab's BatchNorm
// statistics (count, mean, M2) per column like the unsplit epilogue does.  block = (32 columns, 8 row groups of 4 rows)
__global__ void __launch_bounds__(256) ker_splitk_reduce_stats(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ C,
                                                                 int ldc, float alpha, const float* __restrict__ bias, float* __restrict__ stats) {
    __shared__ float sh[8][33];
    __shared__ float mu_s[32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.y * 32 + cx, r0 = blockIdx.x * 32 + ry * 4;
    const size_t n = (size_t)M * N;
    const float bs = (bias != nullptr && col < N) ? bias[col] : 0.f;
    const int nrow = min(32, M - blockIdx.x * 32);
    float v[4];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = 0.f;
        const int r = r0 + j;
        if (r < M && col < N) {
            float a = partial[(size_t)r * N + col];
            for (int z = 1; z < splits; ++z) a += partial[(size_t)z * n + (size_t)r * N + col];
            v[j] = alpha * a + bs;
            C[(size_t)r * ldc + col] = v[j];
            sum += v[j];
        }
    }
    sh[ry][cx] = sum;
    __syncthreads();
    if (ry == 0) {
        float t = 0.f;
        for (int j = 0; j < 8; ++j) t += sh[j][cx];
        mu_s[cx] = t / (float)nrow;
    }
    __syncthreads();
    const float mu = mu_s[cx];
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (r0 + j < M) m2 += (v[j] - mu) * (v[j] - mu);
    __syncthreads();
    sh[ry][cx] = m2;
    __syncthreads();
    if (ry == 0 && col < N) {
        float t = 0.f;
        for (int j = 0; j < 8; ++j) t += sh[j][cx];
        float* so = stats + (size_t)blockIdx.x * 3 * N + col;
        so[0] = (float)nrow;
        so[N] = mu;
        so[2 * (size_t)N] = t;
    }
}

// Wp[k][co][ci] <- W[co][ci][k] for up to 8 layers in one launch (blockIdx.y = layer)
struct RepackTab {
    const float* src[8];
    float* dst[8];
    int co[8], ci[8], k[8];
};
// block = (output channel co, chunk of 256 input channels): the chunk's [ci][k] run is contiguous in W (coalesced reads into shared
// memory), every tap's [ci] run is contiguous in Wp (coalesced writes)
constexpr int kRepackCi = 256, kRepackKmax = 15;
__global__ void __launch_bounds__(256) ker_conv_repack(RepackTab tab) {
    __shared__ float sh[kRepackCi * kRepackKmax];
    const int l = blockIdx.y;
    const float* __restrict__ src = tab.src[l];
    float* __restrict__ dst = tab.dst[l];
    const int Co = tab.co[l], Ci = tab.ci[l], K = tab.k[l];
    const int chunks = (Ci + kRepackCi - 1) / kRepackCi;
    for (int job = blockIdx.x; job < Co * chunks; job += gridDim.x) {
        const int co = job / chunks, ci0 = (job - co * chunks) * kRepackCi, nci = min(kRepackCi, Ci - ci0);
        const float* s0 = src + ((int64_t)co * Ci + ci0) * K;
        for (int i = threadIdx.x; i < nci * K; i += blockDim.x) sh[i] = s0[i];
        __syncthreads();
        for (int i = threadIdx.x; i < nci * K; i += blockDim.x) {
            const int k = i / nci, ci = i - k * nci;
            dst[((int64_t)k * Co + co) * Ci + ci0 + ci] = sh[ci * K + k];
        }
        __syncthreads();
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

// channels-last activation [B][Tn][C] as a 3-D map {C, Tn, B}; box {32 channels, rows time steps, 1}
static int make_map_act(CUtensorMap* map, const float* base, int B, int Tn, int C, int rows, bool mn) {
    EncodeTiledFn enc = get_encode();
    MSA_CHECK(enc != nullptr, MSA_E_NODEVICE, "conv_tc: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Tn, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)C * sizeof(float), (cuuint64_t)Tn * C * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)kTcBK, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSA_CHECK(r == CUDA_SUCCESS, MSA_E_ARG, "conv_tc: cuTensorMapEncodeTiled failed (%d) for a [%d][%d][%d] activation", (int)r, B, Tn, C);
    return 0;
}

bool conv_tc_supported(int B, int Tn, int Ci, int Co, int K, const float* x, const float* w) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return B >= 1 && Tn >= 1 && Ci >= 4 && Co >= 4 && (Ci % 4) == 0 && (Co % 4) == 0 && K >= 1 && (K & 1) == 1 && K <= 15 && al16(x) && al16(w);
}
static int conv_plan_splits(int B, int Tn, int N, int num_kb) {
    const int64_t tiles = (int64_t)B * ((Tn + kTcBM - 1) / kTcBM) * ((N + kTcBN - 1) / kTcBN);
    int64_t splits = std::min<int64_t>(148 / std::max<int64_t>(tiles, 1), num_kb / 4);
    if (splits < 2) return 1;
    const int64_t per = (num_kb + splits - 1) / splits;
    return (int)((num_kb + per - 1) / per);
}
// K-split scratch of the forward / input-gradient product with N output channels and C contraction channels
size_t conv_tc_scratch_floats(int B, int Tn, int C, int N, int K) {
    const int splits = conv_plan_splits(B, Tn, N, K * ((C + kTcBK - 1) / kTcBK));
    return splits > 1 ? (size_t)splits * B * Tn * N + 64 : 64;
}
// slabs of the forward statistics: stats[slabs][3][Co]
int conv_tc_stat_slabs(int B, int Tn, int Ci, int Co, int K, bool have_scratch) {
    const int splits = have_scratch ? conv_plan_splits(B, Tn, Co, K * ((Ci + kTcBK - 1) / kTcBK)) : 1;
    return splits > 1 ? (B * Tn + 31) / 32 : B * ((Tn + kTcBM - 1) / kTcBM) * 4;
}

static bool tc_pdl() {
    static const bool on = !(getenv("MSA_GEMM_PDL") && atoi(getenv("MSA_GEMM_PDL")) == 0);
    return on;
}
template <int kMode>
static int launch_tc(dim3 grid, cudaStream_t st, const CUtensorMap& mA, const CUtensorMap& mB, float* C, int ldc, int M, int N, int K, float alpha,
                     float beta, float* partial, int per, const float* bias1, const float* bias2, int amn, int bmn, const TcConv& cv, float* stats) {
    const size_t smem = (size_t)(kMode == 0 ? kTcStages * 4 : kTcStagesPlain * 2) * kTcTileBytes + sizeof(TcSmem) + 1024;
    static bool attr = false;
    if (!attr) { MSA_CUDA(cudaFuncSetAttribute(k_gemm_tf32_nt<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = tc_pdl() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    MSA_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_tf32_nt<kMode>, mA, mB, C, ldc, M, N, K, alpha, beta, partial, per, bias1, bias2, amn, bmn, cv, stats));
    MSA_LAUNCH_CHECK();
    return 0;
}

// y[B*Tn][Co] = conv1d_same(x[B][Tn][Ci], Wp[K][Co][Ci]) + bias; stats (optional): conv_tc_stat_slabs x [3][Co] slab statistics of y
int conv_tc_fwd(const float* x, int B, int Tn, int Ci, const float* wp, int Co, int K, const float* bias, float* y, int mode, float* scratch,
                float* stats, cudaStream_t st) {
    CUtensorMap mA, mB;
    MSA_TRY(make_map_act(&mA, x, B, Tn, Ci, kTcBM, false));
    MSA_TRY(make_map(&mB, wp, K * Co, Ci, Ci, false));
    const int cb = (Ci + kTcBK - 1) / kTcBK, tilesT = (Tn + kTcBM - 1) / kTcBM, num_kb = K * cb, M = B * Tn;
    const int splits = scratch != nullptr ? conv_plan_splits(B, Tn, Co, num_kb) : 1;
    const int per = (num_kb + splits - 1) / splits;
    float* partial = splits > 1 ? scratch : nullptr;
    const TcConv cv{1, K, cb, (K - 1) / 2, 1, Tn, tilesT, Co, 1};
    const dim3 grid((unsigned)(B * tilesT), (unsigned)((Co + kTcBN - 1) / kTcBN), (unsigned)splits);
    if (mode == 0) MSA_TRY(launch_tc<0>(grid, st, mA, mB, y, Co, M, Co, num_kb * kTcBK, 1.f, 0.f, partial, per, bias, nullptr, 0, 0, cv, stats));
    else if (mode == 2) MSA_TRY(launch_tc<2>(grid, st, mA, mB, y, Co, M, Co, num_kb * kTcBK, 1.f, 0.f, partial, per, bias, nullptr, 0, 0, cv, stats));
    else if (mode == 3) MSA_TRY(launch_tc<3>(grid, st, mA, mB, y, Co, M, Co, num_kb * kTcBK, 1.f, 0.f, partial, per, bias, nullptr, 0, 0, cv, stats));
    else MSA_TRY(launch_tc<1>(grid, st, mA, mB, y, Co, M, Co, num_kb * kTcBK, 1.f, 0.f, partial, per, bias, nullptr, 0, 0, cv, stats));
    if (splits > 1) {
        if (stats != nullptr) {
            ker_splitk_reduce_stats<<<dim3((unsigned)((M + 31) / 32), (unsigned)((Co + 31) / 32)), 256, 0, st>>>(partial, splits, M, Co, y, Co, 1.f, bias, stats);
        } else {
            const int64_t n = (int64_t)M * Co;
            ker_splitk_reduce<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(partial, splits, M, Co, y, Co, 1.f, 0.f, bias, nullptr);
        }
        MSA_LAUNCH_CHECK();
    }
    return 0;
}
// dx[B*Tn][Ci] = sum_k dy[b, t - k + pad, :] . Wp[k]   (the transposed convolution of the forward)
int conv_tc_dx(const float* dy, int B, int Tn, int Co, const float* wp, int Ci, int K, float* dx, int mode, float* scratch, cudaStream_t st) {
    CUtensorMap mA, mB;
    MSA_TRY(make_map_act(&mA, dy, B, Tn, Co, kTcBM, false));
    MSA_TRY(make_map(&mB, wp, Ci, K * Co, Ci, true));          // [K rows = (tap, co)][N = ci], N contiguous
    const int cb = (Co + kTcBK - 1) / kTcBK, tilesT = (Tn + kTcBM - 1) / kTcBM, num_kb = K * cb, M = B * Tn;
    const int splits = scratch != nullptr ? conv_plan_splits(B, Tn, Ci, num_kb) : 1;
    const int per = (num_kb + splits - 1) / splits;
    float* partial = splits > 1 ? scratch : nullptr;
    const TcConv cv{1, K, cb, (K - 1) / 2, -1, Tn, tilesT, Co, 1};
    const dim3 grid((unsigned)(B * tilesT), (unsigned)((Ci + kTcBN - 1) / kTcBN), (unsigned)splits);
    if (mode == 0) MSA_TRY(launch_tc<0>(grid, st, mA, mB, dx, Ci, M, Ci, num_kb * kTcBK, 1.f, 0.f, partial, per, nullptr, nullptr, 0, 1, cv, nullptr));
    else if (mode == 2) MSA_TRY(launch_tc<2>(grid, st, mA, mB, dx, Ci, M, Ci, num_kb * kTcBK, 1.f, 0.f, partial, per, nullptr, nullptr, 0, 1, cv, nullptr));
    else if (mode == 3) MSA_TRY(launch_tc<3>(grid, st, mA, mB, dx, Ci, M, Ci, num_kb * kTcBK, 1.f, 0.f, partial, per, nullptr, nullptr, 0, 1, cv, nullptr));
    else MSA_TRY(launch_tc<1>(grid, st, mA, mB, dx, Ci, M, Ci, num_kb * kTcBK, 1.f, 0.f, partial, per, nullptr, nullptr, 0, 1, cv, nullptr));
    if (splits > 1) {
        const int64_t n = (int64_t)M * Ci;
        ker_splitk_reduce<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(partial, splits, M, Ci, dx, Ci, 1.f, 0.f, nullptr, nullptr);
        MSA_LAUNCH_CHECK();
    }
    return 0;
}
// dW[Co][Ci][K] = (accumulate ? dW : 0) + scale * sum_{b,t} dy[b,t,co] x[b,t+k-pad,ci], native parameter layout
int conv_tc_dw(const float* dy, const float* x, int B, int Tn, int Co, int Ci, int K, float scale, int accumulate, float* dw, int mode,
               cudaStream_t st) {
    CUtensorMap mA, mB;
    MSA_TRY(make_map_act(&mA, dy, B, Tn, Co, kTcBK, true));
    MSA_TRY(make_map_act(&mB, x, B, Tn, Ci, kTcBK, true));
    const int tb = (Tn + kTcBK - 1) / kTcBK, num_kb = B * tb;
    const TcConv cv{2, K, tb, (K - 1) / 2, 1, Tn, 1, 0, K};
    const dim3 grid((unsigned)((Co + kTcBM - 1) / kTcBM), (unsigned)((Ci + kTcBN - 1) / kTcBN), (unsigned)K);
    const float beta = accumulate ? 1.f : 0.f;
    if (mode == 0) MSA_TRY(launch_tc<0>(grid, st, mA, mB, dw, Ci * K, Co, Ci, num_kb * kTcBK, scale, beta, nullptr, num_kb, nullptr, nullptr, 1, 1, cv, nullptr));
    else if (mode == 2) MSA_TRY(launch_tc<2>(grid, st, mA, mB, dw, Ci * K, Co, Ci, num_kb * kTcBK, scale, beta, nullptr, num_kb, nullptr, nullptr, 1, 1, cv, nullptr));
    else if (mode == 3) MSA_TRY(launch_tc<3>(grid, st, mA, mB, dw, Ci * K, Co, Ci, num_kb * kTcBK, scale, beta, nullptr, num_kb, nullptr, nullptr, 1, 1, cv, nullptr));
    else MSA_TRY(launch_tc<1>(grid, st, mA, mB, dw, Ci * K, Co, Ci, num_kb * kTcBK, scale, beta, nullptr, num_kb, nullptr, nullptr, 1, 1, cv, nullptr));
    return 0;
}
// Wp[k][co][ci] <- W[co][ci][k], n <= 8 layers in one launch
int conv_tc_repack(int n, const float* const* src, float* const* dst, const int* co, const int* ci, const int* k, cudaStream_t st) {
    MSA_CHECK(n >= 1 && n <= 8, MSA_E_ARG, "conv_tc_repack: %d layers", n);
    RepackTab tab{};
    int64_t big = 0;
    for (int i = 0; i < n; ++i) {
        tab.src[i] = src[i]; tab.dst[i] = dst[i]; tab.co[i] = co[i]; tab.ci[i] = ci[i]; tab.k[i] = k[i];
        MSA_CHECK(k[i] <= kRepackKmax, MSA_E_UNSUPPORTED, "conv_tc_repack: %d taps", k[i]);
        big = std::max<int64_t>(big, (int64_t)co[i] * ((ci[i] + kRepackCi - 1) / kRepackCi));
    }
    ker_conv_repack<<<dim3((unsigned)std::min<int64_t>(big, 148 * 2), (unsigned)n), 256, 0, st>>>(tab);
    MSA_LAUNCH_CHECK();
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

// mode 0: 3xTF32 (fp32-accurate); mode 1: single TF32 product (operands truncated by the tensor core); mode 2: single TF32 product
// with round-to-nearest operands (cuBLAS's TF32 accuracy); mode 3: the same when A already holds TF32-exact values (only B is rounded
// in shared memory: the kernel is bound by shared-memory traffic, the rounding pass of a tile costs as much as its TMA write + MMA read).  `scratch` (gemm_tc_scratch_floats(M, N, K) floats, 16-byte
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
    const TcConv cv{0, 0, 0, 0, 0, 0, 0, 0, 1};
    if (mode == 0) MSA_TRY(launch_tc<0>(grid, st, mA, mB, C, (int)ldc, (int)M, (int)N, (int)K, alpha, beta, partial, per, bias1, bias2, amn, bmn, cv, nullptr));
    else if (mode == 2) MSA_TRY(launch_tc<2>(grid, st, mA, mB, C, (int)ldc, (int)M, (int)N, (int)K, alpha, beta, partial, per, bias1, bias2, amn, bmn, cv, nullptr));
    else if (mode == 3) MSA_TRY(launch_tc<3>(grid, st, mA, mB, C, (int)ldc, (int)M, (int)N, (int)K, alpha, beta, partial, per, bias1, bias2, amn, bmn, cv, nullptr));
    else MSA_TRY(launch_tc<1>(grid, st, mA, mB, C, (int)ldc, (int)M, (int)N, (int)K, alpha, beta, partial, per, bias1, bias2, amn, bmn, cv, nullptr));
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

size_t msa_conv1d_scratch_floats(int B, int T, int Cin, int Cout, int K) {
    return std::max(msa::conv_tc_scratch_floats(B, T, Cin, Cout, K), msa::conv_tc_scratch_floats(B, T, Cout, Cin, K));
}
int msa_conv1d_stat_slabs(int B, int T, int Cin, int Cout, int K, int have_scratch) {
    return msa::conv_tc_stat_slabs(B, T, Cin, Cout, K, have_scratch != 0);
}
int msa_conv1d_repack(const float* w, float* wp, int Cout, int Cin, int K, void* stream) {
    MSA_CHECK(w && wp, MSA_E_ARG, "msa_conv1d_repack: null argument");
    const float* src[1] = {w};
    float* dst[1] = {wp};
    return msa::conv_tc_repack(1, src, dst, &Cout, &Cin, &K, (cudaStream_t)stream);
}
int msa_conv1d_fwd(const float* x, int B, int T, int Cin, const float* wp, int Cout, int K, const float* bias, float* y, int mode,
                   float* scratch, float* stats, void* stream) {
    MSA_CHECK(x && wp && y, MSA_E_ARG, "msa_conv1d_fwd: null argument");
    MSA_CHECK(msa::conv_tc_supported(B, T, Cin, Cout, K, x, wp), MSA_E_UNSUPPORTED, "msa_conv1d_fwd: channels must be multiples of 4, K odd, 16-byte aligned operands");
    return msa::conv_tc_fwd(x, B, T, Cin, wp, Cout, K, bias, y, mode, scratch, stats, (cudaStream_t)stream);
}
int msa_conv1d_dx(const float* dy, int B, int T, int Cout, const float* wp, int Cin, int K, float* dx, int mode, float* scratch, void* stream) {
    MSA_CHECK(dy && wp && dx, MSA_E_ARG, "msa_conv1d_dx: null argument");
    MSA_CHECK(msa::conv_tc_supported(B, T, Cin, Cout, K, dy, wp), MSA_E_UNSUPPORTED, "msa_conv1d_dx: channels must be multiples of 4, K odd, 16-byte aligned operands");
    return msa::conv_tc_dx(dy, B, T, Cout, wp, Cin, K, dx, mode, scratch, (cudaStream_t)stream);
}
int msa_conv1d_dw(const float* dy, const float* x, int B, int T, int Cout, int Cin, int K, float scale, int accumulate, float* dw, int mode,
                  void* stream) {
    MSA_CHECK(dy && x && dw, MSA_E_ARG, "msa_conv1d_dw: null argument");
    MSA_CHECK(msa::conv_tc_supported(B, T, Cin, Cout, K, dy, x), MSA_E_UNSUPPORTED, "msa_conv1d_dw: channels must be multiples of 4, K odd, 16-byte aligned operands");
    return msa::conv_tc_dw(dy, x, B, T, Cout, Cin, K, scale, accumulate, dw, mode, (cudaStream_t)stream);
}
}
