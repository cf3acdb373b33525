// Fused per-step kernels of free-running inference (Decoder.infer -> Decoder.decode, modules_tacotron2nv/decoder.py:234-274,
// 334-411; ForwardAttention.forward, forward_attn.py:178-219).
//
// One decoder step is SIX launches (prenet 1, prenet 2, attention LSTMCell, attention, decoder LSTMCell, projections + stop
// logic) instead of eight library GEMMs and seven glue kernels.  The step is sequential by nature (the next prenet input is
// this step's mel frame), its weights (80 MB fp32 at the default dimensions) do not fit in shared memory, and the batch is
// skinny (B <= 32 rows per tile), so every matrix product here is a "few rows of W per CTA, all of x" kernel:
//   ker_infer_rows : out[b][n] = sum_k W[n][k] x[b][k] for the 4*RPG rows of W a CTA owns (all 148 SMs stream their slice of W),
//                    K streamed through a multi-stage cp.async ring in shared memory (weights come from L2 after the first step:
//                    80 MB < 126 MB), the product on the tensor cores as 3 x TF32 mma.sync m16n8k8 with a hi/lo operand split
//                    (fp32-accurate; B <= 32 is too skinny for a tcgen05 tile per SM and the kernel is bound by the L2 -> SM
//                    stream, not by the tensor pipe), and the consumer of the product fused as the epilogue: relu + always-on
//                    dropout (prenet, decoder.py:9-20), the LSTMCell point-wise update (decoder.py:253-255,262-264), or bias +
//                    the stop-gate logic of the step run by the last CTA to finish (decoder.py:267-270,381-395).
//   ker_infer_attn : one CLUSTER of 4 CTAs per batch row: q = Wq.h_a', location conv + dense, energies, normalisation, cum/prev
//                    update and the context vector (forward_attn.py:121-131,200-219); q and the energies are exchanged through
//                    distributed shared memory.
// All kernels read the step index from device memory and return at once when the device-side done flag is set.
#include "common.cuh"
#include "kernels.h"

namespace msa {

constexpr int kIrThreads = 512;
constexpr int kIrBT = 32;               // batch rows per tile (blockIdx.y)
// K is streamed in chunks of KC floats through an NS-stage cp.async ring (16-byte copies, zero-filled beyond K and for missing
// rows); the shared-memory row stride is KC + 4 floats, so consecutive rows start 4 banks apart (conflict-free ldmatrix).  Every
// thread copies the same (row, 16-byte column) slots of every chunk, so its source pointers live in registers.  Measured on B200
// (profiles/r01_infer_notes.txt): recomputing them per copy cost twice the issue slots of the products themselves, and one 1-D
// bulk copy (TMA) per 512-byte row was slower still (~100 cycles per copy through the copy engine).
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 16 : 0;       // src-size 0: the 16 destination bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// tensor-core building blocks of the skinny products: ldmatrix (an 8x8 b16 matrix is 8 rows x 4 fp32 words), the hi/lo TF32 split
// and mma.sync m16n8k8.  Three products A_lo.B_hi + A_hi.B_lo + A_hi.B_hi with fp32 accumulation reproduce the fp32 product to
// ~2^-20 relative (a single TF32 product would spend the whole 1e-3 budget of the free-running loop).
__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], const float* smem_ptr) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// hi = the upper 19 bits of v (what the tensor core reads of a TF32 operand), lo = v - hi exactly; the tensor core truncates lo
// to its own 19 bits, so hi + lo carries ~21 mantissa bits.  Two instructions per value (cvt.rna.tf32 is a ~6-instruction
// emulation on sm_100a, measured: it dominated the issue slots of this kernel).
__device__ __forceinline__ void split_tf32(unsigned v, unsigned& hi, unsigned& lo) {
    hi = v & 0xffffe000u;
    lo = __float_as_uint(__uint_as_float(v) - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Programmatic dependent launch: every kernel of the decoder step is launched with programmatic stream serialization, lets its
// successor start launching at once (launch_dependents) and waits for its predecessor's memory before the first global read
// (wait): the launch latency / ramp of kernel n+1 overlaps the tail of kernel n.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ float dot4f(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __host__ __forceinline__ int ir_part_lo(int i, int n, int parts) { return (int)(((long long)i * n) / parts); }

// Per CTA: D[16*MT rows of W][32 batch rows] += W_tile . x_tile^T; warp w takes the k8 steps w, w+16, ... of every chunk (split-K
// over the 16 warps, summed through shared memory in the epilogue) and holds the whole 16*MT x 32 accumulator tile in registers.
// Fragments come from the padded row-major shared tiles by ldmatrix.x4 (conflict-free: consecutive rows start 4 banks apart).
template <int RPG, int kIrKC, int NS>
__global__ void __launch_bounds__(kIrThreads, 1) ker_infer_rows(InferRowsParams p) {
    pdl_prologue();
    if (p.state[1]) return;
    constexpr int R = 4 * RPG, MT = (R + 15) / 16, ROWS = kIrBT + 16 * MT, kIrLd = kIrKC + 4, KSTEPS = kIrKC / 8;
    static_assert(KSTEPS % (kIrThreads / 32) == 0, "every warp takes the same number of k8 steps per chunk");
    extern __shared__ __align__(16) float sm[];
    __shared__ const float* rowp[2][ROWS];
    __shared__ int grow_s[R];
    __shared__ int last_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = p.state[0];
    const int b0 = blockIdx.y * kIrBT, nb = min(kIrBT, p.B - b0);
    int u0 = 0, U = 0;
    if (p.epi == IR_EPI_LSTM) {
        u0 = ir_part_lo(blockIdx.x, p.H, gridDim.x);
        U = ir_part_lo(blockIdx.x + 1, p.H, gridDim.x) - u0;
    }
    // global row of every local row (-1: none) and the source row pointers of both K segments
    for (int rl = threadIdx.x; rl < R; rl += kIrThreads) {
        int gr;
        if (p.epi == IR_EPI_LSTM) {
            const int g = rl / RPG, ul = rl - g * RPG;
            gr = ul < U ? g * p.H + u0 + ul : -1;
        } else {
            gr = blockIdx.x * R + rl;
            if (gr >= p.N) gr = -1;
        }
        grow_s[rl] = gr;
        for (int s = 0; s < p.nseg; ++s) {
            const float* q = nullptr;
            if (gr >= 0) q = (s == 0 && p.nsplit > 0 && gr >= p.nsplit) ? p.Wb + (size_t)(gr - p.nsplit) * p.ldw[0] : p.W[s] + (size_t)gr * p.ldw[s];
            rowp[s][kIrBT + rl] = q;
        }
    }
    for (int r = threadIdx.x; r < kIrBT; r += kIrThreads)
        for (int s = 0; s < p.nseg; ++s) rowp[s][r] = r < nb ? p.x[s] + (size_t)(b0 + r) * p.ldx[s] : nullptr;
    for (int r = kIrBT + R + threadIdx.x; r < ROWS; r += kIrThreads)      // padding rows of the last 16-row tile
        for (int s = 0; s < 2; ++s) rowp[s][r] = nullptr;
    __syncthreads();

    const int nch0 = (p.K[0] + kIrKC - 1) / kIrKC, nch1 = p.nseg > 1 ? (p.K[1] + kIrKC - 1) / kIrKC : 0, nch_all = nch0 + nch1;
    // K split over blockIdx.z (few output rows, long K: the projection): this CTA takes the chunks [coff, coff + nch)
    const int coff = (int)(((long long)blockIdx.z * nch_all) / gridDim.z);
    const int nch = (int)(((long long)(blockIdx.z + 1) * nch_all) / gridDim.z) - coff;
    // every CTA walks the chunks in a different rotation (all CTAs read the same x rows: spreads those requests over the L2 slices)
    const int rot = (p.rotate && nch > 0) ? (int)((blockIdx.x * 7u) % (unsigned)nch) : 0;
    // copy slots of this thread: rows r0, r0 + RSTEP, ... at the 16-byte column c4
    constexpr int C4 = kIrKC / 4, RSTEP = kIrThreads / C4, NSLOT = (ROWS + RSTEP - 1) / RSTEP;
    const int c4 = threadIdx.x % C4, r0 = threadIdx.x / C4;
    const float* srcp[2][NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
        const int row = r0 + j * RSTEP;
        srcp[0][j] = (row < ROWS && rowp[0][row]) ? rowp[0][row] + c4 * 4 : nullptr;
        srcp[1][j] = (p.nseg > 1 && row < ROWS && rowp[1][row]) ? rowp[1][row] + c4 * 4 : nullptr;
    }
    auto issue = [&](int cc, int stage) {
        int c = cc + rot;
        if (c >= nch) c -= nch;
        c += coff;
        const int s = c < nch0 ? 0 : 1, k0 = (s ? c - nch0 : c) * kIrKC;
        const bool kin = k0 + c4 * 4 < p.K[s];
        float* dst = sm + ((size_t)stage * ROWS + r0) * kIrLd + c4 * 4;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
            if (r0 + j * RSTEP < ROWS) {
                const float* sp = s ? srcp[1][j] : srcp[0][j];
                const bool ok = kin && sp != nullptr;
                cp_async16(dst + (size_t)j * RSTEP * kIrLd, ok ? sp + k0 : p.x[0], ok);
            }
        }
        cp_async_commit();
    };

    // epilogue operands of the LSTM point-wise threads, fetched before the K loop (thread = (unit, batch row))
    float pre_b[4] = {0.f, 0.f, 0.f, 0.f}, pre_c = 0.f;
    if (p.epi == IR_EPI_LSTM && (int)threadIdx.x < RPG * 32) {
        const int ul = threadIdx.x >> 5, b = threadIdx.x & 31, u = u0 + ul;
        if (ul < U && b < nb) {
#pragma unroll
            for (int g = 0; g < 4; ++g) pre_b[g] = p.bias1[g * p.H + u] + p.bias2[g * p.H + u];
            pre_c = p.c[(size_t)(b0 + b) * p.H + u];
        }
    }
    float acc[MT][4][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
    // ldmatrix source rows / columns of this lane: matrix mi = lane >> 3, row lane & 7 of that matrix
    const int mi = lane >> 3, mr = lane & 7;
    const int a_off = (kIrBT + (mi & 1) * 8 + mr) * kIrLd + (mi >> 1) * 4;      // A: {rows 0-7, 8-15} x {k 0-3, 4-7}
    const int b_off = ((mi >> 1) * 8 + mr) * kIrLd + (mi & 1) * 4;              // B: {n-tile, n-tile + 1} x {k 0-3, 4-7}

    // ring: chunks c .. c+NS-2 are in flight while chunk c is consumed; one (possibly empty) commit group per iteration keeps the
    // wait_group count uniform
    for (int c = 0; c < NS - 1; ++c) {
        if (c < nch) issue(c, c);
        else cp_async_commit();
    }
    for (int c = 0; c < nch; ++c) {
        cp_async_wait<NS - 2>();
        __syncthreads();      // chunk c has landed
        if (c + NS - 1 < nch) issue(c + NS - 1, (c + NS - 1) % NS);
        else cp_async_commit();
        const int stage = c % NS;
        int kvalid;
        {
            int cr = c + rot;
            if (cr >= nch) cr -= nch;
            cr += coff;
            const int s = cr < nch0 ? 0 : 1;
            kvalid = min(kIrKC, p.K[s] - (s ? cr - nch0 : cr) * kIrKC);
        }
        const float* base = sm + (size_t)stage * ROWS * kIrLd;
#pragma unroll
        for (int ks = 0; ks < KSTEPS / (kIrThreads / 32); ++ks) {
            const int k0 = (ks * (kIrThreads / 32) + w) * 8;
            if (k0 >= kvalid) continue;              // (columns beyond K are zero-filled anyway)
            unsigned bh[4][2], bl[4][2];
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                unsigned r4[4];
                ldsm_x4(r4, base + b_off + (size_t)np * 16 * kIrLd + k0);
                split_tf32(r4[0], bh[2 * np][0], bl[2 * np][0]);
                split_tf32(r4[1], bh[2 * np][1], bl[2 * np][1]);
                split_tf32(r4[2], bh[2 * np + 1][0], bl[2 * np + 1][0]);
                split_tf32(r4[3], bh[2 * np + 1][1], bl[2 * np + 1][1]);
            }
            unsigned ah[MT][4], al[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                unsigned r4[4];
                ldsm_x4(r4, base + a_off + (size_t)m * 16 * kIrLd + k0);
#pragma unroll
                for (int i = 0; i < 4; ++i) split_tf32(r4[i], ah[m][i], al[m][i]);
            }
            // three passes over the 4*MT accumulator tiles (small terms first): consecutive mma.sync never share an accumulator
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) mma_tf32(acc[m][n], al[m], bh[n][0], bh[n][1]);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) mma_tf32(acc[m][n], ah[m], bl[n][0], bl[n][1]);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) mma_tf32(acc[m][n], ah[m], bh[n][0], bh[n][1]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // split-K partials -> shared: red[w][rl][b]; accumulator fragment: c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, ...)
    constexpr int RS = 40;      // row stride of the partial tiles: 64-bit stores of a half warp hit 32 distinct banks
    float* red = sm;
    {
        const int g = lane >> 2, tq = lane & 3;
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int rl = m * 16 + g + hh * 8, b = n * 8 + 2 * tq;
                    if (rl < R) *reinterpret_cast<float2*>(red + ((size_t)w * R + rl) * RS + b) = make_float2(acc[m][n][2 * hh], acc[m][n][2 * hh + 1]);
                }
    }
    __syncthreads();
    auto total = [&](int rl, int b) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < kIrThreads / 32; ++q) s += red[((size_t)q * R + rl) * RS + b];
        return s;
    };

    if (p.epi == IR_EPI_LSTM) {
        // LSTMCell point-wise for the owned units (eval: no dropout on h), gate order i, f, g, o
        if ((int)threadIdx.x < RPG * 32) {
            const int ul = threadIdx.x >> 5, b = threadIdx.x & 31, u = u0 + ul, H = p.H;
            if (ul < U && b < nb) {
                const float zi = total(0 * RPG + ul, b) + pre_b[0];
                const float zf = total(1 * RPG + ul, b) + pre_b[1];
                const float zg = total(2 * RPG + ul, b) + pre_b[2];
                const float zo = total(3 * RPG + ul, b) + pre_b[3];
                const float gi = sigmoidf_(zi), gf = sigmoidf_(zf), gg = tanhf(zg), go = sigmoidf_(zo);
                const size_t ci = (size_t)(b0 + b) * H + u;
                const float cn = gf * pre_c + gi * gg;
                p.c[ci] = cn;
                const float hv = go * tanhf(cn);
                p.h1[(size_t)(b0 + b) * p.ldh1 + u] = hv;
                if (p.h2) p.h2[(size_t)(b0 + b) * p.ldh2 + u] = hv;
            }
        }
        return;
    }
    for (int o = threadIdx.x; o < R * 32; o += kIrThreads) {
        const int rl = o >> 5, b = o & 31, gr = grow_s[rl];
        if (gr < 0 || b >= nb) continue;
        float v = total(rl, b);
        if (p.epi == IR_EPI_RELU_DROP) {
            // prenet layer: relu, then dropout with p = 0.5 that stays on in eval mode (decoder.py:17-19)
            v = fmaxf(v, 0.f);
            const uint8_t mk = p.mask[(((size_t)t * 2 + p.mask_layer) * p.B + b0 + b) * p.N + gr];
            p.out[(size_t)(b0 + b) * p.ldo + gr] = mk ? v * 2.f : 0.f;
        } else if (gridDim.z > 1) {
            p.part[((size_t)blockIdx.z * p.B + b0 + b) * p.N + gr] = v;      // K-split partial; summed in fixed order by the last CTA
        } else if (p.nsplit > 0 && gr >= p.nsplit) {
            p.out2[(size_t)(b0 + b) * p.ldo2 + gr - p.nsplit] = v + (p.bias2 ? p.bias2[gr - p.nsplit] : 0.f);
        } else {
            v += p.bias1 ? p.bias1[gr] : 0.f;
            p.out[(size_t)(b0 + b) * p.ldo + gr] = v;
            if (p.finish) {      // the mel frame goes straight to the output and to the next prenet input (off the serial tail)
                p.mel_tm[((size_t)t * p.B + b0 + b) * p.M + gr] = v;
                p.frame[(size_t)(b0 + b) * p.M + gr] = v;
            }
        }
    }
    if (!p.finish) return;
    // ---- end of the step, run by the last CTA to arrive: mel frame -> output and next prenet input; stop-gate logic ----
    // (decoder.py:381-395): dec = sigmoid(gate) <= threshold; not_finished *= dec; mel_lengths += not_finished
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_s = atomicAdd(p.counter, 1u) == gridDim.x * gridDim.y * gridDim.z - 1 ? 1 : 0;
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int BM = p.B * p.M;
    if (gridDim.z > 1) {
        // sum the K-split partials (fixed order: deterministic) + bias -> mel frame [B][M] and gate [B]
        for (int i = threadIdx.x; i < p.B * p.N; i += kIrThreads) {
            const int b = i / p.N, gr = i - b * p.N;
            float v = 0.f;
            for (int z = 0; z < (int)gridDim.z; ++z) v += __ldcg(p.part + ((size_t)z * p.B + b) * p.N + gr);
            if (p.nsplit > 0 && gr >= p.nsplit) p.out2[(size_t)b * p.ldo2 + gr - p.nsplit] = v + (p.bias2 ? p.bias2[gr - p.nsplit] : 0.f);
            else p.out[(size_t)b * p.ldo + gr] = v + (p.bias1 ? p.bias1[gr] : 0.f);
        }
        __syncthreads();
    }
    if (gridDim.z > 1)
        for (int i = threadIdx.x; i < BM; i += kIrThreads) {
            const float v = __ldcg(p.out + i);
            p.mel_tm[(size_t)t * BM + i] = v;
            p.frame[i] = v;
        }
    for (int b = threadIdx.x; b < p.B; b += kIrThreads) {
        const float g = __ldcg(p.out2 + b);
        const int dec = (1.f / (1.f + expf(-g))) <= p.threshold ? 1 : 0;
        const int nf = p.not_finished[b] * dec;
        p.not_finished[b] = nf;
        p.mel_lengths[b] += nf;
        if (nf) atomicOr(&any, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *p.counter = 0u;
        p.state_rw[2] = t + 1;
        if ((p.early && !any) || t + 1 >= p.max_steps) p.state_rw[1] = 1;
        p.state_rw[0] = t + 1;
    }
}

template <int RPG, int KC, int NS>
static int launch_rows(const InferRowsParams& p, int gx, cudaStream_t st) {
    constexpr size_t ring = sizeof(float) * (size_t)NS * (kIrBT + 16 * ((4 * RPG + 15) / 16)) * (KC + 4);
    constexpr size_t red = sizeof(float) * (size_t)(kIrThreads / 32) * 4 * RPG * 40;
    constexpr size_t smem = ring > red ? ring : red;
    static_assert(smem <= 227 * 1024, "infer_rows: ring does not fit in shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        MSA_CUDA(cudaFuncSetAttribute(ker_infer_rows<RPG, KC, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // the same (maximal) shared-memory carve-out for every kernel of the step: no L1 / shared reconfiguration between launches
        MSA_CUDA(cudaFuncSetAttribute(ker_infer_rows<RPG, KC, NS>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(gx, (p.B + kIrBT - 1) / kIrBT, p.ksplit > 1 ? p.ksplit : 1);
    cfg.blockDim = dim3(kIrThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_rows<RPG, KC, NS>, p));
    MSA_LAUNCH_CHECK();
    return 0;
}
static int ir_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MSA_IR_CFG");
        v = e ? atoi(e) : 0;      // 1: 128-column chunks, 4 stages (measured equal within 2 % on B200)
    }
    return v;
}

int k_infer_rows(const InferRowsParams& p_in, int sm_count, cudaStream_t st) {
    InferRowsParams p = p_in;
    static int rot = -1;
    if (rot < 0) rot = getenv("MSA_IR_NOROT") ? 0 : 1;
    p.rotate = rot;
    MSA_CHECK(p.nseg >= 1 && p.nseg <= 2, MSA_E_ARG, "infer_rows: %d K segments", p.nseg);
    for (int s = 0; s < p.nseg; ++s) {
        MSA_CHECK(p.K[s] >= 4 && p.K[s] % 4 == 0 && p.ldx[s] % 4 == 0 && p.ldw[s] % 4 == 0 && ((uintptr_t)p.x[s] & 15) == 0 && ((uintptr_t)p.W[s] & 15) == 0,
                  MSA_E_UNSUPPORTED, "infer_rows: segment %d (K=%d, ldx=%d, ldw=%d) is not 16-byte aligned", s, p.K[s], p.ldx[s], p.ldw[s]);
    }
    MSA_CHECK(p.nsplit == 0 || ((uintptr_t)p.Wb & 15) == 0, MSA_E_UNSUPPORTED, "infer_rows: second weight block is not 16-byte aligned");
    MSA_CHECK(p.ksplit <= 1 || (p.finish && p.part && p.epi == IR_EPI_BIAS), MSA_E_ARG, "infer_rows: K split needs the finishing CTA and a partial buffer");
    if (p.epi == IR_EPI_LSTM) {
        // units split over min(#SMs, H) CTAs when that leaves <= 7 units per CTA, else one CTA per 7 units
        int gx = p.H < sm_count ? p.H : sm_count;
        if ((p.H + gx - 1) / gx > 7) gx = (p.H + 6) / 7;
        return ir_variant() == 1 ? launch_rows<7, 128, 4>(p, gx, st) : launch_rows<7, 256, 3>(p, gx, st);
    }
    // few output rows: 4 per CTA spreads them over more SMs (each CTA streams all of x either way)
    if (p.N <= 4 * 7 * 8) {
        return launch_rows<1, 128, 8>(p, (p.N + 3) / 4, st);
    }
    return launch_rows<7, 256, 3>(p, (p.N + 27) / 28, st);
}

// =====================================================================================================================
// attention: one CLUSTER of kIaCl CTAs per batch row (B = 32 rows alone would leave 116 of the 148 SMs idle and every
// phase four times longer).  CTA r of the cluster owns the attention dims d = r (mod kIaCl) of the query projection, the
// text positions [r*Lc, (r+1)*Lc) of the location features / energies and the memory channels [r*Ec, (r+1)*Ec) of the
// context vector; q and the energies are all-gathered by remote shared-memory stores (DSMEM) + a cluster barrier, the
// normalisation is recomputed by every CTA.
constexpr int kIaThreads = 512;
constexpr int kIaCl = 4;
constexpr int kIaDJ = 4;          // attention dims per lane in the energy phase: A <= 32 * kIaDJ

struct IaSmem {
    int LHp, FS, FP4, Lc, LcP;
    size_t in_s, h_s, q_s, v_s, wloc_s, wld4_s, cf_s, e_s, part_s, alpha_s, anew_s, ta_s, mem_s, total;
};
// mem_res: the CTA's [L][E/4] tile of the encoder memory is prefetched into shared memory at kernel start
__host__ __device__ inline bool ia_mem_tile_ok(int E) { return E % 4 == 0 && ((E + kIaCl - 1) / kIaCl) % 4 == 0; }
__host__ __device__ inline IaSmem ia_layout(int L, int Ha, int A, int F, int Kl, int E, bool mem_res = false) {
    IaSmem s;
    s.Lc = (((L + kIaCl - 1) / kIaCl) + 3) & ~3;          // owned positions per CTA, multiple of 4
    s.LcP = s.Lc + 4;
    s.LHp = ((s.Lc * kIaCl + Kl - 1 + 8) + 3) & ~3;
    s.FP4 = (F + 3) >> 2;
    s.FS = s.FP4 * 4;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.in_s = take((size_t)2 * s.LHp);
    s.h_s = take(Ha);
    s.q_s = take(A);
    s.v_s = take(A);
    s.wloc_s = take((size_t)2 * Kl * F);
    s.wld4_s = take((size_t)s.FP4 * A * 4);
    s.cf_s = take((size_t)s.LcP * s.FS);
    s.e_s = take((size_t)s.Lc * kIaCl);
    s.part_s = take((size_t)8 * (((E + kIaCl - 1) / kIaCl + 3) & ~3));
    s.alpha_s = take((size_t)s.Lc * kIaCl + 4);      // forward attention: alpha(t-1) with a zero in front (the shifted copy)
    s.anew_s = take((size_t)s.Lc * kIaCl);
    s.ta_s = take(64);
    s.mem_s = take(mem_res ? (size_t)L * ((E + kIaCl - 1) / kIaCl) : 0);
    s.total = o;
    return s;
}

__device__ __forceinline__ unsigned int cluster_ctarank_() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v at the same shared-memory offset as local_ptr in CTA `rank` of the cluster
__device__ __forceinline__ void st_cluster(float* local_ptr, unsigned int rank, float v) {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(local_ptr);
    unsigned int ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kIaThreads, 1) ker_infer_attn(InferAttnParams p) {
    // no early return before the cluster barriers: every CTA of a cluster takes the same path (the flag is launch-uniform)
    pdl_prologue();
    if (p.state[1]) return;
    extern __shared__ __align__(16) float sm[];
    const int t = p.state[0], b = blockIdx.x / kIaCl;
    const int r = (int)cluster_ctarank_();
    const int L = p.L, A = p.A, F = p.F, Kl = p.Kl, Ha = p.Ha, E = p.E, pl = (Kl - 1) / 2;
    const IaSmem lay = ia_layout(L, Ha, A, F, Kl, E, p.mem_res != 0);
    const int LHp = lay.LHp, FS = lay.FS, FP4 = lay.FP4, Lc = lay.Lc;
    const int lbeg = r * Lc, lend = min(L, lbeg + Lc);      // owned positions
    float* in_s = sm + lay.in_s;        // [2][LHp]      prev / cum of the whole row with a zero halo of pl in front
    float* h_s = sm + lay.h_s;          // [Ha]          h_a'(t) of this row
    float* q_s = sm + lay.q_s;          // [A]           all-gathered
    float* v_s = sm + lay.v_s;
    float* wloc_s = sm + lay.wloc_s;    // [2*Kl][F]     wloc[f][c][k] -> [c*Kl+k][f]
    float* wld4_s = sm + lay.wld4_s;    // [FP4][A][4]   wld[d][f] -> [f/4][d][f%4], zero beyond F
    float* cf_s = sm + lay.cf_s;        // [LcP][FS]     location features of the owned positions, zero beyond lend / F
    float* e_s = sm + lay.e_s;          // [kIaCl*Lc]    energies, all-gathered; then the alignment
    float* part_s = sm + lay.part_s;    // [8][Ec]       partial context sums
    float* alpha_s = sm + lay.alpha_s + 1;   // [-1..L)    alpha(t-1), alpha_s[-1] = 0
    float* anew_s = sm + lay.anew_s;
    float* ta_s = sm + lay.ta_s;        // [kIaCl] partial transition-agent dots (rank 0), [32..] block-reduction scratch
    float* mem_s = sm + lay.mem_s;      // [L][Ec]       memory[b][:, owned channels] (prefetched)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = kIaThreads / 32;

    // Everything this CTA reads that does not depend on its own progress is requested up front as four cp.async groups, in the
    // order of first use: h_a' (query projection), location conv weights, location dense weights, the memory tile of the context
    // phase.  The phases wait for their group only; the loads overlap the phases before them.
    const int Ec = (E + kIaCl - 1) / kIaCl, e0 = r * Ec, ne = max(0, min(Ec, E - e0));
    for (int i = threadIdx.x; i < (Ha >> 2); i += kIaThreads) cp_async16(h_s + i * 4, p.h + (size_t)b * p.ldh + i * 4, true);
    cp_async_commit();
    for (int i = threadIdx.x; i < (F * 2 * Kl) >> 2; i += kIaThreads) cp_async16(wloc_s + i * 4, p.wloc_t + i * 4, true);
    cp_async_commit();
    for (int i = threadIdx.x; i < FP4 * A; i += kIaThreads) cp_async16(wld4_s + i * 4, p.wld4 + i * 4, true);
    cp_async_commit();
    if (p.mem_res) {
        const int E4 = Ec >> 2;
        for (int i = threadIdx.x; i < L * E4; i += kIaThreads) {
            const int l = i / E4, c4 = i - l * E4;
            const bool ok = c4 * 4 < ne;
            cp_async16(mem_s + (size_t)l * Ec + c4 * 4, ok ? p.memory + ((size_t)b * L + l) * E + e0 + c4 * 4 : p.memory, ok);
        }
    }
    cp_async_commit();
    for (int i = threadIdx.x; i < 2 * LHp; i += kIaThreads) {
        const int c = i / LHp, l = i - c * LHp - pl;
        in_s[i] = (l >= 0 && l < L) ? (c == 0 ? p.prev[(size_t)b * L + l] : p.cum[(size_t)b * L + l]) : 0.f;
    }
    for (int i = threadIdx.x; i < A; i += kIaThreads) v_s[i] = __ldg(p.v + i);
    for (int i = ((F * 2 * Kl) & ~3) + threadIdx.x; i < F * 2 * Kl; i += kIaThreads) wloc_s[i] = __ldg(p.wloc_t + i);
    for (int i = threadIdx.x; i < lay.LcP * FS; i += kIaThreads) cf_s[i] = 0.f;
    if (p.forward_attn)
        for (int i = threadIdx.x; i <= L; i += kIaThreads) alpha_s[i - 1] = i == 0 ? 0.f : p.alpha[(size_t)b * L + i - 1];
    cp_async_wait<2>();      // h_a' and the location conv weights have landed (this thread's copies; the barrier covers the rest)
    cluster_sync_();         // staging done; every CTA of the cluster is running (required before remote shared-memory stores)

    // ---- query projection (forward_attn.py:125) for the dims d = r (mod kIaCl): two dims per warp pass, 8 loads in flight ----
    const int Ha4 = Ha >> 2;
    for (int di = w; di * kIaCl + r < A; di += 2 * NW) {
        const int d0 = di * kIaCl + r, d1 = (di + NW) * kIaCl + r;
        const bool two = d1 < A;
        const float4* w0 = reinterpret_cast<const float4*>(p.wq + (size_t)d0 * Ha);
        const float4* w1 = reinterpret_cast<const float4*>(p.wq + (size_t)(two ? d1 : d0) * Ha);
        float a0 = 0.f, a1 = 0.f;
        for (int k4 = lane; k4 < Ha4; k4 += 256) {      // 16 independent 128-bit loads in flight per lane
            float4 r0[8], r1[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int idx = k4 + 32 * m;
                r0[m] = idx < Ha4 ? __ldg(w0 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                r1[m] = (two && idx < Ha4) ? __ldg(w1 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int idx = k4 + 32 * m;
                if (idx < Ha4) {
                    const float4 hv = reinterpret_cast<const float4*>(h_s)[idx];
                    a0 += dot4f(r0[m], hv);
                    a1 += dot4f(r1[m], hv);
                }
            }
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane < kIaCl) {
            st_cluster(q_s + d0, lane, a0);
            if (two) st_cluster(q_s + d1, lane, a1);
        }
    }
    // ---- location conv (forward_attn.py:46-50) for the owned positions: thread = (4 consecutive positions, filter) ----
    for (int item = threadIdx.x; item < (Lc >> 2) * F; item += kIaThreads) {
        const int f = item % F, lo = (item / F) * 4, l0 = lbeg + lo;
        if (l0 >= L) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < 2; ++c) {
            const float* in = in_s + c * LHp + l0;        // in[j + k] = input position l0 + j + k - pl
            float win[4];
#pragma unroll
            for (int j = 0; j < 3; ++j) win[j] = in[j];
            for (int k = 0; k < Kl; ++k) {
                win[3] = in[k + 3];
                const float wk = wloc_s[(c * Kl + k) * F + f];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] += wk * win[j];
#pragma unroll
                for (int j = 0; j < 3; ++j) win[j] = win[j + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (l0 + j < L) cf_s[(lo + j) * FS + f] = acc[j];
    }
    cp_async_wait<1>();      // location dense weights
    cluster_sync_();         // q complete in every CTA (also a CTA barrier: cf_s complete)
    // ---- energies (forward_attn.py:126-131) for the owned positions: warp = 2 positions, lane = attention dims lane + 32 j ----
    const float bv = __ldg(p.bv);
    for (int lo = w * 2; lbeg + lo < lend; lo += NW * 2) {
        const int l0 = lbeg + lo;
        float pmv[2][kIaDJ];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < kIaDJ; ++j) {
                const int d = lane + 32 * j;
                pmv[i][j] = (l0 + i < L && d < A) ? __ldg(p.pm + ((size_t)b * L + l0 + i) * A + d) : 0.f;
            }
        float loc[2][kIaDJ];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < kIaDJ; ++j) loc[i][j] = 0.f;
        for (int fc = 0; fc < FP4; ++fc) {
            float4 cf[2], wl[kIaDJ];
#pragma unroll
            for (int i = 0; i < 2; ++i) cf[i] = *reinterpret_cast<const float4*>(cf_s + (lo + i) * FS + fc * 4);
#pragma unroll
            for (int j = 0; j < kIaDJ; ++j) {
                const int d = lane + 32 * j;
                wl[j] = d < A ? *reinterpret_cast<const float4*>(wld4_s + ((size_t)fc * A + d) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < kIaDJ; ++j) loc[i][j] += dot4f(cf[i], wl[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float e = 0.f;
#pragma unroll
            for (int j = 0; j < kIaDJ; ++j) {
                const int d = lane + 32 * j;
                if (d < A) e += v_s[d] * tanhf(q_s[d] + loc[i][j] + pmv[i][j]);
            }
            e = warp_sum(e);
            if (lane < kIaCl && l0 + i < L) st_cluster(e_s + l0 + i, lane, e + bv);
        }
    }
    cluster_sync_();         // all energies of the row in every CTA; no remote store after this point
    // ---- windowing, normalisation, forward attention (forward_attn.py:139-176,200-219): recomputed by warp 0 of every CTA of
    // the cluster; the owner of a position writes the state (cum += a, prev = a', alpha = a') ----
    if (w == 0) {
        if (p.windowing) {
            // eval only: the window follows the argmax of BATCH ROW 0 of the previous step for every row
            const int win = p.win[t & 1], back = win - 2, front = win + 6;
            float m = -INFINITY;
            for (int l = lane; l < L; l += 32) {
                float e = e_s[l];
                if ((back > 0 && l < back) || (front < L && l >= front)) e = -INFINITY;
                e_s[l] = e;
                m = fmaxf(m, e);
            }
            m = warp_max(m);
            if (p.phase == 1) {
                // first step (win = -1): attention[:, 0] = attention.max() is a maximum over the WHOLE batch: this launch only
                // contributes the row maximum, the step is then run again with phase 2
                if (r == 0 && lane == 0) {
                    if (m >= 0.f) atomicMax(reinterpret_cast<int*>(p.gmax), __float_as_int(m));
                    else atomicMin(reinterpret_cast<unsigned int*>(p.gmax), __float_as_uint(m));
                }
            } else {
                __syncwarp();
                if (win == -1 && lane == 0) e_s[0] = *reinterpret_cast<volatile float*>(p.gmax);
                __syncwarp();
                if (b == 0 && r == 0) {
                    // argmax(attention, 1)[0]: first index of the maximum
                    float bm = -INFINITY;
                    int bi = 0x7fffffff;
                    for (int l = lane; l < L; l += 32)
                        if (e_s[l] > bm) { bm = e_s[l]; bi = l; }
                    for (int o = 16; o > 0; o >>= 1) {
                        const float om = __shfl_xor_sync(0xffffffffu, bm, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
                    }
                    if (lane == 0) p.win[(t + 1) & 1] = bi == 0x7fffffff ? 0 : bi;
                }
            }
        }
        if (p.phase != 1) {
            float m = 0.f;
            if (p.norm == 0) {
                m = -INFINITY;
                for (int l = lane; l < L; l += 32) m = fmaxf(m, e_s[l]);
                m = warp_max(m);
            }
            float s = 0.f;
            for (int l = lane; l < L; l += 32) {
                const float x = p.norm == 0 ? expf(e_s[l] - m) : sigmoidf_(e_s[l]);
                e_s[l] = x;
                s += x;
            }
            s = warp_sum(s);
            for (int l = lane; l < L; l += 32) {
                const float a = e_s[l] / s;
                e_s[l] = a;
                if (l >= lbeg && l < lend) p.cum[(size_t)b * L + l] += a;      // the location features see the plain alignment
            }
            if (p.forward_attn) {
                // alpha' = ((1-u) alpha + u shift(alpha) + 1e-8) * a, renormalised (forward_attn.py:154-176)
                const float u = p.u[b];
                float vmax = -INFINITY, smax = -INFINITY;
                int sidx = 0x7fffffff;
                __syncwarp();
                for (int l = lane; l < L; l += 32) {
                    const float sh = alpha_s[l - 1];
                    const float an = (((1.f - u) * alpha_s[l] + u * sh) + 1e-8f) * e_s[l];
                    anew_s[l] = an;
                    vmax = fmaxf(vmax, an);
                    if (sh > smax) { smax = sh; sidx = l; }
                }
                if (p.forward_attn_mask) {
                    // eval only: keep the states around n = argmax(shifted alpha), with the reference's Python slice /
                    // negative-index semantics: alpha[n+3:] = 0; alpha[:n-1] = 0; alpha[n-2] = 0.01 max(alpha')
                    vmax = warp_max(vmax);
                    for (int o = 16; o > 0; o >>= 1) {
                        const float om = __shfl_xor_sync(0xffffffffu, smax, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, sidx, o);
                        if (om > smax || (om == smax && oi < sidx)) { smax = om; sidx = oi; }
                    }
                    const int n = sidx == 0x7fffffff ? 0 : sidx;
                    const int stop = n - 1 >= 0 ? n - 1 : max(L + n - 1, 0);
                    int idx = n - 2;
                    if (idx < 0) idx += L;
                    __syncwarp();
                    for (int l = lane; l < L; l += 32) {
                        float an = anew_s[l];
                        if (l >= n + 3 || l < stop) an = 0.f;
                        if (l == idx) an = 0.01f * vmax;
                        anew_s[l] = an;
                    }
                }
                __syncwarp();
                float sa = 0.f;
                for (int l = lane; l < L; l += 32) sa += anew_s[l];
                sa = warp_sum(sa);
                for (int l = lane; l < L; l += 32) {
                    const float a = anew_s[l] / sa;
                    e_s[l] = a;
                    if (l >= lbeg && l < lend) p.alpha[(size_t)b * L + l] = a;
                }
            }
            for (int l = lane; l < L; l += 32)
                if (l >= lbeg && l < lend) {
                    p.prev[(size_t)b * L + l] = e_s[l];
                    p.align_out[((size_t)b * p.max_steps + t) * L + l] = e_s[l];
                }
        }
    }
    cp_async_wait<0>();      // memory tile (also before the early return: no copy may land after the CTA has exited)
    if (p.phase == 1) return;      // launch-uniform
    __syncthreads();
    // ---- context (forward_attn.py:217) for the owned memory channels: thread = (channel, one of 8 interleaved position sets) ----
    const int EcP = (Ec + 3) & ~3;
    for (int it = threadIdx.x; it < 8 * EcP; it += kIaThreads) {
        const int ls = it / EcP, ec = it - ls * EcP;
        float acc = 0.f;
        if (ec < ne && p.mem_res) {
            float c0 = 0.f, c1 = 0.f;
            int l = ls;
            for (; l + 8 < L; l += 16) {
                c0 += e_s[l] * mem_s[(size_t)l * Ec + ec];
                c1 += e_s[l + 8] * mem_s[(size_t)(l + 8) * Ec + ec];
            }
            if (l < L) c0 += e_s[l] * mem_s[(size_t)l * Ec + ec];
            acc = c0 + c1;
        } else if (ec < ne) {
            const float* mrow = p.memory + (size_t)b * L * E + e0 + ec;
            // positions ls, ls + 8, ls + 16, ...: eight independent loads in flight per pass
            float cacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cacc[j] = 0.f;
            for (int l0 = ls; l0 < L; l0 += 64) {
                float mv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) mv[j] = l0 + 8 * j < L ? __ldg(mrow + (size_t)(l0 + 8 * j) * E) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (l0 + 8 * j < L) cacc[j] += e_s[l0 + 8 * j] * mv[j];
            }
            acc = ((cacc[0] + cacc[1]) + (cacc[2] + cacc[3])) + ((cacc[4] + cacc[5]) + (cacc[6] + cacc[7]));
        }
        part_s[it] = acc;
    }
    __syncthreads();
    for (int ec = threadIdx.x; ec < ne; ec += kIaThreads) {
        float acc = 0.f;
#pragma unroll
        for (int ls = 0; ls < 8; ++ls) acc += part_s[ls * EcP + ec];
        const int e = e0 + ec;
        p.ctx1[(size_t)b * p.ld1 + e] = acc;
        p.ctx2[(size_t)b * p.ld2 + e] = acc;
        p.ctx3[(size_t)b * p.ld3 + e] = acc;
    }
    if (p.trans_agent) {
        // u = sigmoid(W_ta . [ctx; h_a'] + b_ta) (forward_attn.py:222-224): every CTA dots its context channels and its quarter of
        // the query, rank 0 adds the four partials
        float acc = 0.f;
        for (int ec = threadIdx.x; ec < ne; ec += kIaThreads) {
            float c = 0.f;
#pragma unroll
            for (int ls = 0; ls < 8; ++ls) c += part_s[ls * EcP + ec];
            acc += __ldg(p.wta + e0 + ec) * c;
        }
        const int Hq = (Ha + kIaCl - 1) / kIaCl, k0 = r * Hq, k1 = min(Ha, k0 + Hq);
        for (int k = k0 + threadIdx.x; k < k1; k += kIaThreads) acc += __ldg(p.wta + E + k) * h_s[k];
        acc = block_sum(acc, ta_s + 16);
        if (threadIdx.x == 0) st_cluster(ta_s + r, 0, acc);
        cluster_sync_();
        if (r == 0 && threadIdx.x == 0) {
            float z = __ldg(p.bta);
            for (int i = 0; i < kIaCl; ++i) z += ta_s[i];
            p.u[b] = sigmoidf_(z);
        }
    }
}

// wloc [F][2][Kl] -> wloc_t [2*Kl][F];  wld [A][F] -> wld4 [ceil(F/4)][A][4] (zero beyond F): the layouts ker_infer_attn keeps in
// shared memory, written once per msa_infer call
__global__ void ker_infer_attn_prep(const float* __restrict__ wloc, const float* __restrict__ wld, float* wloc_t, float* wld4, int F,
                                    int Kl, int A) {
    const int FP4 = (F + 3) >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F * 2 * Kl; i += gridDim.x * blockDim.x) {
        const int f = i / (2 * Kl), ck = i - f * (2 * Kl);
        wloc_t[ck * F + f] = wloc[i];
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < FP4 * A * 4; i += gridDim.x * blockDim.x) {
        const int fq = i & 3, d = (i >> 2) % A, fc = (i >> 2) / A, f = fc * 4 + fq;
        wld4[i] = f < F ? wld[(size_t)d * F + f] : 0.f;
    }
}
int k_infer_attn_prep(const float* wloc, const float* wld, float* wloc_t, float* wld4, int F, int Kl, int A, cudaStream_t st) {
    ker_infer_attn_prep<<<8, 256, 0, st>>>(wloc, wld, wloc_t, wld4, F, Kl, A);
    MSA_LAUNCH_CHECK();
    return 0;
}

// =====================================================================================================================
// Both prenet layers (decoder.py:9-20,366) in ONE launch: a cluster of 8 CTAs per batch tile, CTA r owns the output columns
// [r*Pc, (r+1)*Pc) of both layers.  Layer 1 = relu(W1 . frame) * mask on the tensor cores (3xTF32), its [32][Pc] slice is written
// into the layer-2 input tile of ALL eight CTAs through distributed shared memory, one cluster barrier, layer 2 the same way from
// shared memory, result into the packed attention-LSTM input.  The weights (42 KB per CTA) are requested before the
// programmatic-dependency wait.  Two separate launches of the generic rows kernel cost ~9 us each for < 1 us of work.
constexpr int kPnThreads = 256;
constexpr int kPnCl = 8;
struct PnSmem {
    int S1, S2, Pc;
    size_t x1, w1, p1, w2, total;
};
__host__ __device__ inline PnSmem pn_layout(int M, int Pd) {
    PnSmem s;
    s.S1 = ((M + 15) & ~15) + 4;
    s.S2 = ((Pd + 15) & ~15) + 4;
    s.Pc = (Pd + kPnCl - 1) / kPnCl;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.x1 = take((size_t)32 * s.S1);
    s.w1 = take((size_t)32 * s.S1);
    s.p1 = take((size_t)32 * s.S2);
    s.w2 = take((size_t)32 * s.S2);
    s.total = o;
    return s;
}
// one warp = one 16 x 8 tile of D[weight row][batch row] = sum_k W[row][k] x[b][k], two k8 steps per iteration
__device__ __forceinline__ void pn_tile(float (&acc)[4], const float* Ws, const float* Xs, int S, int K16, int mt, int nt, int lane) {
    const int mi = lane >> 3, mr = lane & 7;
    const float* ap = Ws + (size_t)(mt * 16 + (mi & 1) * 8 + mr) * S + (mi >> 1) * 4;      // {rows 0-7, 8-15} x {k 0-3, 4-7}
    const float* bp = Xs + (size_t)(nt * 8 + mr) * S + mi * 4;                            // k 0-3, 4-7, 8-11, 12-15
    for (int k = 0; k < K16; k += 16) {
        unsigned b4[4], a4[4], ah[4], al[4], bh[4], bl[4];
        ldsm_x4(b4, bp + k);
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(b4[i], bh[i], bl[i]);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            ldsm_x4(a4, ap + k + half * 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(a4[i], ah[i], al[i]);
            mma_tf32(acc, al, bh[2 * half], bh[2 * half + 1]);
            mma_tf32(acc, ah, bl[2 * half], bl[2 * half + 1]);
            mma_tf32(acc, ah, bh[2 * half], bh[2 * half + 1]);
        }
    }
}
__global__ void __launch_bounds__(kPnThreads, 1) ker_infer_prenet(InferPrenetParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) float sm[];
    const int M = p.M, Pd = p.Pd;
    const PnSmem lay = pn_layout(M, Pd);
    const int S1 = lay.S1, S2 = lay.S2, Pc = lay.Pc;
    float* X1 = sm + lay.x1;      // [32][S1] frame rows (zero beyond B / M)
    float* W1 = sm + lay.w1;      // [32][S1] rows of W1 owned by this CTA (zero beyond Pc / M)
    float* P1 = sm + lay.p1;      // [32][S2] layer-1 output of the whole cluster (zero beyond Pd)
    float* W2 = sm + lay.w2;      // [32][S2] rows of W2 owned by this CTA
    const int r = (int)cluster_ctarank_(), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b0 = blockIdx.y * 32, nb = min(32, p.B - b0);
    const int c0 = r * Pc, nc = max(0, min(Pc, Pd - c0));
    const int M4 = (S1 - 4) >> 2, P4 = (S2 - 4) >> 2;
    // weights first (independent of the previous kernel), zero-filled beyond the owned rows / K
    for (int i = threadIdx.x; i < 32 * M4; i += kPnThreads) {
        const int row = i / M4, c4 = i - row * M4;
        const bool ok = row < nc && c4 * 4 < M;
        cp_async16(W1 + (size_t)row * S1 + c4 * 4, ok ? p.w1 + (size_t)(c0 + row) * M + c4 * 4 : p.w1, ok);
    }
    for (int i = threadIdx.x; i < 32 * P4; i += kPnThreads) {
        const int row = i / P4, c4 = i - row * P4;
        const bool ok = row < nc && c4 * 4 < Pd;
        cp_async16(W2 + (size_t)row * S2 + c4 * 4, ok ? p.w2 + (size_t)(c0 + row) * Pd + c4 * 4 : p.w2, ok);
    }
    cp_async_commit();
    for (int i = threadIdx.x; i < 32 * S2; i += kPnThreads) P1[i] = 0.f;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool done = p.state[1] != 0;      // launch-uniform; no early return before the cluster barriers / pending copies
    const int t = p.state[0];
    for (int i = threadIdx.x; i < 32 * M4; i += kPnThreads) {
        const int row = i / M4, c4 = i - row * M4;
        const bool ok = !done && row < nb && c4 * 4 < M;
        cp_async16(X1 + (size_t)row * S1 + c4 * 4, ok ? p.frame + (size_t)(b0 + row) * M + c4 * 4 : p.frame, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    cluster_sync_();          // tiles complete; every CTA of the cluster is running (remote stores below)
    if (done) return;
    const int mt = w & 1, nt = w >> 1, g = lane >> 2, tq = lane & 3;      // 8 warps = 2 x 4 output tiles
    // ---- layer 1 ----
    {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        pn_tile(acc, W1, X1, S1, S1 - 4, mt, nt, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int col = mt * 16 + g + (i >> 1) * 8, b = nt * 8 + 2 * tq + (i & 1);      // local column, batch row
            if (col < nc && b < nb) {
                const uint8_t mk = p.mask[(((size_t)t * 2 + 0) * p.B + b0 + b) * Pd + c0 + col];
                const float v = mk ? fmaxf(acc[i], 0.f) * 2.f : 0.f;
#pragma unroll
                for (int dst = 0; dst < kPnCl; ++dst) st_cluster(P1 + (size_t)b * S2 + c0 + col, dst, v);
            }
        }
    }
    cluster_sync_();          // the whole layer-1 output is in every CTA
    // ---- layer 2 ----
    {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        pn_tile(acc, W2, P1, S2, S2 - 4, mt, nt, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int col = mt * 16 + g + (i >> 1) * 8, b = nt * 8 + 2 * tq + (i & 1);
            if (col < nc && b < nb) {
                const uint8_t mk = p.mask[(((size_t)t * 2 + 1) * p.B + b0 + b) * Pd + c0 + col];
                p.out[(size_t)(b0 + b) * p.ldo + c0 + col] = mk ? fmaxf(acc[i], 0.f) * 2.f : 0.f;
            }
        }
    }
}
bool infer_prenet_supported(int M, int Pd) { return M % 4 == 0 && Pd % 4 == 0 && (Pd + kPnCl - 1) / kPnCl <= 32; }
int k_infer_prenet(const InferPrenetParams& p, cudaStream_t st) {
    MSA_CHECK(infer_prenet_supported(p.M, p.Pd), MSA_E_UNSUPPORTED, "infer prenet: n_mel %d / prenet_dim %d", p.M, p.Pd);
    MSA_CHECK(((uintptr_t)p.w1 & 15) == 0 && ((uintptr_t)p.w2 & 15) == 0 && ((uintptr_t)p.frame & 15) == 0, MSA_E_ARG,
              "infer prenet: operands must be 16-byte aligned");
    const size_t smem = sizeof(float) * pn_layout(p.M, p.Pd).total;
    MSA_CHECK(smem <= 200 * 1024, MSA_E_UNSUPPORTED, "infer prenet: prenet_dim %d too large for the shared-memory tiles", p.Pd);
    if (smem > 48 * 1024) MSA_CUDA(cudaFuncSetAttribute(ker_infer_prenet, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kPnCl, (p.B + 31) / 32);
    cfg.blockDim = dim3(kPnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kPnCl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_prenet, p));
    MSA_LAUNCH_CHECK();
    return 0;
}

// =====================================================================================================================
// Mel / gate projection + stop logic (decoder.py:267-270, 381-395) in one 8-CTA cluster: CTA r multiplies the K slice
// [r*Ks, (r+1)*Ks) of [h_d; ctx] with the same slice of all N = n_mel + 1 weight rows (3xTF32 mma.sync), pushes the partial
// rows [j*Rc, (j+1)*Rc) to CTA j over distributed shared memory (reduce-scatter), and after one cluster barrier CTA j sums its
// rows in a fixed order, adds the bias and writes the mel frame (output + next prenet input); the CTA that owns the gate row
// runs the stop logic.  The generic rows kernel streams ALL of x through each of its 21 CTAs (17.8 us for 0.3 MFLOP per row).
constexpr int kPjThreads = 256;
constexpr int kPjCl = 8;
constexpr int kPjMT = 6;          // 16-row tiles of weight rows: N <= 96
struct PjSmem {
    int S, Ks, Rc;
    size_t x, w, part, total;
};
__host__ __device__ inline PjSmem pj_layout(int N, int K) {
    PjSmem s;
    s.Ks = (((K + kPjCl - 1) / kPjCl) + 15) & ~15;
    s.S = s.Ks + 4;
    s.Rc = (N + kPjCl - 1) / kPjCl;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
    s.x = take((size_t)32 * s.S);
    s.w = take((size_t)16 * kPjMT * s.S);
    s.part = take((size_t)kPjCl * s.Rc * 32);
    s.total = o;
    return s;
}
__global__ void __launch_bounds__(kPjThreads, 1) ker_infer_proj(InferProjParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(16) float sm[];
    __shared__ int any_s;
    const int N = p.M + 1, K = p.K;
    const PjSmem lay = pj_layout(N, K);
    const int S = lay.S, Ks = lay.Ks, Rc = lay.Rc;
    float* Xs = sm + lay.x;          // [32][S]      K slice of [h_d; ctx] (zero beyond B / K)
    float* Ws = sm + lay.w;          // [96][S]      K slice of the N weight rows (zero beyond N / K)
    float* part = sm + lay.part;     // [8][Rc][32]  partial sums of the owned rows from every CTA of the cluster
    const int r = (int)cluster_ctarank_(), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int k0 = r * Ks, kn = max(0, min(Ks, K - k0)), K4 = Ks >> 2;
    for (int i = threadIdx.x; i < 16 * kPjMT * K4; i += kPjThreads) {
        const int row = i / K4, c4 = i - row * K4;
        const bool ok = row < N && c4 * 4 < kn;
        const float* src = row < p.M ? p.wp + (size_t)row * K : p.wg;      // rows 0..M-1: mel projection, row M: gate layer
        cp_async16(Ws + (size_t)row * S + c4 * 4, ok ? src + k0 + c4 * 4 : p.wp, ok);
    }
    cp_async_commit();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool done = p.state[1] != 0;      // launch-uniform
    const int t = p.state[0];
    for (int i = threadIdx.x; i < 32 * K4; i += kPjThreads) {
        const int row = i / K4, c4 = i - row * K4;
        const bool ok = !done && row < p.B && c4 * 4 < kn;
        cp_async16(Xs + (size_t)row * S + c4 * 4, ok ? p.x + (size_t)row * p.ldx + k0 + c4 * 4 : p.x, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    cluster_sync_();
    if (done) return;
    // 8 warps: n-tile = w & 3, m-tiles (w >> 2), (w >> 2) + 2, (w >> 2) + 4
    const int nt = w & 3, g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int mt = (w >> 2) + 2 * j;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        pn_tile(acc, Ws, Xs, S, Ks, mt, nt, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = mt * 16 + g + (i >> 1) * 8, b = nt * 8 + 2 * tq + (i & 1);
            if (row < N) st_cluster(part + ((size_t)r * Rc + row % Rc) * 32 + b, row / Rc, acc[i]);
        }
    }
    cluster_sync_();          // every partial of the owned rows has arrived
    if (threadIdx.x == 0) any_s = 0;
    __syncthreads();
    const int row0 = r * Rc;
    bool gate_owner = false;
    for (int i = threadIdx.x; i < Rc * 32; i += kPjThreads) {
        const int rl = i >> 5, b = i & 31, row = row0 + rl;
        if (row >= N || b >= p.B) continue;
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < kPjCl; ++q) v += part[((size_t)q * Rc + rl) * 32 + b];
        if (row < p.M) {
            v += p.bp[row];
            p.mel_tm[((size_t)t * p.B + b) * p.M + row] = v;
            p.frame[(size_t)b * p.M + row] = v;
        } else {
            // stop gate (decoder.py:381-395): dec = sigmoid(gate) <= threshold; not_finished *= dec; mel_lengths += not_finished
            v += p.bg[0];
            const int dec = (1.f / (1.f + expf(-v))) <= p.threshold ? 1 : 0;
            const int nf = p.not_finished[b] * dec;
            p.not_finished[b] = nf;
            p.mel_lengths[b] += nf;
            if (nf) atomicOr(&any_s, 1);
        }
    }
    if (row0 <= p.M && p.M < row0 + Rc) gate_owner = true;
    __syncthreads();
    if (gate_owner && threadIdx.x == 0) {
        p.state_rw[2] = t + 1;
        if ((p.early && !any_s) || t + 1 >= p.max_steps) p.state_rw[1] = 1;
        p.state_rw[0] = t + 1;
    }
}
bool infer_proj_supported(int B, int M, int K, int ldx) {
    return B <= 32 && M + 1 <= 16 * kPjMT && K % 4 == 0 && ldx % 4 == 0 && sizeof(float) * pj_layout(M + 1, K).total <= 200 * 1024;
}
int k_infer_proj(const InferProjParams& p, cudaStream_t st) {
    MSA_CHECK(infer_proj_supported(p.B, p.M, p.K, p.ldx), MSA_E_UNSUPPORTED, "infer proj: B %d, n_mel %d, K %d", p.B, p.M, p.K);
    MSA_CHECK(((uintptr_t)p.wp & 15) == 0 && ((uintptr_t)p.wg & 15) == 0 && ((uintptr_t)p.x & 15) == 0, MSA_E_ARG,
              "infer proj: operands must be 16-byte aligned");
    const size_t smem = sizeof(float) * pj_layout(p.M + 1, p.K).total;
    if (smem > 48 * 1024) MSA_CUDA(cudaFuncSetAttribute(ker_infer_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kPjCl);
    cfg.blockDim = dim3(kPjThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kPjCl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_proj, p));
    MSA_LAUNCH_CHECK();
    return 0;
}

__global__ void ker_init_fwd_attn(float* alpha, float* u, int* win, float* gmax, int B, int L) {
    // init_forward_attn / init_win_idx (forward_attn.py:85-96): alpha = [1, 1e-7, ...], u = 0.5, win_idx = -1
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * L; i += gridDim.x * blockDim.x) alpha[i] = (i % L) == 0 ? 1.f : 1e-7f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) u[i] = 0.5f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        win[0] = -1;
        win[1] = -1;
        *gmax = -INFINITY;
    }
}
int k_init_fwd_attn(float* alpha, float* u, int* win, float* gmax, int B, int L, cudaStream_t st) {
    ker_init_fwd_attn<<<cdiv((int64_t)B * L, 256), 256, 0, st>>>(alpha, u, win, gmax, B, L);
    MSA_LAUNCH_CHECK();
    return 0;
}

size_t infer_attention_smem(int L, int Ha, int A, int F, int Kl, int E, bool mem_res) {
    return sizeof(float) * ia_layout(L, Ha, A, F, Kl, E, mem_res).total;
}

int k_infer_attention(const InferAttnParams& p, cudaStream_t st) {
    MSA_CHECK(p.A <= 32 * kIaDJ, MSA_E_UNSUPPORTED, "infer attention: attention_dim %d > %d", p.A, 32 * kIaDJ);
    MSA_CHECK(p.Ha % 4 == 0 && p.ldh % 4 == 0 && ((uintptr_t)p.h & 15) == 0 && ((uintptr_t)p.wq & 15) == 0, MSA_E_UNSUPPORTED,
              "infer attention: attention_rnn_dim %d / row stride %d must be multiples of 4", p.Ha, p.ldh);
    InferAttnParams pm = p;
    pm.mem_res = ia_mem_tile_ok(p.E) && infer_attention_smem(p.L, p.Ha, p.A, p.F, p.Kl, p.E, true) <= 200 * 1024;
    const size_t smem = infer_attention_smem(p.L, p.Ha, p.A, p.F, p.Kl, p.E, pm.mem_res != 0);
    MSA_CHECK(smem <= 200 * 1024, MSA_E_UNSUPPORTED, "infer attention: text length %d too long for the shared-memory tile", p.L);
    MSA_CHECK(((uintptr_t)p.wloc_t & 15) == 0 && ((uintptr_t)p.wld4 & 15) == 0 && ((uintptr_t)p.memory & 15) == 0, MSA_E_ARG,
              "infer attention: prepared weights / memory must be 16-byte aligned");
    if (smem > 48 * 1024) MSA_CUDA(cudaFuncSetAttribute(ker_infer_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static bool carve_set = false;
    if (!carve_set) {
        MSA_CUDA(cudaFuncSetAttribute(ker_infer_attn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        carve_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.B * kIaCl);
    cfg.blockDim = dim3(kIaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kIaCl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    MSA_CUDA(cudaLaunchKernelEx(&cfg, ker_infer_attn, pm));
    MSA_LAUNCH_CHECK();
    return 0;
}

}  // namespace msa
