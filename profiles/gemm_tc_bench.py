"""Time msa_gemm_nt (hand-written tcgen05 / TMA GEMM, 3xTF32 and single TF32) against cuBLAS (fp32 SIMT, TF32 tensor-op) on the
x.W^T contractions of the pass at the bench dimensions.    python profiles/gemm_tc_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from msa_tts_b200 import _lib

lib = _lib.load()
SHAPES = [("dec-RNN gates W_hh part", 800, 4096, 1024), ("dec-RNN gates ctx part", 800, 4096, 768), ("attn-LSTM prenet part", 804, 4096, 256),
          ("postnet conv 512->512 k5", 800, 512, 2560), ("postnet conv 80->512", 800, 512, 400), ("postnet conv 512->80", 800, 80, 2560),
          ("encoder conv", 256, 512, 2560), ("MW = Wc.memory", 256, 4096, 768), ("mel/gate projection", 800, 81, 1792), ("prenet 2", 804, 256, 256)]


def timed(fn, reps=20):
    """GPU time per call: `reps` calls captured in one CUDA graph (host launch overhead would dominate the small shapes)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


P = lambda t: C.c_void_p(t.data_ptr())
# ---- the same products with the leading dimensions they have inside the pass, L2-hot and with the L2 flushed before every call ----
flush_buf = torch.empty(64 * 1024 * 1024, device="cuda")           # 256 MB > 126 MB of L2
PASS_LD = [("dec gates h_a part, ldb 1792", 800, 4096, 1024, 1024, 1792, 0.0), ("dec gates ctx part, beta 1", 800, 4096, 768, 768, 1792, 1.0),
           ("attn-LSTM prenet part, ldb 1024", 804, 4096, 256, 256, 1024, 0.0)]
print(f"{'contraction (pass leading dims)':36s} | tc 3xTF32 hot / cold | cuBLAS fp32 hot / cold   (us, cold = after an L2 flush, flush time subtracted)")
for name, M, N, K, lda, ldb, beta in PASS_LD:
    A, B = torch.randn(M, lda, device="cuda"), torch.randn(N, ldb, device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    scratch = torch.empty(int(lib.msa_gemm_nt_scratch_floats(M, N, K)) + 4, device="cuda")

    def tc0():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.msa_gemm_nt(M, N, K, C.c_float(1.0), P(A), lda, P(B), ldb, C.c_float(beta), P(Cm), N, 0, P(scratch), st), "msa_gemm_nt")

    def cb():
        if beta:
            torch.addmm(Cm, A[:, :K], B[:, :K].t(), out=Cm)
        else:
            torch.mm(A[:, :K], B[:, :K].t(), out=Cm)
    torch.backends.cuda.matmul.allow_tf32 = False
    tf = timed(lambda: flush_buf.zero_())
    res = []
    for fn in (tc0, cb):
        hot = timed(fn)
        cold = timed(lambda: (flush_buf.zero_(), fn())) - tf
        res += [hot, cold]
    print(f"{name:36s} | {res[0]:9.1f} / {res[1]:6.1f}  | {res[2]:9.1f} / {res[3]:6.1f}")
print()
print(f"{'contraction':28s} {'M':>5s} {'N':>5s} {'K':>5s} | tc 3xTF32  tc TF32 | cuBLAS fp32  cuBLAS TF32   (us)   err(3x)")
for name, M, N, K in SHAPES:
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    Cm = torch.empty(M, N, device="cuda")
    scratch = torch.empty(int(lib.msa_gemm_nt_scratch_floats(M, N, K)) + 4, device="cuda")

    def tc(mode):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.msa_gemm_nt(M, N, K, C.c_float(1.0), P(A), K, P(B), K, C.c_float(0.0), P(Cm), N, mode, P(scratch), st), "msa_gemm_nt")
    t0, t1 = timed(lambda: tc(0)), timed(lambda: tc(1))
    tc(0)
    ref = A.double() @ B.double().t()
    err = float((Cm.double() - ref).norm() / ref.norm())
    torch.backends.cuda.matmul.allow_tf32 = False
    t2 = timed(lambda: torch.mm(A, B.t(), out=Cm))
    torch.backends.cuda.matmul.allow_tf32 = True
    t3 = timed(lambda: torch.mm(A, B.t(), out=Cm))
    print(f"{name:28s} {M:5d} {N:5d} {K:5d} | {t0:9.1f} {t1:8.1f} | {t2:11.1f} {t3:12.1f}          {err:.1e}")
