"""The hand-written tcgen05 / TMA GEMM (msa_gemm_nt) against a float64 matmul.

Tolerances: mode 0 (3xTF32 split): relative Frobenius error < 4e-6 + 1e-8*K (the tensor core accumulates in fp32 with
truncation, so the error grows linearly with K: measured 2.3e-6 at K=256, 1.9e-5 at K=2560 -- two orders of magnitude below the
2e-4 parity tolerance of the pass); mode 1 (single TF32 product) < 2e-3."""
import ctypes as C

import pytest
import torch

from msa_tts_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(M, N, K, mode, alpha=1.0, beta=0.0, lda=None, ldb=None, ldc=None, seed=0):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    lda, ldb, ldc = lda or K, ldb or K, ldc or N
    A = torch.randn(M, lda, device="cuda", generator=g)
    B = torch.randn(N, ldb, device="cuda", generator=g)
    Cm = torch.randn(M, ldc, device="cuda", generator=g)
    C0 = Cm.clone()
    scratch = torch.empty(int(lib.msa_gemm_nt_scratch_floats(M, N, K)) + 4, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.msa_gemm_nt(M, N, K, C.c_float(alpha), P(A), lda, P(B), ldb, C.c_float(beta), P(Cm), ldc, mode, P(scratch),
                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "msa_gemm_nt")
    torch.cuda.synchronize()
    ref = alpha * (A[:, :K].double() @ B[:, :K].double().t()) + beta * C0[:, :N].double()
    err = float((Cm[:, :N].double() - ref).norm() / ref.norm())
    assert torch.equal(Cm[:, N:], C0[:, N:]), "columns beyond N must not be touched"
    return err


@pytest.mark.parametrize("shape", [(128, 128, 32), (128, 128, 256), (800, 4096, 1024), (804, 256, 80), (256, 512, 2560),
                                   (37, 130, 100), (1, 8, 4), (300, 81, 1792)])
def test_gemm_3xtf32_is_fp32_accurate(shape):
    assert _run(*shape, mode=0) < 4e-6 + 1e-8 * shape[2]


@pytest.mark.parametrize("shape", [(128, 128, 64), (800, 4096, 1024), (37, 130, 100)])
def test_gemm_single_tf32(shape):
    assert _run(*shape, mode=1) < 2e-3


def test_gemm_alpha_beta_and_leading_dimensions():
    assert _run(200, 300, 768, mode=0, alpha=0.5, beta=1.0, lda=1024, ldb=1792, ldc=304) < 4e-6 + 1e-8 * 768
    assert _run(200, 300, 768, mode=1, alpha=2.0, beta=-0.5, lda=1024, ldb=1792, ldc=304) < 2e-3


def test_pass_parity_with_tcgen05_forward_gemms(monkeypatch):
    """The whole teacher-forced pass with its forward projections on the tcgen05 3xTF32 kernel (MSA_GEMM_TC=1, GEMM policy 1)
    against the oracle: same tolerances as the cuBLAS route (outputs 3e-4, gradients 2e-3 of the global norm)."""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from msa_tts_b200.engine import Engine
    from helpers import cuda_pass, oracle_pass, rel
    monkeypatch.setenv("MSA_GEMM_TC", "1")
    cfg = pkg.small_params()
    crit = dict(reduction="none", pos_weight=10.0)
    B, T, L = 4, 40, 24
    eng = Engine(cfg, reduction="none", pos_weight=10.0, gemm_tf32=1)
    P = synth.init_params(cfg, 9)
    batch = synth.make_batch(cfg, B, T, L, 109)
    masks = synth.make_masks(cfg, B, T, L, 209)
    o_out, o_loss, o_g, _, _ = oracle_pass(cfg, P, batch, masks, crit)
    c_out, c_loss, c_g, _ = cuda_pass(eng, cfg, P, batch, masks)
    for a, b in zip(c_out, o_out):
        assert rel(a, b) < 3e-4
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in o_g.values())))
    assert max(float((c_g[n].double().cpu() - o_g[n].double()).norm()) / gn for n in o_g) < 2e-3


def _run_general(ta, tb, M, N, K, mode, alpha=1.0, beta=0.0, pad=0, seed=0):
    """C = alpha op(A) op(B) + beta C through msa_gemm; operands stored [K][M] / [K][N] when their contraction index is the row."""
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    a_shape = (K, M + pad) if ta else (M, K + pad)
    b_shape = (N, K + pad) if tb else (K, N + pad)
    A = torch.randn(*a_shape, device="cuda", generator=g)
    B = torch.randn(*b_shape, device="cuda", generator=g)
    Cm = torch.randn(M, N + pad, device="cuda", generator=g)
    C0 = Cm.clone()
    scratch = torch.empty(int(lib.msa_gemm_nt_scratch_floats(M, N, K)) + 4, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.msa_gemm(int(ta), int(tb), M, N, K, C.c_float(alpha), P(A), a_shape[1], P(B), b_shape[1], C.c_float(beta), P(Cm), N + pad,
                      mode, P(scratch), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "msa_gemm")
    torch.cuda.synchronize()
    Ad = (A[:, :M].double().t() if ta else A[:, :K].double())
    Bd = (B[:, :K].double().t() if tb else B[:, :N].double())
    ref = alpha * (Ad @ Bd) + beta * C0[:, :N].double()
    return float((Cm[:, :N].double() - ref).norm() / ref.norm())


@pytest.mark.parametrize("tt", [(0, 0), (1, 0), (1, 1), (0, 1)])
@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 384, 160), (4096, 1024, 800), (800, 1792, 4096), (100, 36, 52), (512, 2560, 800)])
def test_gemm_transposed_operands_as_mn_major_tiles(tt, shape):
    """The backward contractions dX = dY . W (0, 0) and dW = dY^T . X (1, 0) (and (1, 1) for completeness) on the tcgen05 kernel:
    operands whose contraction index is the row index are MN-major shared-memory tiles, nothing is transposed in memory."""
    M, N, K = shape
    if (tt[0] and M % 4) or (not tt[1] and N % 4):
        pytest.skip("leading dimension of an MN-major operand must be a multiple of 4 floats")
    assert _run_general(tt[0], tt[1], M, N, K, mode=0) < 4e-6 + 1e-8 * K
    assert _run_general(tt[0], tt[1], M, N, K, mode=1, alpha=0.5, beta=1.0, pad=4) < 2e-3
