"""The implicit-GEMM convolutions (msa_conv1d_fwd / _dx / _dw, gemm_tc.cu) against torch.nn.functional.conv1d in float64 and its
autograd: the contraction of every Encoder / Postnet ConvNorm (modules_tacotron2nv/encoder.py:36-37, decoder.py:63-72) and of its
backward, plus the BatchNorm slab statistics the forward epilogue emits."""
import ctypes as C

import pytest
import torch

from msa_tts_b200 import _lib

pytestmark = pytest.mark.gpu

P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _case(B, T, Ci, Co, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, T, Ci, device="cuda", generator=g)
    w = torch.randn(Co, Ci, K, device="cuda", generator=g) / (Ci * K) ** 0.5
    b = torch.randn(Co, device="cuda", generator=g)
    dy = torch.randn(B, T, Co, device="cuda", generator=g)
    return x, w, b, dy


def _ref(x, w, b, dy):
    xd = x.double().transpose(1, 2).requires_grad_(True)
    wd = w.double().requires_grad_(True)
    y = torch.nn.functional.conv1d(xd, wd, b.double(), padding=(w.shape[2] - 1) // 2)
    gx, gw = torch.autograd.grad(y, (xd, wd), dy.double().transpose(1, 2))
    return y.transpose(1, 2).contiguous(), gx.transpose(1, 2).contiguous(), gw


SHAPES = [(4, 200, 512, 512, 5), (4, 200, 80, 512, 5), (4, 200, 512, 80, 5), (4, 64, 512, 512, 5), (1, 7, 16, 8, 3), (3, 130, 36, 52, 5),
          (2, 257, 64, 128, 1), (5, 33, 24, 40, 7), (32, 96, 128, 256, 5)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("split", [True, False])
def test_conv_forward_input_gradient_weight_gradient(shape, split):
    B, T, Ci, Co, K = shape
    lib = _lib.load()
    x, w, b, dy = _case(B, T, Ci, Co, K, 11)
    y_ref, gx_ref, gw_ref = [t.detach() for t in _ref(x, w, b, dy)]
    wp = torch.empty(K, Co, Ci, device="cuda")
    _lib.check(lib.msa_conv1d_repack(P(w), P(wp), Co, Ci, K, _stream()), "repack")
    torch.cuda.synchronize()
    assert torch.equal(wp, w.permute(2, 0, 1).contiguous())
    scratch = torch.empty(int(lib.msa_conv1d_scratch_floats(B, T, Ci, Co, K)) + 4, device="cuda") if split else None
    nslab = int(lib.msa_conv1d_stat_slabs(B, T, Ci, Co, K, int(split)))
    for mode, tol in ((0, 4e-6 + 1e-8 * K * max(Ci, Co)), (1, 2e-3), (2, 8e-4), (3, 1.2e-3)):      # 3xTF32: fp32-accurate; single TF32 truncated / rounded to nearest / B rounded only (the pass feeds it a TF32-exact A)
        y = torch.full((B, T, Co), float("nan"), device="cuda")
        stats = torch.full((nslab, 3, Co), float("nan"), device="cuda")
        _lib.check(lib.msa_conv1d_fwd(P(x), B, T, Ci, P(wp), Co, K, P(b), P(y), mode, P(scratch), P(stats), _stream()), "fwd")
        dx = torch.full((B, T, Ci), float("nan"), device="cuda")
        _lib.check(lib.msa_conv1d_dx(P(dy), B, T, Co, P(wp), Ci, K, P(dx), mode, P(scratch), _stream()), "dx")
        dw = torch.full((Co, Ci, K), float("nan"), device="cuda")
        _lib.check(lib.msa_conv1d_dw(P(dy), P(x), B, T, Co, Ci, K, C.c_float(1.0), 0, P(dw), mode, _stream()), "dw")
        dw2 = dw.clone()
        _lib.check(lib.msa_conv1d_dw(P(dy), P(x), B, T, Co, Ci, K, C.c_float(0.5), 1, P(dw2), mode, _stream()), "dw acc")
        torch.cuda.synchronize()
        e_y = float((y.double() - y_ref).norm() / y_ref.norm())
        e_x = float((dx.double() - gx_ref).norm() / gx_ref.norm())
        e_w = float((dw.double() - gw_ref).norm() / gw_ref.norm())
        e_w2 = float((dw2.double() - 1.5 * gw_ref).norm() / gw_ref.norm())
        print(f"{shape} split={split} mode={mode}: y {e_y:.2e} dx {e_x:.2e} dw {e_w:.2e} dw(acc) {e_w2:.2e}")
        tol_w = tol if mode else 4e-6 + 1e-8 * B * T          # the weight gradient contracts over the B * T rows
        assert e_y < tol and e_x < tol and e_w < tol_w and e_w2 < 1.5 * tol_w
        # slab statistics -> mean / biased variance of every channel (Chan's merge in slab order, as the normalisation kernel does)
        n = torch.zeros(Co, dtype=torch.float64, device="cuda")
        m = torch.zeros_like(n)
        m2 = torch.zeros_like(n)
        for sl in range(nslab):
            cnt, mu, q = stats[sl, 0].double(), stats[sl, 1].double(), stats[sl, 2].double()
            nn = n + cnt
            delta = mu - m
            safe = torch.where(nn > 0, nn, torch.ones_like(nn))
            m = m + delta * cnt / safe
            m2 = m2 + q + delta * delta * n * cnt / safe
            n = nn
        yf = y.double().reshape(-1, Co)
        assert torch.all(n == B * T)
        assert float((m - yf.mean(0)).abs().max()) < 1e-5 * (1 + float(yf.abs().max()))
        assert float((m2 / (B * T) - yf.var(0, unbiased=False)).abs().max()) < 1e-4 * float(yf.var(0, unbiased=False).max())
