"""Flat-buffer kernels (msa_flat_*, msa_ewc_*) against the reference formulas
(utils/grad_utils.py:8-31, reptile.py:73-77, torch.optim.SGD/Adam, continual_ewc.py:59-89) -- bit-level
agreement is not expected for FMA-contracted fp32, tolerance 1e-6 relative."""
import pytest
import torch

import msa_tts_b200 as pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from msa_tts_b200.engine import Engine
    return Engine(pkg.small_params())


def _r(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g).cuda()


N = 4 * 100003 * 4  # not a multiple of the grid, > 1M elements


def close(a, b, tol=1e-6):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)) < tol


def test_sgd_plain_and_momentum(eng):
    p, g = _r(N, 1), _r(N, 2)
    out = eng.sgd_step(p.clone(), g, lr=0.01)
    assert close(out, p - 0.01 * g)
    # momentum + weight decay + nesterov, two steps, against torch.optim.SGD
    pt = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([pt], lr=0.05, momentum=0.9, weight_decay=0.01, nesterov=True)
    q, buf = p.clone(), torch.empty_like(p)
    for step in range(2):
        gg = _r(N, 10 + step)
        pt.grad = gg.clone()
        opt.step()
        eng.sgd_step(q, gg, lr=0.05, momentum=0.9, weight_decay=0.01, nesterov=True, buf=buf, first_step=(step == 0))
    assert close(q, pt.data)


def test_axpy_mix_grad(eng):
    gs = [_r(N, 20 + i) for i in range(4)]
    acc = torch.empty(N, device="cuda")
    for i, g in enumerate(gs):
        eng.axpy(acc, g, 0.25, init=(i == 0))
    ref = torch.stack([0.25 * g for g in gs]).sum(dim=0)          # mix_grad, grad_utils.py:23-31
    assert close(acc, ref)


def test_reptile_delta(eng):
    p0, pT = _r(N, 30), _r(N, 31)
    acc = torch.empty(N, device="cuda")
    eng.reptile_delta(acc, pT, p0, 1.0, init=True)
    assert torch.equal(acc, -(pT - p0))                            # reptile.py:76, exact
    eng.reptile_delta(acc, pT, p0, 0.5, init=False)
    assert close(acc, -1.5 * (pT - p0))


def test_sumsq_and_clip_sgd(eng):
    p, g = _r(N, 40), _r(N, 41)
    ss = eng.sumsq(g)
    assert abs(float(ss) - float((g.double() ** 2).sum())) / float((g.double() ** 2).sum()) < 1e-6
    ss2 = eng.sumsq(g)
    assert torch.equal(ss, ss2)                                    # deterministic reduction
    pt = torch.nn.Parameter(p.clone())
    pt.grad = g.clone()
    torch.nn.utils.clip_grad_norm_([pt], 1.0)
    torch.optim.SGD([pt], lr=0.1).step()
    q = p.clone()
    eng.clip_sgd(q, g, ss, lr=0.1, max_norm=1.0)
    assert close(q, pt.data)


def test_clip_adam(eng):
    p = _r(N, 50)
    pt = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([pt], lr=1e-3)
    q, m, v = p.clone(), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = _r(N, 50 + step)
        pt.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([pt], 5.0)
        opt.step()
        eng.clip_adam(q, g, m, v, eng.sumsq(g), lr=1e-3, step=step, max_norm=5.0)
    assert close(q, pt.data, 2e-6)


def test_functional_adam_step(eng):
    """Inner-loop Adam (msa_flat_adam_step): torch.optim.Adam's rule with weight decay, out of place (p stays intact)."""
    p = _r(N, 70)
    pt = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([pt], lr=2e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.01)
    cur, m, v = p.clone(), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = _r(N, 70 + step)
        pt.grad = g.clone()
        opt.step()
        keep, new = cur.clone(), torch.empty_like(cur)
        eng.adam_step(cur, g, new, m, v, lr=2e-3, step=step, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.01)
        assert torch.equal(cur, keep), "functional step: the input fast weights are not modified"
        cur = new
    assert close(cur, pt.data, 2e-6)


def test_device_dropout_masks(eng):
    """msa_masks_generate: one launch for all sections; keep rate 1 - p per section, values a function of (seed, byte index) only,
    bytes between sections untouched."""
    B, T, L = 3, 20, 9
    secs = eng.mask_sections(B, T, L)
    a = torch.full((eng.mask_bytes(B, T, L),), 7, dtype=torch.uint8, device="cuda")
    eng.generate_masks(B, T, L, 1234, out=a)
    b = eng.generate_masks(B, T, L, 1234)
    c = eng.generate_masks(B, T, L, 1235)
    covered = torch.zeros_like(a, dtype=torch.bool)
    for name, off, ne, p in secs:
        m = a[off:off + ne]
        assert int(m.max()) <= 1
        keep = float(m.float().mean())
        sigma = (p * (1 - p) / ne) ** 0.5
        assert abs(keep - (1 - p)) < 5 * sigma + 1e-9, (name, keep, p)
        assert torch.equal(m, b[off:off + ne]) and not torch.equal(m, c[off:off + ne])
        covered[off:off + ne] = True
    assert bool((a[~covered] == 7).all()), "alignment gaps are not written"


def test_ewc(eng):
    p, mu, g = _r(N, 60), _r(N, 61), _r(N, 62)
    f = torch.empty(N, device="cuda")
    gs = [_r(N, 70 + i) for i in range(3)]
    for i, x in enumerate(gs):
        eng.ewc_fisher_accum(f, x, 1.0 / 3, init=(i == 0))
    fref = sum(x ** 2 / 3 for x in gs)                              # continual_ewc.py:78-79
    assert close(f, fref)
    pen = eng.ewc_penalty(p, mu, f)
    pref = float((fref.double() * (p.double() - mu.double()) ** 2).sum())   # continual_ewc.py:84-89
    assert abs(float(pen) - pref) / pref < 1e-5
    q = p.clone()
    pen2 = eng.ewc_sgd_step(q, g, mu, f, lr=0.01, lam=3.0)
    assert abs(float(pen2) - pref) / pref < 1e-5
    assert close(q, p - 0.01 * (g + 2 * 3.0 * fref * (p - mu)))


def test_argument_errors(eng):
    with pytest.raises(RuntimeError):
        eng.axpy(torch.empty(6, device="cuda"), torch.empty(6, device="cuda"), 1.0, True)   # n % 4 != 0
