"""The oracle restatement against the committed golden fixtures (reference outputs).

The fixtures were written by oracle/gen_golden.py from the unmodified reference
(msa_tts/models/tacotron2nv.py, tacotron2nv_loss.py); here the restatement is
re-run on the same seeded inputs and compared.  Tolerance: fp32 CPU, 2e-5
relative (outputs: per tensor; gradients: relative to the global gradient norm,
SURVEY.md Q17)."""
import copy
import os

import numpy as np
import pytest
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from oracle import model as OM
from oracle.gen_cases import CASES, INFER_CASES, infer_stats, speaker_input

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", [c for c in CASES if c != "default_train_b4_t200"])
def test_train_small(name):
    cfg, seed, (B, T, L), crit = CASES[name]()
    z = np.load(os.path.join(GOLD, name + ".npz"))
    P = synth.init_params(cfg, seed)
    batch = synth.make_batch(cfg, B, T, L, seed + 100)
    masks = synth.make_masks(cfg, B, T, L, seed + 200)
    _, inp, inp_len, mels, mel_len, _, _, stop = batch
    spk = speaker_input(cfg, batch)
    Pl = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    stats = OM.fresh_bn_stats(Pl, cfg)
    out = OM.forward(Pl, cfg, inp, inp_len, mels, mel_len, spk, masks, stats, True)
    loss = OM.loss_fn(out, (mels, stop), mel_len, **crit)
    names = list(P.keys())
    g = torch.autograd.grad(loss, [Pl[n] for n in names], allow_unused=True)
    for key, o in zip(("mel", "mel_post", "gate", "align"), out):
        assert rel(o.detach(), z[key]) < TOL, key
    assert abs(float(loss.detach()) - float(z["loss"])) < TOL * abs(float(z["loss"]))
    gn = np.sqrt(sum(float((z["grad/" + n].astype(np.float64) ** 2).sum()) for n in names))
    for n, gi in zip(names, g):
        gi = torch.zeros_like(Pl[n]) if gi is None else gi
        assert float((gi.double() - torch.as_tensor(z["grad/" + n]).double()).norm()) / gn < TOL, n
    for k in stats:
        if "running" in k:
            assert rel(stats[k], z["stat/" + k]) < TOL, k
        else:
            assert int(stats[k]) == int(z["stat/" + k])


def test_train_default_compact():
    name = "default_train_b4_t200"
    cfg, seed, (B, T, L), crit = CASES[name]()
    z = np.load(os.path.join(GOLD, name + ".npz"))
    P = synth.init_params(cfg, seed)
    np.testing.assert_allclose(np.array([float(v.double().sum()) for v in P.values()]), z["param_checksum"], rtol=1e-9,
                               err_msg="synthetic init differs from the one the fixture was generated with")
    batch = synth.make_batch(cfg, B, T, L, seed + 100)
    masks = synth.make_masks(cfg, B, T, L, seed + 200)
    _, inp, inp_len, mels, mel_len, _, spk, stop = batch
    Pl = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    out = OM.forward(Pl, cfg, inp, inp_len, mels, mel_len, spk, masks, OM.fresh_bn_stats(Pl, cfg), True)
    loss = OM.loss_fn(out, (mels, stop), mel_len, **crit)
    names = list(P.keys())
    g = torch.autograd.grad(loss, [Pl[n] for n in names], allow_unused=True)
    for key, o in zip(("mel", "mel_post", "gate"), out):
        assert rel(o.detach(), z[key]) < TOL, key
    assert rel(out[3].detach()[:, ::8], z["align_sample"]) < TOL
    assert abs(float(loss.detach()) - float(z["loss"])) < TOL * abs(float(z["loss"]))
    gn = float(np.sqrt((z["grad_norms"] ** 2).sum()))
    for i, (n, gi) in enumerate(zip(names, g)):
        assert abs(float(gi.double().norm()) - z["grad_norms"][i]) / gn < TOL, n
        flat = gi.flatten()
        samp = flat[:: max(1, flat.numel() // 64)][:64]
        assert float((samp.double() - torch.as_tensor(z["gsample/" + n]).double()).abs().max()) < 1e-4 * gn, n


@pytest.mark.parametrize("name", list(INFER_CASES))
def test_infer_small(name):
    cfg, seed, (B, L), steps = INFER_CASES[name]()
    z = np.load(os.path.join(GOLD, name + ".npz"))
    P = synth.init_params(cfg, seed)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, seed + 100)
    stats = infer_stats(P, cfg, seed)
    pm = synth.make_infer_masks(cfg, B, steps, seed + 300)
    post, lens, align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert post.shape[2] == int(z["steps"])
    assert np.array_equal(lens.numpy(), z["mel_lengths"])          # bit-exact stop bookkeeping
    assert lens.dtype == torch.int32
    assert rel(post, z["mel_post"]) < TOL and rel(align, z["align"]) < TOL


def test_second_order_maml_oracle_matches_a_float64_hessian_vector_difference():
    """oracle.meta.maml2_task (maml.py:70-71, track_higher_grads=True: the test loss differentiated through the inner SGD step) against
    an independent evaluation of the same quantity in float64: g_test - lr * H_train . g_test with the Hessian-vector product taken as a
    central difference of first-order gradients at a relative step of 1e-7.  (The same difference in float32 is off by 20-100 % at
    every step size -- profiles/r02_second_order_fd.txt -- which is why the CUDA path offers no finite-difference second-order mode.)"""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from oracle import meta as OMeta
    from oracle import model as OM
    cfg = pkg.small_params()
    crit = dict(reduction="none", pos_weight=10.0)
    B, T, L, lr = 3, 10, 8, 0.05
    names = OM.param_names(cfg)
    to64 = lambda d: {k: (v.double() if v.dtype.is_floating_point else v) for k, v in d.items()}
    P = to64(synth.init_params(cfg, 5))
    task = {k: tuple(x.double() if hasattr(x, "dtype") and x.dtype == torch.float32 else x for x in v)
            for k, v in synth.make_task(cfg, B, T, L, 3).items()}
    masks = [synth.make_masks(cfg, B, T, L, 900 + i) for i in range(2)]
    _, g2 = OMeta.maml2_task(P, cfg, task, masks, crit, names, 1, lr)
    stats = OM.fresh_bn_stats(P, cfg)
    _, g, _ = OMeta.loss_and_grads(P, cfg, task["train"], masks[0], stats, crit, names)
    th1 = OMeta.sgd_step(P, g, names, lr)
    _, v, _ = OMeta.loss_and_grads(th1, cfg, task["test"], masks[1], stats, crit, names)
    norm = lambda d: float(torch.sqrt(sum((d[n] ** 2).sum() for n in names)))
    eps = 1e-7 * norm(P) / norm(v)
    gp = OMeta.loss_and_grads({n: P[n] + eps * v[n] for n in names}, cfg, task["train"], masks[0], OM.fresh_bn_stats(P, cfg), crit, names)[1]
    gm = OMeta.loss_and_grads({n: P[n] - eps * v[n] for n in names}, cfg, task["train"], masks[0], OM.fresh_bn_stats(P, cfg), crit, names)[1]
    fd = {n: v[n] - lr * (gp[n] - gm[n]) / (2 * eps) for n in names}
    err = norm({n: fd[n] - g2[n] for n in names}) / norm(g2)
    first_order_gap = norm({n: v[n] - g2[n] for n in names}) / norm(g2)
    assert err < 1e-6, err
    assert first_order_gap > 0.1          # the second-order term is not a rounding-level correction at this shape
