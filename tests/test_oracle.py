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
    assert abs(float(loss) - float(z["loss"])) < TOL * abs(float(z["loss"]))
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
    assert abs(float(loss) - float(z["loss"])) < TOL * abs(float(z["loss"]))
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
