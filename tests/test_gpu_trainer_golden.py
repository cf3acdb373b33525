"""The CUDA trainer-level path against the reference's own results (tests/golden/trainer.npz, see test_trainer_golden.py):
``mix_grad`` / ``apply_grad`` on flat buffers, the fused clip + SGD / Adam outer step, the device-side mcd metric, EWC Fisher /
penalty / BatchNorm buffers -- and the abort word of the persistent kernels reaching the trainer."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from oracle.gen_golden_trainer import CLIP, CRIT, LR_ADAM, LR_SGD, N_TASKS, trainer_inputs

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "trainer.npz"))


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _model(cfg, P):
    m = pkg.Tacotron2NV(cfg)
    m.flat.copy_(m.engine.flat_from_dict(P))
    return m


def test_mix_grad_apply_grad_and_outer_steps_match_the_reference():
    """maml.py:94-105 with this package's drop-ins: mix_grad(grad_list, weight) -> apply_grad(model, grads) -> stock
    clip_grad_norm_ + torch.optim step on the parameter views, AND the fused flat path the trainers use."""
    from msa_tts_b200.grad_utils import apply_grad, mix_grad
    cfg, P, names, rounds, _, _, _, _ = trainer_inputs()
    model = _model(cfg, P)
    eng = model.engine
    grad_lists = [model.layout_views(eng.flat_from_dict(g)) for g in rounds[0]]
    weight = torch.ones(N_TASKS) / N_TASKS
    model.zero_grad()
    mixed = mix_grad(grad_lists, weight)                                   # reference signature
    assert max(_rel(m, Z["mixed/" + n]) for n, m in zip(names, mixed)) < 2e-6
    norm = apply_grad(model, mixed)
    assert abs(norm - float(Z["grad_norm"])) < 1e-5 * float(Z["grad_norm"])
    opt = torch.optim.SGD(model.parameters(), lr=LR_SGD)
    torch.nn.utils.clip_grad_norm_(model.parameters(), CLIP)
    opt.step()
    got = eng.dict_from_flat(model.flat)
    upd = {n: torch.as_tensor(Z["sgd1/" + n]) - P[n] for n in names}
    scale = float(torch.sqrt(sum((u.double() ** 2).sum() for u in upd.values())))
    assert max(float(((got[n].cpu() - P[n]).double() - upd[n].double()).norm()) for n in names) / scale < 1e-4
    # fused flat path (MetaTrainer._outer_update): axpy accumulation, sumsq, clip + step in one kernel
    for opt_name, lr, n_rounds in (("sgd", LR_SGD, 1), ("adam", LR_ADAM, 2)):
        theta = eng.flat_from_dict(P)
        m, v, acc, ss = eng.new_flat(), eng.new_flat(), eng.new_flat(None), torch.zeros(1, device="cuda")
        for r in range(n_rounds):
            for i, g in enumerate(rounds[r]):
                eng.axpy(acc, eng.flat_from_dict(g), 1.0 / N_TASKS, init=(i == 0))
            eng.sumsq(acc, ss)
            if opt_name == "sgd":
                eng.clip_sgd(theta, acc, ss, lr=lr, max_norm=CLIP)
            else:
                eng.clip_adam(theta, acc, m, v, ss, lr=lr, step=r + 1, max_norm=CLIP)
            got = eng.dict_from_flat(theta)
            upd = {n: torch.as_tensor(Z[f"{opt_name}{r + 1}/" + n]) - P[n] for n in names}
            scale = float(torch.sqrt(sum((u.double() ** 2).sum() for u in upd.values())))
            err = max(float(((got[n].cpu() - P[n]).double() - upd[n].double()).norm()) for n in names) / scale
            assert err < 1e-4, (opt_name, r, err)


def test_ewc_matches_the_reference_class_including_the_bn_buffers():
    """EWC(model, buffer, criterion, device): Fisher, means, penalty(model') and the BatchNorm running statistics /
    num_batches_tracked the Fisher passes leave in the model (continual_ewc.py:28-89; the reference forwards the model itself)."""
    cfg, P, names, _, _, buf, buf_masks, P_moved = trainer_inputs()
    model = _model(cfg, P)
    ewc = pkg.EWC(model, buf, None, None, masks=buf_masks)
    F_ = model.engine.dict_from_flat(ewc.fisher)
    scale = np.sqrt(sum(float((Z["fisher/" + n].astype(np.float64) ** 2).sum()) for n in names))
    assert max(float((F_[n].double().cpu() - torch.as_tensor(Z["fisher/" + n]).double()).norm()) for n in names) / scale < 2e-4
    sd = model.state_dict()
    for k in Z.files:
        if k.startswith("ewc_stat/"):
            name = k[len("ewc_stat/"):]
            if name.endswith("num_batches_tracked"):
                assert int(sd[name]) == int(Z[k]), name
            else:
                assert _rel(sd[name], Z[k]) < 2e-4, name
    model.flat.copy_(model.engine.flat_from_dict(P_moved))
    pen = float(ewc.penalty(model))
    assert abs(pen - float(Z["penalty"])) < 2e-4 * float(Z["penalty"])


def test_abort_word_reaches_the_trainer_and_protects_the_weights():
    """A persistent kernel that gives up polling raises the handle's abort word.  The trainer must (a) not let the garbage
    gradients of that meta-step reach theta or the optimizer state and (b) raise -- without a per-pass host sync."""
    from msa_tts_b200.maml import MAML
    cfg = pkg.small_params()
    B, T, L = 3, 8, 7
    tasks = {f"spk{i}": synth.make_task(cfg, B, T, L, 60 + i) for i in range(2)}
    sgd = lambda lr: {"optimizer_name": "SGD", "optim_params": {"lr": str(lr)}}
    adam = {"optimizer_name": "Adam", "optim_params": {"lr": "0.01"}}
    tr = MAML(model=cfg, criterion={"criterion_type": "Tacotron2Loss", **CRIT}, optim_inner=sgd(0.05), optim_outer=adam,
              n_inner_train=1, track_higher_grads=False, clip_grad_norm=True, grad_clip_thresh=1.0)
    tr._metatrain_step(tasks)                     # a healthy step
    tr.engine.abort_flush()
    theta, m, v = tr.theta.clone(), tr.outer_m.clone(), tr.outer_v.clone()
    tr.engine.debug_raise_abort()                 # exactly what a polling thread that timed out does
    log = tr._metatrain_step(tasks)
    assert not bool(torch.isfinite(log["grad_sumsq"]).all()), "the gradient norm of the aborted step is poisoned on the device"
    assert torch.equal(tr.theta, theta) and torch.equal(tr.outer_m, m) and torch.equal(tr.outer_v, v), "update skipped"
    with pytest.raises(RuntimeError, match="polling time-out"):
        tr._metatrain_step(tasks)                 # the deferred host check fires on the next step (or at the end of the epoch)
    tr.engine.abort_flush()
    tr._metatrain_step(tasks)                     # cleared: training goes on
    tr.engine.abort_flush()
    assert not torch.equal(tr.theta, theta)
    with pytest.raises(RuntimeError, match="polling"):
        tr.engine.debug_raise_abort()
        tr.engine.check_abort()
    tr.engine.abort_clear()
    tr.engine.check_abort()
