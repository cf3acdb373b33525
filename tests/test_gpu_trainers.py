"""Whole meta-steps through the trainer classes (MAML / Reptile: maml.py:33-105, reptile.py:33-89) against the CPU oracle
(oracle/meta.py) on the same seeded tasks and injected dropout masks: averaged meta-gradient, clipping and the outer update.

Tolerances (fp32 GEMM policy): averaged meta-gradient / Reptile delta 2e-4 of its global norm per tensor, updated weights
2e-4 of the size of the UPDATE (not of the weights -- that would hide a wrong step), test losses 2e-4 relative.  Task order,
task indices and the checkpoint keys are exact.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H  # noqa: E402

pytestmark = pytest.mark.gpu

CRIT = dict(reduction="none", pos_weight=10.0)
TOL = 2e-4
B, T, L = 3, 12, 9


def _params(cfg, inner, outer, n_inner, clip=None, **kw):
    p = {"model": cfg, "criterion": {"criterion_type": "Tacotron2Loss", **CRIT},
         "optim_inner": inner, "optim_outer": outer, "n_inner_train": n_inner, "track_higher_grads": False,
         "clip_grad_norm": clip is not None, "grad_clip_thresh": clip if clip is not None else 0.0, "init_seed": 5}
    p.update(kw)
    return p


def _setup(n_tasks, n_inner):
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from oracle import model as OM
    cfg = pkg.small_params()
    tasks = {f"spk{i}": synth.make_task(cfg, B, T, L, 40 + i) for i in range(n_tasks)}
    masks = {(i, p): synth.make_masks(cfg, B, T, L, 7000 + 16 * i + p) for i in range(n_tasks) for p in range(n_inner + 1)}
    P0 = synth.init_params(cfg, 5)
    return cfg, tasks, masks, P0, OM.param_names(cfg)


def _gnorm(d, names):
    return float(torch.sqrt(sum((d[n].double() ** 2).sum() for n in names)))


def _check_tensors(got, want, names, scale, what):
    worst = max((float((got[n].double().cpu() - want[n].double()).norm()) / scale, n) for n in names)
    assert worst[0] < TOL, f"{what}: {worst[1]} off by {worst[0]:.2e} of the global norm"


def _sgd(lr, **kw):
    return {"optimizer_name": "SGD", "optim_params": {"lr": str(lr), **{k: str(v) for k, v in kw.items()}}}


def _adam(lr):
    return {"optimizer_name": "Adam", "optim_params": {"lr": str(lr)}}


@pytest.mark.parametrize("outer", ["sgd_clip", "adam_clip", "sgd_noclip"])
def test_fomaml_meta_step_matches_oracle(outer):
    """3 speakers, 2 inner SGD steps; meta-gradient = mean of the test-loss gradients at theta_T (maml.py:73-74, 94-98)."""
    from msa_tts_b200.maml import MAML
    from oracle import meta as OMeta
    n_tasks, n_inner, lr_in, lr_out = 3, 2, 0.05, 0.02
    cfg, tasks, masks, P0, names = _setup(n_tasks, n_inner)
    clip = None if outer == "sgd_noclip" else 0.5
    tr = MAML(**_params(cfg, _sgd(lr_in), _adam(lr_out) if outer == "adam_clip" else _sgd(lr_out), n_inner, clip))
    tr.injected_masks = masks
    log = tr._metatrain_step(tasks)
    torch.cuda.synchronize()
    tr.engine.check_abort()

    o_losses, o_grads, o_mcd = [], [], []
    for i, spk in enumerate(tasks):
        loss, g, out, _, _ = OMeta.fomaml_task(P0, cfg, tasks[spk], [masks[(i, p)] for p in range(n_inner + 1)], CRIT, names, n_inner, lr_in)
        o_losses.append(float(loss))
        o_grads.append(g)
        test = tasks[spk]["test"]          # maml.py:78-82: mcd_batch of the FIRST model output against the test mels
        o_mcd.append(OMeta.mcd_batch(out[0].transpose(1, 2), test[3].transpose(1, 2), test[4].tolist()))
    mixed = OMeta.mix_grad(o_grads, [1.0 / n_tasks] * n_tasks, names)
    gn = OMeta.grad_norm(mixed, names)
    assert log["task_index"] == list(range(n_tasks))
    for a, b in zip(log["loss_test"].tolist(), o_losses):
        assert abs(a - b) < TOL * abs(b)
    for a, b in zip(log["mcd"].tolist(), o_mcd):                                # device-side metric of the per-task log
        assert abs(a - b) < TOL * abs(b)
    _check_tensors(tr.engine.dict_from_flat(tr.meta_grad), mixed, names, gn, "meta-gradient")
    assert abs(float(log["grad_sumsq"]) ** 0.5 - gn) < TOL * gn                 # apply_grad's norm (grad_utils.py:8-20)
    assert clip is None or gn > clip, "the clip must be active in this case"
    if outer == "adam_clip":
        # Adam normalises the step, so rounding noise in near-zero gradient entries moves a weight by +-lr: the update rule is
        # checked on the SAME gradient (the CUDA meta-gradient, verified above), the rule itself bit-level in test_gpu_flat.py
        mixed_dev = {n: v.cpu().clone() for n, v in tr.engine.dict_from_flat(tr.meta_grad).items()}
        P1 = OMeta.outer_adam(P0, mixed_dev, names, {}, lr_out, clip=clip)
    else:
        P1 = OMeta.outer_sgd(P0, mixed, names, lr_out, clip=clip)
    theta = tr.engine.dict_from_flat(tr.theta)
    upd = _gnorm({n: P1[n] - P0[n] for n in names}, names)
    _check_tensors(theta, P1, names, upd, "weights after the outer step")
    assert tr.step_global == 1


def test_fomaml_two_meta_steps_adam_state_carries_over():
    """Second meta-step: Adam's moments and step count persist across meta-steps (maml.py:105)."""
    from msa_tts_b200.maml import MAML
    from oracle import meta as OMeta
    n_tasks, n_inner, lr_in, lr_out = 2, 1, 0.05, 0.01
    cfg, tasks, masks, P0, names = _setup(n_tasks, n_inner)
    tr = MAML(**_params(cfg, _sgd(lr_in), _adam(lr_out), n_inner, 1.0))
    tr.injected_masks = masks
    P, state = P0, {}
    for step in range(2):
        tr._metatrain_step(tasks)
        grads = [OMeta.fomaml_task(P, cfg, tasks[s], [masks[(i, p)] for p in range(n_inner + 1)], CRIT, names, n_inner, lr_in)[1]
                 for i, s in enumerate(tasks)]
        mixed = OMeta.mix_grad(grads, [0.5, 0.5], names)
        torch.cuda.synchronize()
        got = {n: v.cpu().clone() for n, v in tr.engine.dict_from_flat(tr.meta_grad).items()}
        _check_tensors(got, mixed, names, OMeta.grad_norm(mixed, names), f"meta-gradient of meta-step {step + 1}")
        Pn = OMeta.outer_adam(P, got, names, state, lr_out, clip=1.0)      # same gradient on both sides (see above)
        upd = _gnorm({n: Pn[n] - P[n] for n in names}, names)
        P = Pn
        _check_tensors(tr.engine.dict_from_flat(tr.theta), P, names, upd, f"weights after meta-step {step + 1}")
    tr.engine.check_abort()


def test_fomaml_inner_adam_and_inner_momentum():
    """The inner optimizer comes from YAML (helpers.py:20-26): Adam with fresh moments per task, SGD with momentum + weight decay."""
    from msa_tts_b200.maml import MAML
    from oracle import meta as OMeta
    n_tasks, n_inner = 2, 2
    cfg, tasks, masks, P0, names = _setup(n_tasks, n_inner)

    def oracle_task(i, spk, step_fn):
        P = {k: v.clone() for k, v in P0.items()}
        stats = H.OM.fresh_bn_stats(P, cfg)
        st = {}
        for it in range(n_inner):
            _, g, _ = OMeta.loss_and_grads(P, cfg, tasks[spk]["train"], masks[(i, it)], stats, CRIT, names)
            P = step_fn(P, g, st)
        return OMeta.loss_and_grads(P, cfg, tasks[spk]["test"], masks[(i, n_inner)], stats, CRIT, names)[1]

    cases = [
        (_adam(3e-3), lambda P, g, st: {k: v.detach() for k, v in OMeta.outer_adam(P, g, names, st, 3e-3).items()}),
        (_sgd(0.05, momentum=0.9, weight_decay=0.01),
         lambda P, g, st: OMeta.sgd_step(P, g, names, 0.05, momentum=0.9, weight_decay=0.01, bufs=st)),
    ]
    for inner, step_fn in cases:
        tr = MAML(**_params(cfg, inner, _sgd(0.02), n_inner, None))
        tr.injected_masks = masks
        tr._metatrain_step(tasks)
        torch.cuda.synchronize()
        tr.engine.check_abort()
        mixed = OMeta.mix_grad([oracle_task(i, s, step_fn) for i, s in enumerate(tasks)], [0.5, 0.5], names)
        _check_tensors(tr.engine.dict_from_flat(tr.meta_grad), mixed, names, OMeta.grad_norm(mixed, names),
                       f"meta-gradient with inner {inner['optimizer_name']}")


def test_reptile_batched_meta_step_matches_oracle():
    """BASELINE configs[2] semantics: every speaker adapts from the same theta_0, deltas averaged, one outer step."""
    from msa_tts_b200.reptile import Reptile
    from oracle import meta as OMeta
    n_tasks, n_inner, lr_in, lr_out = 3, 3, 0.05, 0.5
    cfg, tasks, masks, P0, names = _setup(n_tasks, n_inner)
    tr = Reptile(**_params(cfg, _sgd(lr_in), _sgd(lr_out), n_inner, 0.05, reptile_sequential=False))
    tr.injected_masks = masks
    log = tr._metatrain_step(tasks)
    torch.cuda.synchronize()
    tr.engine.check_abort()
    deltas, o_losses = [], []
    for i, spk in enumerate(tasks):
        mk = [masks[(i, p)] for p in range(n_inner + 1)]
        d, PT, stats = OMeta.reptile_task(P0, cfg, tasks[spk], mk, CRIT, names, n_inner, lr_in)       # reptile.py:42, 73-77
        deltas.append(d)
        o_losses.append(float(OMeta.loss_and_grads(PT, cfg, tasks[spk]["test"], mk[n_inner], stats, CRIT, names)[0]))
    mixed = OMeta.mix_grad(deltas, [1.0 / n_tasks] * n_tasks, names)
    gn = OMeta.grad_norm(mixed, names)
    _check_tensors(tr.engine.dict_from_flat(tr.meta_grad), mixed, names, gn, "averaged Reptile delta")
    for a, b in zip(log["loss_test"].tolist(), o_losses):
        assert abs(a - b) < TOL * abs(b)
    P1 = OMeta.outer_sgd(P0, mixed, names, lr_out, clip=0.05)
    _check_tensors(tr.engine.dict_from_flat(tr.theta), P1, names, _gnorm({n: P1[n] - P0[n] for n in names}, names), "weights")


def test_reptile_sequential_is_the_reference_literal_loop():
    """reptile.py:37-39, 82-89: an outer step after EACH speaker, the next one starts from the updated weights."""
    from msa_tts_b200.reptile import Reptile
    from oracle import meta as OMeta
    n_tasks, n_inner, lr_in, lr_out = 2, 2, 0.05, 0.5
    cfg, tasks, masks, P0, names = _setup(n_tasks, n_inner)
    # no flag: on one GPU the reference's per-speaker outer steps are the default (the batched variant is opt-in)
    tr = Reptile(**_params(cfg, _sgd(lr_in), _sgd(lr_out), n_inner, None))
    assert tr.sequential
    tr.injected_masks = masks
    tr._metatrain_step(tasks)
    torch.cuda.synchronize()
    tr.engine.check_abort()
    P = P0
    total = {n: torch.zeros_like(P0[n]) for n in names}
    for i, spk in enumerate(tasks):
        d, _, _ = OMeta.reptile_task(P, cfg, tasks[spk], [masks[(i, p)] for p in range(n_inner + 1)], CRIT, names, n_inner, lr_in)
        Pn = OMeta.outer_sgd(P, d, names, lr_out)
        total = {n: total[n] + (Pn[n] - P[n]) for n in names}
        P = Pn
    _check_tensors(tr.engine.dict_from_flat(tr.theta), P, names, _gnorm(total, names), "weights after the sequential loop")
    assert tr.step_global == n_tasks


def test_checkpoint_round_trip_keeps_the_reference_state_dict_keys(tmp_path):
    """metatrainer.py:119-122, 138-146: torch.save(model.state_dict()) / load_state_dict with the reference's key set."""
    from msa_tts_b200.maml import MAML
    n_tasks, n_inner = 2, 1
    cfg, tasks, masks, P0, names = _setup(n_tasks, n_inner)
    tr = MAML(**_params(cfg, _sgd(0.05), _sgd(0.02), n_inner, None, output_path=str(tmp_path)))
    tr.injected_masks = masks
    tr._metatrain_step(tasks)
    path = tr._save_checkpoint()
    sd = torch.load(path, map_location="cpu")
    assert [k for k in sd if k in set(names)] == names                     # parameter keys, model.parameters() order
    extra = set(sd) - set(names)
    assert extra and all(k.endswith(("running_mean", "running_var", "num_batches_tracked")) for k in extra)
    tr2 = MAML(**_params(cfg, _sgd(0.05), _sgd(0.02), n_inner, None, finetune=True, finetune_checkpoint_path=path, init_seed=99))
    assert torch.equal(tr2.theta, tr.theta)                                  # bit-exact through the file
    # metatrainer.py:141-146: a tensor the checkpoint lacks is reported and keeps its initial value, the others still load
    sd.pop(names[0])
    torch.save(sd, path)
    tr3 = MAML(**_params(cfg, _sgd(0.05), _sgd(0.02), n_inner, None, finetune=True, finetune_checkpoint_path=path, init_seed=99))
    ref99 = MAML(**_params(cfg, _sgd(0.05), _sgd(0.02), n_inner, None, init_seed=99))
    d3, d99, d1 = (t.engine.dict_from_flat(t.theta) for t in (tr3, ref99, tr))
    assert torch.equal(d3[names[0]], d99[names[0]]) and all(torch.equal(d3[n], d1[n]) for n in names[1:])


def test_meta_step_on_collated_ragged_batches_from_pinned_memory():
    """The data contract end to end: raw items -> MetaCollator (pinned host memory, the reference's tuple layout,
    dataloader_meta.py:133-179) -> MAML._metatrain_step, with different B / T / L for the train and the test split of each speaker,
    against the oracle on the same collated tuples."""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from msa_tts_b200.data import MetaCollator
    from msa_tts_b200.maml import MAML
    from oracle import meta as OMeta
    from oracle import model as OM
    from oracle.gen_cases import collate_items
    cfg = pkg.small_params()
    kw = dict(n_mels=cfg["n_mel_channels"], n_symbols=cfg["n_symbols"], spk_dim=cfg["speaker_embedding_dim"])
    raw = [("spkA", {"train": collate_items(70, 4, **kw), "test": collate_items(71, 3, **kw)}),
           ("spkB", {"train": collate_items(72, 2, **kw), "test": collate_items(73, 4, **kw)})]
    items = MetaCollator(cfg["n_frames_per_step"], pin_memory=True)(raw)
    assert all(t.is_pinned() for spk in items.values() for b in spk.values() for t in b[1:])
    masks = {}
    for i, spk in enumerate(items):
        for p, mode in enumerate(("train", "test")):
            b = items[spk][mode]
            masks[(i, p)] = synth.make_masks(cfg, b[1].shape[0], b[3].shape[2], b[1].shape[1], 9000 + 8 * i + p)
    tr = MAML(**_params(cfg, _sgd(0.05), _sgd(0.02), 1, None))
    tr.injected_masks = masks
    log = tr._metatrain_step(items)
    torch.cuda.synchronize()
    tr.engine.check_abort()
    names = OM.param_names(cfg)
    P0 = synth.init_params(cfg, 5)
    grads, losses = [], []
    for i, spk in enumerate(items):
        loss, g, _, _, _ = OMeta.fomaml_task(P0, cfg, items[spk], [masks[(i, 0)], masks[(i, 1)]], CRIT, names, 1, 0.05)
        grads.append(g)
        losses.append(float(loss))
    mixed = OMeta.mix_grad(grads, [0.5, 0.5], names)
    for a, b in zip(log["loss_test"].tolist(), losses):
        assert abs(a - b) < TOL * abs(b)
    _check_tensors(tr.engine.dict_from_flat(tr.meta_grad), mixed, names, OMeta.grad_norm(mixed, names), "meta-gradient")


@pytest.mark.parametrize("mode", ["grouped", "plain", "outputs"])
def test_metatest_matches_oracle_and_moves_nothing(mode):
    """maml.py:115-179 / reptile.py:108-172: ``n_inner_test`` inner steps on the train split, then the test loss and MCD of the adapted
    weights without a gradient; theta, the base BatchNorm statistics, the outer optimizer state and step_global stay untouched.
    grouped: the three speakers share grouped passes; plain: speaker by speaker; outputs: the test-pass outputs come back too."""
    from msa_tts_b200.maml import MAML
    from msa_tts_b200.reptile import Reptile
    from oracle import meta as OMeta
    n_tasks, n_inner_test, lr_in = 3, 2, 0.05
    cfg, tasks, _, P0, names = _setup(n_tasks, 1)
    p0 = MAML.METATEST_PASS0
    import msa_tts_b200.synth as synth
    masks = {(i, p0 + p): synth.make_masks(cfg, B, T, L, 9000 + 16 * i + p) for i in range(n_tasks) for p in range(n_inner_test + 1)}
    cls = Reptile if mode == "plain" else MAML
    tr = cls(**_params(cfg, _sgd(lr_in), _adam(0.01), 1, 1.0, n_inner_test=n_inner_test, group_tasks=(mode != "plain")))
    tr.injected_masks = masks
    theta0, bn0 = tr.theta.clone(), tr.base_bn.clone()
    if mode == "outputs":
        log = tr._metatest_step(tasks, return_outputs=True)
    else:
        log = tr._metatest(1, [tasks])[0]
    torch.cuda.synchronize()
    tr.engine.check_abort()
    assert log["task_index"] == list(range(n_tasks)) and log["speakers"] == list(tasks)
    for i, spk in enumerate(tasks):
        loss, _, out, _, _ = OMeta.fomaml_task(P0, cfg, tasks[spk], [masks[(i, p0 + p)] for p in range(n_inner_test + 1)], CRIT, names,
                                               n_inner_test, lr_in)
        test = tasks[spk]["test"]
        mcd = OMeta.mcd_batch(out[0].transpose(1, 2), test[3].transpose(1, 2), test[4].tolist())
        assert abs(float(log["loss_test"][i]) - float(loss)) < TOL * abs(float(loss))
        assert abs(float(log["mcd"][i]) - mcd) < TOL * abs(mcd)
        if mode == "outputs":
            for a, b in zip(log["outputs"][i], out):
                assert float((a.cpu().double() - b.detach().double()).norm()) < TOL * float(b.detach().double().norm())
    assert torch.equal(tr.theta, theta0) and torch.equal(tr.base_bn, bn0) and tr.step_global == 0
    assert float(tr.outer_m.abs().max()) == 0.0 and float(tr.outer_v.abs().max()) == 0.0


def test_run_epoch_loop_like_the_reference(tmp_path):
    """maml.py:19-31: per epoch the meta-train loader, a checkpoint every ckpt_save_epoch_interval and a meta-test every
    metatest_epoch_interval epochs."""
    from msa_tts_b200.maml import MAML
    cfg, tasks, _, _, _ = _setup(2, 1)
    tr = MAML(**_params(cfg, _sgd(0.05), _sgd(0.02), 1, None, n_epochs=2, ckpt_save_epoch_interval=2, metatest_epoch_interval=2,
                        n_inner_test=1, output_path=str(tmp_path)))
    tr.dataloader_metatrain = [tasks, tasks]
    tr.dataloader_metatest = [tasks]
    logs = tr.run()
    assert len(logs) == 4 and tr.step_global == 4
    assert os.path.exists(os.path.join(str(tmp_path), "checkpoint_0.pt"))          # step_global // 100 (metatrainer.py:120)
    assert len(tr.last_metatest) == 1 and tr.last_metatest[0]["loss_test"].numel() == 2
    # a second run() restarts the step counter of the logs (maml.py:20) but not the outer optimizer's own state / step count
    tr.run(n_epochs=1)
    assert tr.step_global == 2 and tr._outer_steps == 6
