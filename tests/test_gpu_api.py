"""The reference-shaped Python surface (Tacotron2NV / Tacotron2Loss / innerloop_ctx / mix_grad / apply_grad / EWC) on the CUDA
path, written the way the reference trainers call it (maml.py:40-76,94-105; continual_ewc.py:28-89,345-357) and checked
against the oracle.  Tolerance: fp32, 2e-4 (outputs per tensor, gradients relative to the global gradient norm)."""
import os

import pytest
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from oracle import meta as OMeta
from oracle import model as OM
from helpers import oracle_pass, rel

pytestmark = pytest.mark.gpu
CRIT = dict(reduction="none", pos_weight=10.0)
TOL = 2e-4


def _inputs(batch):
    _, inp, inp_len, mels, mel_len, _, spk, stop = batch
    return dict(inputs=inp, input_lengths=inp_len, melspecs=mels, melspec_lengths=mel_len, speaker_vecs=spk), (mels, stop), mel_len


def _model(cfg, P):
    model = pkg.Tacotron2NV(cfg)
    sd = model.state_dict()
    assert [k for k in sd if "running" not in k and "num_batches" not in k] == list(P.keys()), "state_dict keys / order = reference"
    missing, unexpected = model.load_state_dict({k: v for k, v in P.items()}, strict=False)
    assert not unexpected and all("running" in k or "num_batches" in k for k in missing)
    return model


def _gerr(grads, o_g, names):
    gn = float(torch.sqrt(sum((v.double() ** 2).sum() for v in o_g.values())))
    return max(float((g.double().cpu() - o_g[n].double()).norm()) / gn for n, g in zip(names, grads))


def test_module_forward_loss_backward_like_baseline_py():
    cfg = pkg.small_params()
    B, T, L = 3, 11, 9
    P = synth.init_params(cfg, 3)
    batch = synth.make_batch(cfg, B, T, L, 103)
    masks = synth.make_masks(cfg, B, T, L, 203)
    o_out, o_loss, o_g, _, _ = oracle_pass(cfg, P, batch, masks, CRIT)
    model = _model(cfg, P)
    model.train()
    model.injected_masks = masks
    criterion = pkg.Tacotron2Loss(1, CRIT["reduction"], CRIT["pos_weight"])
    kw, targets, mel_len = _inputs(batch)
    out = model(**kw)
    loss = criterion(out, targets, mel_len)
    loss.backward()
    for a, b in zip(out, o_out):
        assert rel(a.detach(), b) < TOL
    assert abs(float(loss.detach()) - float(o_loss)) < TOL * abs(float(o_loss))
    names = list(P.keys())
    assert _gerr([p.grad for p in model.parameters()], o_g, names) < TOL
    # stock torch calls of the trainers keep working on the views (maml.py:101-105)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    before = model.flat.clone()
    opt.step()
    assert not torch.equal(before, model.flat), "optimizer.step() on the per-tensor views updates the flat buffer"
    with pytest.raises(NotImplementedError):
        model.eval()(**kw)


def test_innerloop_fomaml_task_like_maml_py():
    cfg = pkg.small_params()
    B, T, L, lr = 3, 10, 8, 0.05
    P = synth.init_params(cfg, 5)
    task = synth.make_task(cfg, B, T, L, 3)
    masks = [synth.make_masks(cfg, B, T, L, 900 + i) for i in range(2)]
    names = OM.param_names(cfg)
    o_loss, o_g, _, _, _ = OMeta.fomaml_task(P, cfg, task, masks, CRIT, names, 1, lr)
    model = _model(cfg, P)
    criterion = pkg.Tacotron2Loss(1, "none", 10.0)
    inner_opt = torch.optim.SGD(model.parameters(), lr=lr)
    theta0 = model.flat.clone()
    with pkg.innerloop_ctx(model, inner_opt, track_higher_grads=False) as (fmodel, diffopt):      # maml.py:40-41
        fmodel.injected_masks = masks
        kw, targets, mel_len = _inputs(task["train"])
        diffopt.step(criterion(fmodel(**kw), targets, mel_len))                                   # maml.py:50-54
        kw, targets, mel_len = _inputs(task["test"])
        loss_test = criterion(fmodel(**kw), targets, mel_len)
        task_grads = torch.autograd.grad(loss_test, fmodel.parameters(time=-1))                  # maml.py:73-74
    assert torch.equal(theta0, model.flat), "the base model is never touched inside the inner loop"
    assert abs(float(loss_test.detach()) - float(o_loss)) < TOL * abs(float(o_loss))
    assert _gerr(task_grads, o_g, names) < TOL
    # mix_grad / apply_grad (maml.py:94-99, utils/grad_utils.py:8-31)
    from msa_tts_b200.grad_utils import apply_grad, mix_grad
    mixed = mix_grad([task_grads, task_grads], [0.25, 0.75])      # the reference's own signature (maml.py:96)
    assert _gerr(mixed, o_g, names) < TOL
    norm = apply_grad(model, mixed)
    gn = float(torch.sqrt(sum((v.double() ** 2).sum() for v in o_g.values())))
    assert abs(norm - gn) < TOL * gn
    assert _gerr([p.grad for p in model.parameters()], o_g, names) < TOL
    with pytest.raises(NotImplementedError):
        with pkg.innerloop_ctx(model, inner_opt, track_higher_grads=True):
            pass


def test_flat_buffer_gaps_are_zero_even_when_uninitialised():
    """The fused kernels stream over whole flat buffers: the alignment gaps between tensors must never hold garbage (it would
    reach gradient norms, Fisher sums and EWC penalties)."""
    from msa_tts_b200.engine import Engine
    eng = Engine(pkg.small_params())
    junk = torch.full((eng.layout.total + 1024,), float("inf"), device="cuda")      # poison the allocator's free list
    del junk
    t = eng.new_flat(None)
    assert eng._gap_idx.numel() > 0
    assert bool((t[eng._gap_idx] == 0).all())
    for n, v in eng.dict_from_flat(t).items():
        v.zero_()
    assert float(eng.sumsq(t)) == 0.0


def test_innerloop_with_adam_inner_optimizer():
    """The reference builds the inner optimizer from YAML with any torch.optim class (utils/helpers.py:20-26): Adam inner
    steps on the fused path against oracle gradients + torch.optim.Adam on the CPU (two inner steps, then the test loss)."""
    from helpers import oracle_pass
    cfg = pkg.small_params()
    B, T, L = 3, 10, 8
    P = synth.init_params(cfg, 6)
    task = synth.make_task(cfg, B, T, L, 4)
    masks = [synth.make_masks(cfg, B, T, L, 950 + i) for i in range(3)]
    names = OM.param_names(cfg)
    hyper = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-8)
    # CPU side: oracle gradients, torch.optim.Adam on per-tensor leaves (fresh state, as inside innerloop_ctx)
    Pc = {n: torch.nn.Parameter(P[n].clone()) for n in names}
    copt = torch.optim.Adam([Pc[n] for n in names], **hyper)
    for i in range(2):
        cur = dict(P)
        cur.update({n: Pc[n].detach().clone() for n in names})
        _, _, g, _, _ = oracle_pass(cfg, cur, task["train"], masks[i], CRIT)
        for n in names:
            Pc[n].grad = g[n].clone()
        copt.step()
    cur = dict(P)
    cur.update({n: Pc[n].detach().clone() for n in names})
    _, o_loss, o_g, _, _ = oracle_pass(cfg, cur, task["test"], masks[2], CRIT)
    # CUDA side, written like maml.py:40-74
    model = _model(cfg, P)
    criterion = pkg.Tacotron2Loss(1, "none", 10.0)
    inner_opt = torch.optim.Adam(model.parameters(), **hyper)
    with pkg.innerloop_ctx(model, inner_opt, track_higher_grads=False) as (fmodel, diffopt):
        fmodel.injected_masks = masks
        for _ in range(2):
            kw, targets, mel_len = _inputs(task["train"])
            diffopt.step(criterion(fmodel(**kw), targets, mel_len))
        kw, targets, mel_len = _inputs(task["test"])
        loss_test = criterion(fmodel(**kw), targets, mel_len)
        task_grads = torch.autograd.grad(loss_test, fmodel.parameters(time=-1))
        fast = fmodel.state_dict()
    # Adam normalises the step: where the gradient is rounding noise (|g| ~ 1e-9, e.g. rows of unused symbols) the two sides move
    # by +-lr in different directions, so the fast weights are compared against the size of the whole update (sanity bound) and
    # the quantities that matter -- test loss and meta-gradient -- at tolerances relative to their own size
    upd = float(torch.sqrt(sum(((Pc[n].detach() - P[n]).double() ** 2).sum() for n in names)))
    dif = float(torch.sqrt(sum(((fast[n].cpu() - Pc[n].detach()).double() ** 2).sum() for n in names)))
    lerr, gerr = abs(float(loss_test.detach()) - float(o_loss)) / abs(float(o_loss)), _gerr(task_grads, o_g, names)
    print(f"inner Adam: fast-weight diff / update {dif / upd:.3e}, loss rel {lerr:.3e}, meta-grad err/|G| {gerr:.3e}")
    assert dif < 0.15 * upd, (dif, upd)
    assert lerr < TOL and gerr < TOL, (lerr, gerr)


def test_ewc_fisher_penalty_and_fused_step_like_continual_ewc_py():
    cfg = pkg.small_params()
    B, T, L = 3, 9, 8
    P = synth.init_params(cfg, 7)
    names = OM.param_names(cfg)
    buf = [synth.make_batch(cfg, B, T, L, 400 + i) for i in range(3)]
    bmasks = [synth.make_masks(cfg, B, T, L, 500 + i) for i in range(3)]
    F_ = OMeta.ewc_fisher(P, cfg, buf, bmasks, CRIT, names)
    model = _model(cfg, P)
    ewc = pkg.EWC(model, buf, masks=bmasks)
    fd = model.engine.dict_from_flat(ewc.fisher)
    fn = float(torch.sqrt(sum((v.double() ** 2).sum() for v in F_.values())))
    assert max(float((fd[n].double().cpu() - F_[n].double()).norm()) / fn for n in names) < 5e-4
    # move the weights, then penalty and the fused EWC step against the oracle (continual_ewc.py:84-89,345-357)
    g = torch.Generator().manual_seed(1)
    P2 = {n: v + 0.01 * torch.randn(v.shape, generator=g) for n, v in P.items()}
    model.load_state_dict(P2, strict=False)
    pen = ewc.penalty(model)
    o_pen = OMeta.ewc_penalty(P2, F_, P, names)
    assert abs(float(pen) - float(o_pen)) < 1e-3 * abs(float(o_pen))
    batch = synth.make_batch(cfg, B, T, L, 600)
    masks = synth.make_masks(cfg, B, T, L, 700)
    lam, lr = 50.0, 0.01
    stats = OM.fresh_bn_stats(P2, cfg)
    o_total, _, o_new = OMeta.ewc_step(P2, cfg, batch, masks, stats, CRIT, names, F_, P, lam, lr)
    from msa_tts_b200.engine import batch_to_device
    eng = model.engine
    bd = batch_to_device(batch, eng.device)
    _, loss = eng.forward(model.flat, model.bn_flat, bd, eng.pack_masks(masks, B, T, L), outputs=False)
    gflat = eng.new_flat()
    eng.backward(model.flat, gflat)
    pen2 = ewc.sgd_step(gflat, lr, lam)
    assert abs(float(loss) + lam * float(pen2) - float(o_total)) < 1e-3 * abs(float(o_total))
    new = eng.dict_from_flat(model.flat)
    pn = float(torch.sqrt(sum((v.double() ** 2).sum() for v in o_new.values())))
    assert max(float((new[n].double().cpu() - o_new[n].double()).norm()) / pn for n in names) < 1e-5


def test_erkd_soft_targets_like_continual_erkd_py():
    """continual_erkd.py:73-83: the stored soft target is the FIRST model output (pre-postnet mel, Q9) trimmed to length."""
    from msa_tts_b200.continual import make_soft_targets, sgd_train_step
    cfg = pkg.small_params()
    B, T, L = 3, 10, 8
    P = synth.init_params(cfg, 8)
    batch = synth.make_batch(cfg, B, T, L, 808)
    masks = synth.make_masks(cfg, B, T, L, 809)
    o_out, o_loss, o_g, _, _ = oracle_pass(cfg, P, batch, masks, CRIT)
    model = _model(cfg, P)
    soft = make_soft_targets(model, [batch], masks=[masks])
    lens = batch[4].tolist()
    assert list(soft.keys()) == list(batch[0])
    for j, item in enumerate(batch[0]):
        assert soft[item].shape == (cfg["n_mel_channels"], lens[j])
        assert rel(soft[item], o_out[0][j, :, :lens[j]]) < TOL
    # and the plain replay training step (continual_erkd.py:318-336 with SGD)
    before = {k: v.clone() for k, v in model.engine.dict_from_flat(model.flat).items()}
    log = sgd_train_step(model, batch, 0.01, masks=masks)
    assert abs(float(log["loss"]) - float(o_loss)) < TOL * abs(float(o_loss))
    from oracle import meta as OMeta          # the step's log metric (continual_erkd.py:338-342): mcd_batch of the FIRST output
    o_mcd = OMeta.mcd_batch(o_out[0].detach().transpose(1, 2), batch[3].transpose(1, 2), batch[4].tolist())
    assert abs(float(log["mcd"]) - o_mcd) < TOL * abs(o_mcd)
    after = model.engine.dict_from_flat(model.flat)
    names = list(P.keys())
    num = sum(float(((after[n] - (before[n] - 0.01 * o_g[n].to(after[n].device))).double() ** 2).sum()) for n in names)
    den = sum(float(((0.01 * o_g[n]).double() ** 2).sum()) for n in names)
    assert (num / den) ** 0.5 < TOL


def test_er_reg_adaptive_weight_decay_step_like_continual_er_reg_py():
    """continual_er_reg.py:213-216: weight_decay = weightdecay_value * (1 - similarity), then torch.optim.SGD's rule
    g <- g + wd * p; p <- p - lr * g (one fused streaming kernel here)."""
    from msa_tts_b200.continual import adaptive_weight_decay, sgd_train_step
    cfg = pkg.small_params()
    B, T, L = 3, 10, 8
    P = synth.init_params(cfg, 9)
    batch = synth.make_batch(cfg, B, T, L, 818)
    masks = synth.make_masks(cfg, B, T, L, 819)
    _, o_loss, o_g, _, _ = oracle_pass(cfg, P, batch, masks, CRIT)
    wd = adaptive_weight_decay(0.2, 0.75)
    assert abs(wd - 0.05) < 1e-12
    model = _model(cfg, P)
    log = sgd_train_step(model, batch, 0.01, masks=masks, weight_decay=wd)
    assert abs(float(log["loss"]) - float(o_loss)) < TOL * abs(float(o_loss))
    after = model.engine.dict_from_flat(model.flat)
    names = list(P.keys())
    want = {n: P[n] - 0.01 * (o_g[n] + wd * P[n]) for n in names}
    num = sum(float(((after[n].cpu() - want[n]).double() ** 2).sum()) for n in names)
    den = sum(float(((want[n] - P[n]).double() ** 2).sum()) for n in names)
    assert (num / den) ** 0.5 < TOL


def test_few_shot_adaptation_then_infer_like_infer_py():
    """infer.py:258-293 (BASELINE configs[4]): adapt on the speaker's train split inside innerloop_ctx, ``fmodel.eval()``, then
    free-running ``fmodel.infer`` -- with the adapted weights AND the functional copy's BatchNorm running statistics, which the
    train-mode adaptation passes have moved (the base model's stay untouched).  mel_lengths / step count exact, mels 2e-4."""
    cfg = dict(pkg.small_params())
    cfg["max_decoder_steps"], cfg["decoder_no_early_stopping"] = 14, True
    B, T, L, n_inner, lr = 3, 10, 8, 2, 0.05
    P = synth.init_params(cfg, 12)
    task = synth.make_task(cfg, B, T, L, 21)
    masks = [synth.make_masks(cfg, B, T, L, 970 + i) for i in range(n_inner)]
    names = OM.param_names(cfg)
    from oracle import meta as OMeta
    P_T, stats, _ = OMeta.adapt_task(P, cfg, task, masks, CRIT, names, n_inner, lr)
    _, q_inp, q_len, _, _, _, q_spk, _ = synth.make_batch(cfg, 2, 6, 7, 333)
    pm = synth.make_infer_masks(cfg, 2, cfg["max_decoder_steps"], 77)
    o_post, o_lens, o_align = OM.infer(P_T, cfg, q_inp, q_len, q_spk, pm, stats)

    model = _model(cfg, P)
    base_bn = model.bn_flat.clone()
    criterion = pkg.Tacotron2Loss(1, "none", 10.0)
    inner_opt = torch.optim.SGD(model.parameters(), lr=lr)
    with pkg.innerloop_ctx(model, inner_opt, track_higher_grads=False) as (fmodel, diffopt):
        fmodel.injected_masks = masks
        for _ in range(n_inner):
            kw, targets, mel_len = _inputs(task["train"])
            diffopt.step(criterion(fmodel(**kw), targets, mel_len))
        fmodel.eval()
        post, lens, align = fmodel.infer(q_inp, q_len, q_spk, prenet_masks=pm)
        adapted = fmodel.state_dict()
    torch.cuda.synchronize()
    assert post.shape == o_post.shape and torch.equal(lens.cpu(), o_lens)
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL
    assert torch.equal(model.bn_flat, base_bn), "the base model's running statistics must not move (higher clones the buffers)"
    for k, v in stats.items():                                   # the functional copy's did, exactly like the oracle's
        if k.endswith("num_batches_tracked"):
            assert int(adapted[k]) == int(v) == n_inner, k
        else:
            assert rel(adapted[k], v) < TOL, k
    # infer.py:287-288: model_spk.load_state_dict(fmodel.state_dict()) -- the adapted copy is a loadable checkpoint
    spk_model = pkg.Tacotron2NV(cfg)
    missing, unexpected = spk_model.load_state_dict(adapted, strict=False)
    assert not unexpected and not missing
    post2, lens2, _ = spk_model.infer(q_inp, q_len, q_spk, prenet_masks=pm)
    assert torch.equal(post2, post) and torch.equal(lens2, lens)


@pytest.mark.parametrize("optim", [{"optimizer_name": "Adam", "optim_params": {"lr": "1e-3", "betas": "(0.9, 0.98)", "weight_decay": "1e-6"}},
                                   {"optimizer_name": "SGD", "optim_params": {"lr": "0.01", "momentum": "0.9", "nesterov": "True"}}])
def test_ewc_training_steps_with_any_torch_optimizer_like_continual_ewc_py(optim):
    """continual_ewc.py:213 builds ``self.optim = get_optimizer(self.model, **params["optim"])`` from ANY torch.optim class and
    345-357 adds ``ewc_importance * penalty`` to the loss before ``backward()``: two training steps through ``train_step`` (penalty
    gradient fused into one pass over the flat buffers, then the Adam / momentum-SGD kernels) against the oracle's gradients +
    the penalty gradient stepped by the real torch.optim class."""
    from msa_tts_b200.continual import train_step
    from msa_tts_b200.helpers import optimizer_hparams
    cfg = pkg.small_params()
    B, T, L = 3, 9, 8
    P = synth.init_params(cfg, 7)
    names = OM.param_names(cfg)
    buf = [synth.make_batch(cfg, B, T, L, 400 + i) for i in range(2)]
    bmasks = [synth.make_masks(cfg, B, T, L, 500 + i) for i in range(2)]
    F_ = OMeta.ewc_fisher(P, cfg, buf, bmasks, CRIT, names)
    model = _model(cfg, P)
    ewc = pkg.EWC(model, buf, masks=bmasks)
    g = torch.Generator().manual_seed(1)
    P2 = {n: v + 0.01 * torch.randn(v.shape, generator=g) for n, v in P.items()}
    model.load_state_dict(P2, strict=False)
    lam = 50.0
    h = optimizer_hparams(optim)
    kw = {k: v for k, v in h.items() if k != "name"}
    ref_p = {n: P2[n].clone().requires_grad_(True) for n in names}
    ref_opt = getattr(torch.optim, h["name"])([ref_p[n] for n in names], **kw)
    stats = OM.fresh_bn_stats(P2, cfg)
    for step in range(2):
        batch = synth.make_batch(cfg, B, T, L, 600 + step)
        masks = synth.make_masks(cfg, B, T, L, 700 + step)
        cur = {n: ref_p[n].detach().clone() for n in names}
        o_loss, o_g, _ = OMeta.loss_and_grads({**P2, **cur}, cfg, batch, masks, stats, CRIT, names)
        o_pen = OMeta.ewc_penalty(cur, F_, P, names)
        out = train_step(model, batch, optim, ewc=ewc, importance=lam, masks=masks)
        assert abs(float(out["loss"]) - float(o_loss)) < 1e-3 * abs(float(o_loss))
        assert abs(float(out["penalty"]) - float(o_pen)) < 1e-3 * abs(float(o_pen)) + 1e-9
        # (1) the completed gradient = task gradient + penalty gradient, against the oracle
        gd = {n: v.cpu().clone() for n, v in model.engine.dict_from_flat(model.grad_flat).items()}
        want = {n: o_g[n] + 2.0 * lam * F_[n] * (cur[n] - P[n]) for n in names}
        gn = float(torch.sqrt(sum((want[n].double() ** 2).sum() for n in names)))
        assert max(float((gd[n].double() - want[n].double()).norm()) / gn for n in names) < TOL
        # (2) the update rule: the real torch.optim class stepped with the SAME gradient values (Adam turns rounding-level
        # differences of near-zero gradient elements into full-size steps, so the rule is checked apart from the gradient)
        for n in names:
            ref_p[n].grad = gd[n]
        ref_opt.step()
        new = model.engine.dict_from_flat(model.flat)
        moved = float(torch.sqrt(sum(((ref_p[n].detach() - P2[n]).double() ** 2).sum() for n in names)))
        err = float(torch.sqrt(sum(((new[n].cpu() - ref_p[n].detach()).double() ** 2).sum() for n in names)))
        print(f"{h['name']} step {step}: |theta - theta_ref| / |theta_ref - theta_0| = {err / moved:.2e}")
        assert err < 1e-5 * moved


def test_model_to_its_own_device_is_a_noop_and_other_conversions_are_loud():
    """metatrainer.py:56 / baseline.py:64 call ``self.model.to(self.device)``: the parameters must stay views of the flat buffer the
    kernels read; a conversion that would detach them (.cpu(), .half()) raises instead of silently training a stale copy."""
    cfg = pkg.small_params()
    model = _model(cfg, synth.init_params(cfg, 3))
    ptrs = [p.data_ptr() for p in model.parameters()]
    assert model.to(model.engine.device) is model and model.cuda() is model and model.float() is model
    assert [p.data_ptr() for p in model.parameters()] == ptrs
    first = next(iter(model.parameters()))
    assert first.data_ptr() == model.flat.data_ptr() + 4 * model.layout.offsets[model.layout.names()[0]]
    for bad in (model.cpu, model.half, model.double):
        with pytest.raises(RuntimeError):
            bad()
    assert [p.data_ptr() for p in model.parameters()] == ptrs


def test_joint_trainer_train_test_run_like_baseline_py(tmp_path):
    """baseline.py:181-296 (BASELINE configs[0]'s caller): ``_train`` = forward, loss, backward, SGD step per batch; ``_test`` = the
    test loader in train mode without gradients, mean loss / MCD, ``checkpoint_best.pt`` on improvement; ``run`` = the epoch loop.
    Against the oracle on the same batches and injected dropout masks."""
    from msa_tts_b200.baseline import JointTrainer
    cfg = pkg.small_params()
    B, T, L, lr = 3, 10, 8, 0.02
    P0 = synth.init_params(cfg, 5)
    names = list(P0.keys())
    batches = [synth.make_batch(cfg, B, T, L, 910 + i) for i in range(2)]
    test_batches = [synth.make_batch(cfg, B, T, L, 920 + i) for i in range(2)]
    masks = synth.make_masks(cfg, B, T, L, 930)
    tr = JointTrainer(model=cfg, criterion={"criterion_type": "Tacotron2Loss", **CRIT}, init_seed=5, n_epochs=1,
                      optim={"optimizer_name": "SGD", "optim_params": {"lr": str(lr)}}, freeze_charemb=False, freeze_encoder=False,
                      freeze_decoder=False, ckpt_save_epoch_interval=1, output_path=str(tmp_path))
    tr.model.injected_masks = masks
    logs = tr._train(1, batches)
    P, stats = P0, OM.fresh_bn_stats(P0, cfg)
    for i, batch in enumerate(batches):
        o_loss, o_g, o_out = OMeta.loss_and_grads(P, cfg, batch, masks, stats, CRIT, names)
        o_mcd = OMeta.mcd_batch(o_out[0].transpose(1, 2), batch[3].transpose(1, 2), batch[4].tolist())
        assert abs(float(logs[i]["loss"]) - float(o_loss)) < TOL * abs(float(o_loss))
        assert abs(float(logs[i]["mcd"]) - o_mcd) < TOL * abs(o_mcd)
        P = OMeta.sgd_step(P, o_g, names, lr)
    got = tr.model.engine.dict_from_flat(tr.model.flat)
    upd = float(torch.sqrt(sum(((P[n] - P0[n]).double() ** 2).sum() for n in names)))
    assert max(float((got[n].double().cpu() - P[n].double()).norm()) for n in names) < TOL * upd
    assert tr.step_global == 2
    # _test: no gradient step, mean over the batches, best checkpoint with the reference's key set
    before = tr.model.flat.clone()
    t = tr._test(1, test_batches)
    want = [OMeta.loss_and_grads(P, cfg, b, masks, stats, CRIT, names) for b in test_batches]
    o_loss = sum(float(w[0]) for w in want) / 2
    o_mcd = sum(OMeta.mcd_batch(w[2][0].transpose(1, 2), b[3].transpose(1, 2), b[4].tolist()) for w, b in zip(want, test_batches)) / 2
    assert abs(t["loss"] - o_loss) < TOL * abs(o_loss) and abs(t["mcd"] - o_mcd) < TOL * abs(o_mcd)
    assert t["best"] and tr.best_test_loss == t["loss"] and torch.equal(tr.model.flat, before)
    sd = torch.load(os.path.join(str(tmp_path), "checkpoint_best.pt"), map_location="cpu")
    assert [k for k in sd if k in set(names)] == names and torch.equal(sd[names[-1]], got[names[-1]].cpu())
    # run(): the epoch loop; a second trainer finetunes from the checkpoint it wrote
    out = tr.run(batches, test_batches, n_epochs=2)
    assert len(out) == 2 and tr.step_global == 4 and len(tr.train_logs) == 2
    path = os.path.join(str(tmp_path), "checkpoint_0.pt")
    assert os.path.exists(path)
    tr2 = JointTrainer(model=cfg, optim=tr.optim, init_seed=99, finetune=True, finetune_checkpoint_path=path)
    assert torch.equal(tr2.model.flat, tr.model.flat)
