"""CPU restatement of the trainer-level arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows msa_tts/maml.py:33-105, msa_tts/reptile.py:33-89,
msa_tts/utils/grad_utils.py:8-31, msa_tts/continual_ewc.py:28-89,338-357,
msa_tts/continual_erkd.py:73-83 and torch.optim.SGD / Adam.

PARITY UNPINNED for the inner-loop optimizer: it lives in the third-party
library ``higher`` (facebookresearch/higher, last release 0.2.1; the reference
pins no version, ships no copy and has no test at that boundary).  Its semantics
are restated from SURVEY.md Appendix C: fast weights = clone of the parameters
in ``model.parameters()`` order, BN buffers cloned privately, and
``diffopt.step`` = ``autograd.grad`` followed by the torch.optim.SGD rule, with
re-leafing when ``track_higher_grads=False``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import model as M


def unpack_batch(batch):
    """metatrainer.py:95-117 -- static / static+linear speaker vectors."""
    _, inp, inp_len, mels, mel_len, spk_ids, spk_embs, stop = batch
    return dict(inputs=inp, input_lengths=inp_len, melspecs=mels, melspec_lengths=mel_len,
                speaker_vecs=spk_embs), stop


def loss_and_grads(P, cfg, batch, masks, stats, crit, names):
    """One fmodel(**inputs) + criterion + autograd.grad w.r.t. every parameter."""
    x, stop = unpack_batch(batch)
    Pl = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    out = M.forward(Pl, cfg, x["inputs"], x["input_lengths"], x["melspecs"], x["melspec_lengths"],
                    x["speaker_vecs"], masks, stats, True)
    loss = M.loss_fn(out, (x["melspecs"], stop), x["melspec_lengths"], **crit)
    grads = torch.autograd.grad(loss, [Pl[n] for n in names], allow_unused=True)
    grads = [torch.zeros_like(Pl[n]) if g is None else g for n, g in zip(names, grads)]
    return loss.detach(), dict(zip(names, grads)), [o.detach() for o in out]


def sgd_step(P, grads, names, lr, momentum=0.0, weight_decay=0.0, dampening=0.0, nesterov=False, bufs=None):
    """torch.optim.SGD rule as applied by higher's DifferentiableSGD (Appendix C)."""
    newP = dict(P)
    for n in names:
        g = grads[n]
        if weight_decay != 0:
            g = g + weight_decay * P[n]
        if momentum != 0:
            if bufs is None or n not in bufs:
                buf = g.clone()
            else:
                buf = momentum * bufs[n] + (1 - dampening) * g
            if bufs is not None:
                bufs[n] = buf
            g = g + momentum * buf if nesterov else buf
        newP[n] = (P[n] - lr * g).detach()
    return newP


def adapt_task(P0, cfg, task, task_masks, crit, names, n_inner, inner_lr, momentum=0.0, weight_decay=0.0):
    """higher.innerloop_ctx + n_inner x diffopt.step on the train split (maml.py:40-54)."""
    P = {k: v.detach().clone() for k, v in P0.items()}
    stats = M.fresh_bn_stats(P, cfg)          # buffers are cloned privately (Q18); base starts at 0/1
    bufs = {}
    losses = []
    for it in range(n_inner):
        loss, g, _ = loss_and_grads(P, cfg, task["train"], task_masks[it], stats, crit, names)
        P = sgd_step(P, g, names, inner_lr, momentum, weight_decay, bufs=bufs)
        losses.append(float(loss))
    return P, stats, losses


def fomaml_task(P0, cfg, task, task_masks, crit, names, n_inner, inner_lr):
    """First-order MAML task gradient: grad of the test loss at theta_T (maml.py:58-76, track_higher_grads=False)."""
    P, stats, _ = adapt_task(P0, cfg, task, task_masks, crit, names, n_inner, inner_lr)
    loss, g, out = loss_and_grads(P, cfg, task["test"], task_masks[n_inner], stats, crit, names)
    return loss, g, out, P, stats


def maml2_task(P0, cfg, task, task_masks, crit, names, n_inner, inner_lr, weight_decay=0.0):
    """Second-order MAML task gradient d loss_test(theta_n) / d theta_0, differentiated THROUGH the inner SGD steps
    (maml.py:44-54 with track_higher_grads=True, maml.py:70-71: autograd.grad(loss_test, fmodel.parameters(time=0))).
    Plain double backward: the inner gradients are taken with create_graph=True and the fast weights stay in the graph."""
    x_tr, stop_tr = unpack_batch(task["train"])
    x_te, stop_te = unpack_batch(task["test"])
    P_init = {k: v.detach().clone().requires_grad_(True) for k, v in P0.items()}
    P = dict(P_init)
    stats = M.fresh_bn_stats(P0, cfg)
    for it in range(n_inner):
        out = M.forward(P, cfg, x_tr["inputs"], x_tr["input_lengths"], x_tr["melspecs"], x_tr["melspec_lengths"], x_tr["speaker_vecs"],
                        task_masks[it], stats, True)
        loss = M.loss_fn(out, (x_tr["melspecs"], stop_tr), x_tr["melspec_lengths"], **crit)
        g = torch.autograd.grad(loss, [P[n] for n in names], create_graph=True, allow_unused=True)
        newP = dict(P)
        for n, gn in zip(names, g):
            gn = torch.zeros_like(P[n]) if gn is None else gn
            if weight_decay != 0:
                gn = gn + weight_decay * P[n]
            newP[n] = P[n] - inner_lr * gn
        P = newP
    out = M.forward(P, cfg, x_te["inputs"], x_te["input_lengths"], x_te["melspecs"], x_te["melspec_lengths"], x_te["speaker_vecs"],
                    task_masks[n_inner], stats, True)
    loss = M.loss_fn(out, (x_te["melspecs"], stop_te), x_te["melspec_lengths"], **crit)
    grads = torch.autograd.grad(loss, [P_init[n] for n in names], allow_unused=True)
    grads = [torch.zeros_like(P_init[n]) if g is None else g for n, g in zip(names, grads)]
    return loss.detach(), dict(zip(names, grads))


def reptile_task(P0, cfg, task, task_masks, crit, names, n_inner, inner_lr):
    """Reptile task 'gradient' -(theta_T - theta_0) (reptile.py:42,73-77)."""
    P, stats, _ = adapt_task(P0, cfg, task, task_masks, crit, names, n_inner, inner_lr)
    return {n: -(P[n] - P0[n]).detach() for n in names}, P, stats


def mix_grad(grad_list: Sequence[dict], weights, names) -> dict:
    """grad_utils.py:23-31: stack(w_i * g_i).sum(0) per parameter."""
    return {n: torch.stack([weights[i] * grad_list[i][n] for i in range(len(grad_list))]).sum(dim=0) for n in names}


def grad_norm(grads: dict, names) -> float:
    """grad_utils.py:8-20: sqrt(sum_p sum(g**2))."""
    tot = 0.0
    for n in names:
        tot = tot + torch.sum(grads[n] ** 2)
    return float(tot ** 0.5)


def clip_coef(total_norm: float, max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1."""
    return min(1.0, max_norm / (total_norm + 1e-6))


def outer_sgd(P, grads, names, lr, clip=None):
    """maml.py:99-105 with an SGD outer optimizer (momentum 0)."""
    coef = 1.0 if clip is None else clip_coef(grad_norm(grads, names), clip)
    return {n: P[n] - lr * (grads[n] * coef) for n in P}


def outer_adam(P, grads, names, state, lr, betas=(0.9, 0.999), eps=1e-8, clip=None):
    """maml.py:99-105 with torch.optim.Adam (no amsgrad, no weight decay)."""
    coef = 1.0 if clip is None else clip_coef(grad_norm(grads, names), clip)
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    out = {}
    for n in names:
        g = grads[n] * coef
        m = state.setdefault("m", {}).get(n, torch.zeros_like(g))
        v = state.setdefault("v", {}).get(n, torch.zeros_like(g))
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        state["m"][n], state["v"][n] = m, v
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        denom = (v.sqrt() / (bc2 ** 0.5)) + eps
        out[n] = P[n] - (lr / bc1) * (m / denom)
    return out


def ewc_fisher(P, cfg, buffer_batches, buffer_masks, crit, names):
    """EWC._diag_fisher, continual_ewc.py:59-82: F = sum_batches grad**2 / n_batches.

    The reference runs the *model itself* in train mode, so its BN running stats move; the
    gradients do not depend on them."""
    F_ = {n: torch.zeros_like(P[n]) for n in names}
    for batch, masks in zip(buffer_batches, buffer_masks):
        _, g, _ = loss_and_grads(P, cfg, batch, masks, None, crit, names)
        for n in names:
            F_[n] += g[n] ** 2 / len(buffer_batches)
    return F_


def ewc_penalty(P, F_, mu, names):
    """EWC.penalty, continual_ewc.py:84-89: sum F (p - mu)^2 (no 1/2)."""
    tot = 0.0
    for n in names:
        tot = tot + (F_[n] * (P[n] - mu[n]) ** 2).sum()
    return tot


def ewc_step(P, cfg, batch, masks, stats, crit, names, F_, mu, lam, lr):
    """continual_ewc.py:338-357 with an SGD optimizer: loss + lam*penalty, backward, step.
    (clip_grad_norm_ there acts on stale grads before zero_grad -- a no-op)"""
    loss, g, _ = loss_and_grads(P, cfg, batch, masks, stats, crit, names)
    pen = ewc_penalty(P, F_, mu, names)
    g = {n: g[n] + lam * 2.0 * F_[n] * (P[n] - mu[n]) for n in names}
    newP = {n: P[n] - lr * g[n] for n in P}
    return float(loss) + lam * float(pen), g, newP


def mcd_batch(output, mel, mel_len):
    """utils/metrics.py:15-22 on torch tensors [B, T, D]."""
    import math
    K = 10 / math.log(10) * math.sqrt(2)
    vals = []
    for i in range(output.shape[0]):
        d = mel[i, :mel_len[i]] - output[i, :mel_len[i]]
        vals.append(float(torch.sqrt((d ** 2).sum(dim=1)).mean()))
    return K * sum(vals) / len(vals)
