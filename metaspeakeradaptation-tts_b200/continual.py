"""Continual-learning steps of the reference on the CUDA hot path (BASELINE.json configs[3]).

* ER-KD soft targets (continual_erkd.py:73-83,105-115): a teacher-forced forward pass under ``no_grad`` -- in TRAIN mode, the
  reference never calls ``eval()`` there -- whose FIRST output (the pre-postnet mel, SURVEY.md Q9: the variable is named
  ``out_post`` but ``Tacotron2NV.forward`` returns ``[mel, mel_post, gate, align]``) trimmed to each item's length becomes the
  stored target of the replay item.
* The training step itself (continual_erkd.py:318-336, continual_ewc.py:338-357) is forward, loss, backward and one optimizer
  step; with EWC the penalty gradient and the SGD update are one fused kernel (``EWC.sgd_step``).  ``train_step`` takes the
  optimizer the way the reference's params.yml names it (``get_optimizer(model, **params["optim"])``, continual_ewc.py:213):
  SGD (momentum / dampening / nesterov / weight decay) or Adam, with the same update rules as torch.optim.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from .engine import batch_to_device


def make_soft_targets(model, batches: Iterable[tuple], masks: Optional[list] = None) -> Dict[str, torch.Tensor]:
    """{item_id: soft mel [n_mel, len]} for every item of the replay batches (continual_erkd.py:73-83)."""
    eng = model.engine
    out: Dict[str, torch.Tensor] = {}
    for i, batch in enumerate(batches):
        item_ids = batch[0]
        bd = batch_to_device(batch, eng.device, model.params["speaker_emb_type"])
        B, L = bd["inputs"].shape
        T = bd["melspecs"].shape[2]
        mk = eng.pack_masks(masks[i], B, T, L) if masks is not None else model._masks(B, T, L)
        outs, _ = eng.forward(model.flat, model.bn_flat, bd, mk, outputs=True)     # train mode: BN batch stats, running stats move
        mel = outs[0]
        lens = bd["melspec_lengths"].tolist()
        for j, item in enumerate(item_ids):
            out[item] = mel[j, :, :lens[j]].detach().cpu()
    return out


def adaptive_weight_decay(weightdecay_value: float, spk_similarity: float) -> float:
    """continual_er_reg.py:213-216 (``regularizaton_method == "adaptive_weightdecay"``): the optimizer's weight decay for the next
    speaker is ``weightdecay_value * (1 - similarity)``; a similarity of exactly 1.0 keeps the configured value (the reference only
    overrides it ``if spk_similarity != 1.0``) -- pass ``None`` as ``weight_decay`` to ``sgd_train_step`` in that case.
    The sibling method "adaptive_weightclipping" (continual_er_reg.py:356-360) clips the STALE gradients of the previous step right
    before ``zero_grad()`` and therefore changes nothing; there is deliberately no kernel for it."""
    return float(weightdecay_value) * (1.0 - float(spk_similarity))


def sgd_train_step(model, batch: tuple, lr: float, ewc=None, importance: float = 0.0, masks: Optional[dict] = None,
                   weight_decay: float = 0.0) -> dict:
    """One continual training step with plain SGD: loss (+ importance * EWC penalty), backward, update; returns the loss and the
    step's MCD log metric as device tensors (no host sync).
    ``weight_decay`` is torch.optim.SGD's (g += wd * p), folded into the same streaming update kernel."""
    eng = model.engine
    bd = batch_to_device(batch, eng.device, model.params["speaker_emb_type"])
    B, L = bd["inputs"].shape
    T = bd["melspecs"].shape[2]
    mk = eng.pack_masks(masks, B, T, L) if masks is not None else model._masks(B, T, L)
    _, loss = eng.forward(model.flat, model.bn_flat, bd, mk, outputs=False)
    mcd = eng.mcd(bd["melspec_lengths"])       # the per-step log metric (baseline.py:217-219, continual_erkd.py:338-342) on the device
    eng.backward(model.flat, model.grad_flat)
    if ewc is not None:
        if weight_decay:
            raise NotImplementedError("EWC step with weight decay: the reference never combines them (continual_ewc.py:338-357)")
        penalty = ewc.sgd_step(model.grad_flat, lr, importance)                     # fused penalty gradient + update
        return {"loss": loss, "mcd": mcd, "penalty": penalty}
    eng.sgd_step(model.flat, model.grad_flat, lr=lr, weight_decay=weight_decay)
    return {"loss": loss, "mcd": mcd}


def train_step(model, batch: tuple, optim: dict, ewc=None, importance: float = 0.0, masks: Optional[dict] = None) -> dict:
    """One continual training step with the optimizer of the reference's ``params["optim"]`` (continual_ewc.py:213, 338-357;
    continual_erkd.py:318-336): ``optim`` = {"name": "SGD" | "Adam", "lr": ..., + that class's keyword arguments}, values may be the
    strings of a params.yml (helpers.py:20-26 evals them).  The optimizer state (momentum buffer / Adam moments, step count) lives
    on the model as flat buffers and is created on first use, like ``torch.optim`` does.  With ``ewc`` the penalty gradient
    2 * importance * F * (theta - mu) is added to the gradient before the update (fused with the plain-SGD update when there is
    no momentum / weight decay)."""
    from .metatrainer import optimizer_hparams
    h = optimizer_hparams(optim) if not isinstance(optim.get("lr", 0.0), float) or "name" not in optim else dict(optim)
    name = h["name"]
    if name not in ("SGD", "Adam"):
        raise NotImplementedError(f"train_step: optimizer {name} (SGD and Adam are implemented)")
    if name == "Adam" and (h.get("amsgrad", False) or h.get("maximize", False)):
        raise NotImplementedError("train_step: Adam with amsgrad / maximize is not implemented")
    eng = model.engine
    plain_sgd = name == "SGD" and not h.get("momentum", 0.0) and not h.get("weight_decay", 0.0)
    if plain_sgd:
        return sgd_train_step(model, batch, h["lr"], ewc=ewc, importance=importance, masks=masks)
    bd = batch_to_device(batch, eng.device, model.params["speaker_emb_type"])
    B, L = bd["inputs"].shape
    T = bd["melspecs"].shape[2]
    mk = eng.pack_masks(masks, B, T, L) if masks is not None else model._masks(B, T, L)
    _, loss = eng.forward(model.flat, model.bn_flat, bd, mk, outputs=False)
    out = {"loss": loss, "mcd": eng.mcd(bd["melspec_lengths"])}
    eng.backward(model.flat, model.grad_flat)
    if ewc is not None:
        out["penalty"] = ewc.add_penalty_grad(model.grad_flat, importance)
    st = model.__dict__.setdefault("_optim_state", {"step": 0})
    if name == "Adam":
        if "m" not in st:
            st["m"], st["v"] = eng.new_flat(), eng.new_flat()
        st["step"] += 1
        eng.adam_step(model.flat, model.grad_flat, model.flat, st["m"], st["v"], lr=h["lr"], step=st["step"],
                      betas=h.get("betas", (0.9, 0.999)), eps=h.get("eps", 1e-8), weight_decay=h.get("weight_decay", 0.0))
    else:
        if h.get("momentum", 0.0) and "buf" not in st:
            st["buf"] = eng.new_flat()
        eng.sgd_step(model.flat, model.grad_flat, lr=h["lr"], momentum=h.get("momentum", 0.0), dampening=h.get("dampening", 0.0),
                     weight_decay=h.get("weight_decay", 0.0), nesterov=h.get("nesterov", False), buf=st.get("buf"),
                     first_step=(st["step"] == 0))
        st["step"] += 1
    return out
