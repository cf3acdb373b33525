"""FOMAML meta-step on the host CPU with the UNMODIFIED reference model, loss and outer-update helpers
(TEST / BENCH INFRASTRUCTURE: used by bench.py's ``--impl reference`` arm and its ``cpu_baseline`` leg only).

What is the reference's own code here (imported from baseline/_ref, installed by oracle/install_reference.py, or from
/root/reference in the build container): ``Tacotron2NV`` (models/tacotron2nv.py), ``Tacotron2Loss``
(models/modules_tacotron2nv/tacotron2nv_loss.py), ``mix_grad`` / ``apply_grad`` (utils/grad_utils.py), and the stock
``clip_grad_norm_`` + ``torch.optim`` step of maml.py:94-105.  What is restated: the inner loop of maml.py:38-76, because it
lives in ``higher`` (absent): a functional copy = ``copy.deepcopy(model)`` (parameters AND BatchNorm buffers cloned, SURVEY.md
Appendix C), ``diffopt.step`` = ``autograd.grad`` + the SGD rule in place, first-order task gradient = ``autograd.grad`` of the
test loss w.r.t. the adapted weights.  Dropout draws from torch's RNG exactly as the reference does (timing only, no parity).
If the reference package cannot be imported the port (oracle/model.py, oracle/meta.py) is used and ``kind`` says "port".
"""
from __future__ import annotations

import copy
import io
import os
import sys
import time
from contextlib import redirect_stdout
from typing import Callable, Tuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_threads() -> int:
    """Threads for the CPU arms: the cores this process may run on, capped at the physical core count (hyper-threads slow the
    oneDNN / MKL GEMMs down: SCALE_r01 saw 0.037 meta-steps/s on 32 logical CPUs against 0.058 on 16)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys)
    except Exception:
        pass
    return max(1, n)


def load_reference():
    """(Tacotron2NV, Tacotron2Loss, mix_grad, apply_grad) of the unmodified reference, or None."""
    for path in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(path, "msa_tts", "models", "tacotron2nv.py")):
            if path not in sys.path:
                sys.path.insert(0, path)
            try:
                with redirect_stdout(io.StringIO()):
                    from msa_tts.models.tacotron2nv import Tacotron2NV
                    from msa_tts.models.modules_tacotron2nv.tacotron2nv_loss import Tacotron2Loss
                    from msa_tts.utils.grad_utils import apply_grad, mix_grad
                return Tacotron2NV, Tacotron2Loss, mix_grad, apply_grad
            except Exception:
                continue
    return None


def make_meta_step(cfg: dict, tasks: list, inner_lr: float, n_inner: int, outer_lr: float, clip: float, n_threads: int):
    """-> (task_fn(i) -> seconds, outer_fn() -> seconds, kind).  ``task_fn(i)`` adapts on task i's train split and appends the
    first-order gradient of its test loss; ``outer_fn`` runs maml.py:94-105 on the collected gradients and clears them."""
    import torch
    torch.set_num_threads(n_threads)
    ref = load_reference()
    crit = dict(reduction="none", pos_weight=10.0)
    if ref is None:
        return _make_port_step(cfg, tasks, inner_lr, n_inner, outer_lr, clip) + ("port",)
    Tacotron2NV, Tacotron2Loss, mix_grad, apply_grad = ref
    from msa_tts_b200 import synth
    with redirect_stdout(io.StringIO()):
        model = Tacotron2NV(copy.deepcopy(cfg))
    sd = model.state_dict()
    for k, v in synth.init_params(cfg, 0).items():
        sd[k] = v.clone()
    model.load_state_dict(sd)
    model.train()
    criterion = Tacotron2Loss(cfg["n_frames_per_step"], crit["reduction"], crit["pos_weight"], "cpu")
    outer = torch.optim.Adam(model.parameters(), lr=outer_lr)
    grad_list = []

    def unpack(batch):
        _, inp, inp_len, mels, mel_len, _, spk, stop = batch
        return dict(inputs=inp, input_lengths=inp_len, melspecs=mels, melspec_lengths=mel_len, speaker_vecs=spk), stop

    def task_fn(i: int) -> float:
        t0 = time.perf_counter()
        task = tasks[i % len(tasks)]
        fmodel = copy.deepcopy(model)                                   # higher.innerloop_ctx: params + buffers cloned
        params = list(fmodel.parameters())
        x, stop = unpack(task["train"])
        for _ in range(n_inner):                                        # maml.py:49-54
            loss = criterion(fmodel(**x), (x["melspecs"], stop), x["melspec_lengths"])
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            with torch.no_grad():
                for p, g in zip(params, grads):
                    if g is not None:
                        p.add_(g, alpha=-inner_lr)
        x, stop = unpack(task["test"])                                  # maml.py:56-76
        loss_test = criterion(fmodel(**x), (x["melspecs"], stop), x["melspec_lengths"])
        g = torch.autograd.grad(loss_test, params, allow_unused=True)
        grad_list.append([torch.zeros_like(p) if gi is None else gi for p, gi in zip(params, g)])
        return time.perf_counter() - t0

    def outer_fn() -> float:
        t0 = time.perf_counter()
        model.zero_grad()                                               # maml.py:94-105
        weight = torch.ones(len(grad_list))
        weight = weight / torch.sum(weight)
        mixed = mix_grad(grad_list, weight)
        apply_grad(model, mixed)
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        outer.step()
        grad_list.clear()
        return time.perf_counter() - t0

    return task_fn, outer_fn, "reference"


def _make_port_step(cfg, tasks, inner_lr, n_inner, outer_lr, clip) -> Tuple[Callable, Callable]:
    import torch
    from msa_tts_b200 import synth
    from oracle import meta as OMeta
    from oracle import model as OM
    B, L = tasks[0]["train"][1].shape
    T = tasks[0]["train"][3].shape[2]
    crit = dict(reduction="none", pos_weight=10.0)
    names = OM.param_names(cfg)
    state = {"P": synth.init_params(cfg, 0), "adam": {}, "grads": []}
    masks = [synth.make_masks(cfg, B, T, L, 77 + i) for i in range(n_inner + 1)]

    def task_fn(i: int) -> float:
        t0 = time.perf_counter()
        _, g, _, _, _ = OMeta.fomaml_task(state["P"], cfg, tasks[i % len(tasks)], masks, crit, names, n_inner, inner_lr)
        state["grads"].append(g)
        return time.perf_counter() - t0

    def outer_fn() -> float:
        t0 = time.perf_counter()
        n = len(state["grads"])
        mixed = OMeta.mix_grad(state["grads"], [1.0 / n] * n, names)
        state["P"] = OMeta.outer_adam(state["P"], mixed, names, state["adam"], outer_lr, clip=clip)
        state["grads"].clear()
        return time.perf_counter() - t0

    return task_fn, outer_fn
