"""Drop-in for the part of ``higher`` the reference trainers use (maml.py:40-54,71-74; reptile.py:40-56; infer.py:266-293):

    with innerloop_ctx(model, inner_optimizer, track_higher_grads=False) as (fmodel, diffopt):
        outputs = fmodel(**inputs); loss = criterion(...); diffopt.step(loss)
        grads = torch.autograd.grad(loss_test, fmodel.parameters(time=-1))

``fmodel`` is a functional copy: it clones the parameters AND the BatchNorm buffers of ``model`` (higher's semantics,
SURVEY.md Appendix C), so the base model is never touched.  The fast weights live in ONE flat buffer; ``diffopt.step`` is
autograd.grad of the loss w.r.t. that flat leaf followed by the fused functional SGD or Adam kernel (msa_flat_sgd_step,
msa_flat_adam_step; the reference builds the inner optimizer from YAML, utils/helpers.py:20-26).  Only the
first-order mode exists (``track_higher_grads=False``): second-order MAML needs a double backward through the fused kernels.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List

import torch

from .model import Tacotron2NV, _PassFn


class FunctionalTacotron2NV:
    def __init__(self, model: Tacotron2NV):
        self.model = model
        self.engine = model.engine
        self.training = model.training
        self._fast: List[torch.Tensor] = []      # flat fast weights, time 0 .. -1
        self._views: List[List[torch.Tensor]] = []  # per-parameter LEAF views of each (what autograd.grad is asked about)
        self._push(model.flat.detach().clone())
        self.bn_flat = model.bn_flat.clone()
        self.injected_masks = None
        self._calls = 0

    def _push(self, flat: torch.Tensor) -> None:
        views = [v.requires_grad_(True) for v in self.model.layout_views(flat.detach())]
        for v in views:
            v._msa_flat = flat
            v._msa_owner = self.model
        self._fast.append(flat)
        self._views.append(views)

    def __call__(self, inputs, input_lengths, melspecs, melspec_lengths, speaker_vecs):
        m = self.model
        if not self.training:      # same loud error as the base model: there is no eval-mode teacher-forced pass on the CUDA path
            raise NotImplementedError("teacher-forced forward in eval() mode is not implemented on the CUDA path; use infer()")
        dev = self.engine.device
        bd = {"inputs": inputs.to(dev).contiguous(), "input_lengths": input_lengths.to(dev).contiguous(),
              "melspecs": melspecs.to(dev).contiguous(), "melspec_lengths": melspec_lengths.to(dev).contiguous(),
              "speaker_vecs": speaker_vecs.to(dev).contiguous()}
        B, L = bd["inputs"].shape
        T = bd["melspecs"].shape[2]
        if self.injected_masks is not None:
            masks = self.engine.pack_masks(self.injected_masks[self._calls], B, T, L)
        else:
            masks = m._masks(B, T, L)
        self._calls += 1
        out = _PassFn.apply(m, self._fast[-1], self.bn_flat, bd, masks, *self._views[-1])
        self.engine._last_out = out
        return list(out)

    def infer(self, inputs, input_lengths, speaker_vecs, prenet_masks=None):
        """``fmodel.eval(); fmodel.infer(...)`` of the few-shot adaptation script (infer.py:266-293, tacotron2nv.py:130-162): free-running
        decoding with the ADAPTED weights and with the functional copy's BatchNorm running statistics -- the ones the adaptation
        passes (train mode) have been moving, not the base model's (higher clones the buffers)."""
        cfg = self.model.params
        steps = cfg["max_decoder_steps"]
        if prenet_masks is None:
            g = torch.Generator().manual_seed(self.model._mask_seed + 7919 * (self._calls + 1))
            prenet_masks = (torch.rand(steps, 2, inputs.shape[0], cfg["prenet_dim"], generator=g) >= 0.5)
        with torch.no_grad():
            return self.engine.infer(self._fast[-1], self.bn_flat, inputs, input_lengths, speaker_vecs, prenet_masks, steps)

    def parameters(self, time: int = -1):
        """Leaf views of the fast weights at ``time`` in ``model.parameters()`` order (maml.py:71-74)."""
        return self._views[time]

    def fast_flat(self, time: int = -1) -> torch.Tensor:
        return self._fast[time]

    def state_dict(self):
        sd = {k: v.detach() for k, v in self.engine.dict_from_flat(self._fast[-1]).items()}
        sd.update(self.engine.bn_dict(self.bn_flat))
        base = self.model.state_dict()
        for name in self.model.layout.bn_names:        # every train-mode forward of the copy counts one batch per BatchNorm layer
            k = name + ".num_batches_tracked"
            sd[k] = base[k].detach().cpu() + self._calls
        return sd

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self


class DifferentiableSGD:
    """``diffopt.step(loss)`` with torch.optim.SGD hyper-parameters read from the passed optimizer (maml.py:54)."""

    def __init__(self, fmodel: FunctionalTacotron2NV, opt: torch.optim.Optimizer):
        g = opt.param_groups[0]
        self.h = dict(lr=g["lr"], momentum=g.get("momentum", 0.0), dampening=g.get("dampening", 0.0),
                      weight_decay=g.get("weight_decay", 0.0), nesterov=g.get("nesterov", False))
        self.fmodel = fmodel
        self.buf = fmodel.engine.new_flat() if self.h["momentum"] else None
        self.steps = 0

    def step(self, loss: torch.Tensor):
        fm = self.fmodel
        grads = torch.autograd.grad(loss, fm._views[-1])             # first-order: the graph is not kept
        g = grads[0]._msa_flat                                       # the 61 gradients are views of ONE flat buffer
        new = fm.engine.new_flat(None)
        fm.engine.sgd_step(fm._fast[-1], g, p_out=new, buf=self.buf, first_step=(self.steps == 0), **self.h)
        self.steps += 1
        fm._push(new)                                                # re-leafed like higher with track_higher_grads=False
        return fm.parameters(-1)


class DifferentiableAdam:
    """``diffopt.step(loss)`` for a torch.optim.Adam inner optimizer: fresh moment buffers per context (higher creates the
    optimizer state inside ``innerloop_ctx``), torch.optim.Adam's update rule, first-order (re-leafed) fast weights."""

    def __init__(self, fmodel: FunctionalTacotron2NV, opt: torch.optim.Optimizer):
        g = opt.param_groups[0]
        if g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("inner Adam: amsgrad / maximize are not implemented")
        self.h = dict(lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"]), weight_decay=float(g.get("weight_decay", 0.0)))
        self.fmodel = fmodel
        self.m = torch.zeros_like(fmodel.fast_flat(0))
        self.v = torch.zeros_like(fmodel.fast_flat(0))
        self.steps = 0

    def step(self, loss: torch.Tensor):
        fm = self.fmodel
        grads = torch.autograd.grad(loss, fm._views[-1])
        g = grads[0]._msa_flat
        new = fm.engine.new_flat(None)
        self.steps += 1
        fm.engine.adam_step(fm._fast[-1], g, new, self.m, self.v, step=self.steps, **self.h)
        fm._push(new)
        return fm.parameters(-1)


@contextmanager
def innerloop_ctx(model: Tacotron2NV, opt: torch.optim.Optimizer, copy_initial_weights: bool = True,
                  track_higher_grads: bool = True):
    """Same keyword defaults as ``higher.innerloop_ctx`` (``track_higher_grads=True``): a call that omits the keyword asks for the
    second-order graph like it does with ``higher`` and gets the loud error, never a silent first-order run.  Every reference
    call site passes ``track_higher_grads=self.params["track_higher_grads"]`` (maml.py:40-41, reptile.py:40-41, infer.py:266-267)."""
    if track_higher_grads:
        raise NotImplementedError("second-order MAML (track_higher_grads=True) needs a double backward through the fused kernels")
    fmodel = FunctionalTacotron2NV(model)
    if isinstance(opt, torch.optim.SGD):
        diffopt = DifferentiableSGD(fmodel, opt)
    elif type(opt) is torch.optim.Adam:
        diffopt = DifferentiableAdam(fmodel, opt)
    else:
        raise NotImplementedError(f"inner optimizer {type(opt).__name__}: SGD and Adam are implemented on the fused path")
    yield fmodel, diffopt
