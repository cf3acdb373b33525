"""Drop-in for ``msa_tts.models.tacotron2nv.Tacotron2NV`` on the CUDA hot path.

Same constructor (``Tacotron2NV(params)`` with the ``params["model"]`` dictionary, tacotron2nv.py:11-66), same
``forward(inputs=, input_lengths=, melspecs=, melspec_lengths=, speaker_vecs=)`` -> ``[mel, mel_post, gate, align]``
(tacotron2nv.py:81-127), same ``infer(inputs, input_lengths, speaker_vecs)`` (130-162) and the same
``named_parameters()`` / ``state_dict()`` keys and shapes, so reference checkpoints load and the
stock ``torch.optim`` / ``clip_grad_norm_`` calls of the trainers keep working.

What is different underneath: every parameter is a VIEW into one flat fp32 buffer (``model.flat``) and every ``.grad`` a
view into ``model.grad_flat``; the forward / backward are single calls into libmsa_b200.so through a
``torch.autograd.Function``.  There is no PyTorch implementation of the model here and no CPU fallback.

Restrictions (loud errors, never silent fallbacks): training-mode forward only (the reference never runs the teacher-forced
pass in ``eval()`` mode on this path; ``infer`` is implemented), one backward per forward.  ``ForwardAttention`` is supported
with every switch the reference has: ``norm``, ``forward_attn`` and ``trans_agent`` in training and inference, ``windowing`` and
``forward_attn_mask`` (eval-only features) in ``infer``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn

from .engine import Engine
from .layout import FlatLayout


class _Node(nn.Module):
    """Container that only gives parameters / buffers their reference names (encoder.convolutions.0.0.conv.weight ...)."""


def _descend(root: nn.Module, parts: List[str]) -> nn.Module:
    node = root
    for p in parts:
        if not hasattr(node, p):
            node.add_module(p, _Node())
        node = getattr(node, p)
    return node


class _PassFn(torch.autograd.Function):
    """One teacher-forced pass: forward = msa_train_forward, backward = msa_train_backward (whole-model granularity)."""

    @staticmethod
    def forward(ctx, owner, flat, bn, batch_dev, masks, *params):
        out, _ = owner.engine.forward(flat, bn, batch_dev, masks, outputs=True)
        ctx.owner, ctx.flat, ctx.n, ctx.ticket = owner, flat, len(params), owner._new_ticket()
        ctx.mark_non_differentiable(out[3])
        return tuple(out)

    @staticmethod
    def backward(ctx, d_mel, d_post, d_gate, _d_align):
        owner = ctx.owner
        if ctx.ticket != owner._ticket:
            raise RuntimeError("Tacotron2NV: backward must follow its own forward (the pass workspace holds one pass)")
        eng = owner.engine
        z = lambda g, like: torch.zeros_like(like) if g is None else g.contiguous()
        outs = eng._last_out
        g = eng.new_flat(None)
        eng.backward(ctx.flat, g, accumulate=False, scale=1.0, d_outputs=(z(d_mel, outs[0]), z(d_post, outs[1]), z(d_gate, outs[2])))
        return (None, None, None, None, None) + tuple(owner.layout_views(g))


class Tacotron2NV(nn.Module):
    def __init__(self, params: dict, device: Optional[torch.device] = None, criterion: Optional[dict] = None,
                 gemm_tf32: int = 0, init_seed: int = 0):
        super().__init__()
        self.params = params
        crit = criterion or {"reduction": "none", "pos_weight": 10.0}
        self.engine = Engine(params, device, reduction=crit["reduction"], pos_weight=crit["pos_weight"], gemm_tf32=gemm_tf32)
        self.layout: FlatLayout = self.engine.layout
        from .synth import init_params
        self.flat = self.engine.flat_from_dict(init_params(params, init_seed))
        self.grad_flat = self.engine.new_flat()
        self.grad_flat_set = False
        self.bn_flat = self.engine.new_bn_stats()
        self._ticket = 0
        self._mask_seed, self._mask_calls = 1234, 0
        self.injected_masks = None               # parity tests: reference-layout mask dict for the next forward
        for name, view in self.engine.dict_from_flat(self.flat).items():
            parts = name.split(".")
            _descend(self, parts[:-1]).register_parameter(parts[-1], nn.Parameter(view))
        for name, view in self.engine.bn_dict(self.bn_flat).items():
            parts = name.split(".")
            _descend(self, parts[:-1]).register_buffer(parts[-1], view)
        for name in self.layout.bn_names:
            _descend(self, name.split(".")).register_buffer("num_batches_tracked", torch.zeros((), dtype=torch.long))

    # ---- plumbing ---------------------------------------------------------------------------------
    def _apply(self, fn, recurse: bool = True):
        """``model.to(self.device)`` of the reference trainers (metatrainer.py:56, baseline.py:64) is a no-op here: the model is born
        on its GPU.  A conversion that would move the parameters to another device or dtype (``.cpu()``, ``.half()``, another GPU)
        would silently detach them from the flat buffer the kernels read, so it is a loud error instead."""
        probe = fn(self.flat[:1])
        if probe.device != self.flat.device or probe.dtype != self.flat.dtype:
            raise RuntimeError(f"Tacotron2NV lives in one flat fp32 buffer on {self.flat.device}: it cannot be converted to "
                               f"{probe.device} / {probe.dtype} (build it with device=... instead)")
        return super()._apply(fn, recurse)

    def _new_ticket(self) -> int:
        self._ticket += 1
        return self._ticket

    def layout_views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        """Per-parameter views of a flat buffer, in ``parameters()`` order, tagged with their flat base."""
        views = list(self.engine.dict_from_flat(flat).values())
        for v in views:
            v._msa_flat = flat
            v._msa_owner = self
        return views

    def bind_grads(self) -> None:
        """Point every ``p.grad`` at its view of ``grad_flat`` (so stock clip_grad_norm_ / optimizers see the fused result)."""
        for p, g in zip(self.parameters(), self.engine.dict_from_flat(self.grad_flat).values()):
            p.grad = g

    def zero_grad(self, set_to_none: bool = True) -> None:  # noqa: D102
        super().zero_grad(set_to_none=set_to_none)
        self.grad_flat_set = False

    def count_bn_batches(self, n: int = 1) -> None:
        """nn.BatchNorm1d bumps ``num_batches_tracked`` on every train-mode forward (checkpoint parity with the reference)."""
        for name in self.layout.bn_names:
            _descend(self, name.split(".")).num_batches_tracked += n

    def _masks(self, B: int, T: int, L: int) -> torch.Tensor:
        if self.injected_masks is not None:
            return self.engine.pack_masks(self.injected_masks, B, T, L)
        self._mask_calls += 1
        return self.engine.generate_masks(B, T, L, self._mask_seed * 1000003 + self._mask_calls)

    # ---- the reference call surface -------------------------------------------------------------------
    def forward(self, inputs, input_lengths, melspecs, melspec_lengths, speaker_vecs):
        if not self.training:
            raise NotImplementedError("teacher-forced forward in eval() mode is not implemented on the CUDA path; use infer()")
        dev = self.engine.device
        bd = {"inputs": inputs.to(dev).contiguous(), "input_lengths": input_lengths.to(dev).contiguous(),
              "melspecs": melspecs.to(dev).contiguous(), "melspec_lengths": melspec_lengths.to(dev).contiguous(),
              "speaker_vecs": speaker_vecs.to(dev).contiguous()}
        B, L = bd["inputs"].shape
        T = bd["melspecs"].shape[2]
        out = _PassFn.apply(self, self.flat, self.bn_flat, bd, self._masks(B, T, L), *self.parameters())
        self.count_bn_batches()
        self.engine._last_out = out
        return list(out)

    def infer(self, inputs, input_lengths, speaker_vecs, prenet_masks: Optional[torch.Tensor] = None):
        cfg = self.params
        B = inputs.shape[0]
        steps = cfg["max_decoder_steps"]
        if prenet_masks is None:
            g = torch.Generator().manual_seed(self._mask_seed + self._mask_calls)
            prenet_masks = (torch.rand(steps, 2, B, cfg["prenet_dim"], generator=g) >= 0.5)
        with torch.no_grad():
            return self.engine.infer(self.flat, self.bn_flat, inputs, input_lengths, speaker_vecs, prenet_masks, steps)
