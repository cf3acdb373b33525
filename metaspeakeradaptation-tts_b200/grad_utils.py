"""Drop-in for msa_tts/utils/grad_utils.py on flat buffers.

``mix_grad`` / ``apply_grad`` keep the reference signatures (lists of per-parameter tensors) for callers that
want them; tensors that are views of flat buffers are combined with ONE fused launch per task instead of
61 x N tiny ones (utils/grad_utils.py:23-31 stacks and sums tensor by tensor)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def _flat_of(tensors: Sequence[torch.Tensor]):
    """The flat buffer a list of per-parameter views came from (set by Engine.param_views), else None."""
    base = getattr(tensors[0], "_msa_flat", None)
    return base


def mix_grad(grad_list, weight_list, engine=None) -> List[torch.Tensor]:
    """Weighted sum of per-task gradient lists (grad_utils.py:23-31)."""
    flats = [_flat_of(g) for g in grad_list]
    if engine is None or any(f is None for f in flats):
        raise RuntimeError("mix_grad: pass gradient lists produced by this package (views of flat buffers) and the Engine; "
                           "there is no per-tensor PyTorch fallback")
    acc = engine.new_flat(None)
    for i, f in enumerate(flats):
        engine.axpy(acc, f, float(weight_list[i]), init=(i == 0))
    return engine.param_views(acc)


def apply_grad(model, grad) -> float:
    """Assign (or accumulate) gradients to the model and return their L2 norm (grad_utils.py:8-20)."""
    eng = model.engine
    gflat = _flat_of(grad)
    if gflat is None:
        raise RuntimeError("apply_grad: gradients must be views of a flat buffer of this package")
    if model.grad_flat_set:
        eng.axpy(model.grad_flat, gflat, 1.0, init=False)
    else:
        model.grad_flat.copy_(gflat)
        model.grad_flat_set = True
    model.bind_grads()
    return float(torch.sqrt(eng.sumsq(gflat)).item())
