"""Drop-in for msa_tts/utils/grad_utils.py on flat buffers.

``mix_grad`` / ``apply_grad`` keep the reference signatures (lists of per-parameter tensors).  Gradient lists produced by
this package are views of ONE flat buffer (``Tacotron2NV.layout_views``), so the weighted sum over tasks is one fused launch
per task instead of 61 x N tiny ones (utils/grad_utils.py:23-31 stacks and sums tensor by tensor), and ``apply_grad`` is one
copy / accumulate plus the fused norm."""
from __future__ import annotations

from typing import List, Sequence

import torch


def _flat_of(tensors: Sequence[torch.Tensor]):
    """The flat buffer a list of per-parameter views came from (set by Tacotron2NV.layout_views), else None."""
    return getattr(tensors[0], "_msa_flat", None)


def mix_grad(grad_list, weight_list, model=None) -> List[torch.Tensor]:
    """Weighted sum of per-task gradient lists (grad_utils.py:23-31): sum_i w_i * g_i.  Same signature as the reference
    (``mix_grad(grad_list, weight_list)``, maml.py:96): the owning model travels with the gradient views."""
    flats = [_flat_of(g) for g in grad_list]
    if any(f is None for f in flats):
        raise RuntimeError("mix_grad: pass gradient lists produced by this package (views of flat buffers); "
                           "there is no per-tensor PyTorch fallback")
    model = model if model is not None else getattr(grad_list[0][0], "_msa_owner", None)
    if model is None:
        raise RuntimeError("mix_grad: the gradient views carry no owner (were they produced by Tacotron2NV.layout_views?)")
    eng = model.engine
    acc = eng.new_flat(None)
    for i, f in enumerate(flats):
        eng.axpy(acc, f, float(weight_list[i]), init=(i == 0))
    return model.layout_views(acc)


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
