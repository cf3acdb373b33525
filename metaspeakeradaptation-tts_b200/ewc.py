"""Drop-in for the EWC pieces of msa_tts/continual_ewc.py on flat buffers.

``EWC(model, buffer_batches, criterion)`` snapshots the means mu = theta and accumulates the diagonal Fisher
F = sum_batches grad(loss_batch)^2 / n_batches (continual_ewc.py:28-41,59-82: the gradient of the BATCH loss is squared, not
per-sample gradients); ``penalty(model)`` = sum F (theta - mu)^2 (84-89, no 1/2); ``sgd_step`` is the fused training update
theta -= lr * (g + 2*lam*F*(theta-mu)) that also returns the penalty (345-357).  All three are single fused launches over
the flat buffers (msa_ewc_*)."""
from __future__ import annotations

from typing import Iterable, Optional

import torch


class EWC:
    def __init__(self, model, buffer_batches: Iterable[tuple], criterion=None, device=None, masks: Optional[list] = None):
        self.model = model
        eng = model.engine
        self.means = model.flat.detach().clone()                         # continual_ewc.py:33-36
        self.fisher = eng.new_flat()
        batches = list(buffer_batches)
        g = eng.new_flat(None)
        from .engine import batch_to_device
        for i, batch in enumerate(batches):                              # continual_ewc.py:59-82
            bd = batch_to_device(batch, eng.device, model.params["speaker_emb_type"])
            B, L = bd["inputs"].shape
            T = bd["melspecs"].shape[2]
            mk = eng.pack_masks(masks[i], B, T, L) if masks is not None else model._masks(B, T, L)
            # the reference runs the model itself in train mode here (continual_ewc.py:66-72): its BatchNorm running statistics
            # and num_batches_tracked DO move with every buffer batch, and the checkpoint / inference after EWC see them
            eng.forward(model.flat, model.bn_flat, bd, mk, outputs=False)
            model.count_bn_batches()
            eng.backward(model.flat, g)
            eng.ewc_fisher_accum(self.fisher, g, 1.0 / len(batches), init=(i == 0))

    def penalty(self, model=None) -> torch.Tensor:
        m = model or self.model
        return m.engine.ewc_penalty(m.flat, self.means, self.fisher).reshape(())

    def add_penalty_grad(self, grads_flat: torch.Tensor, importance: float) -> torch.Tensor:
        """grads += d(importance * penalty)/d theta = 2*importance*F*(theta - mu) (what ``loss += ewc_importance * penalty`` adds to
        ``loss.backward()``, continual_ewc.py:345-355) for optimizers other than plain SGD; returns the penalty (fused, one pass)."""
        m = self.model
        return m.engine.ewc_penalty_grad(m.flat, grads_flat, self.means, self.fisher, importance).reshape(())

    def sgd_step(self, grads_flat: torch.Tensor, lr: float, importance: float) -> torch.Tensor:
        """theta -= lr * (g + 2*importance*F*(theta - mu)); returns the penalty at the OLD theta (fused, one pass)."""
        m = self.model
        return m.engine.ewc_sgd_step(m.flat, grads_flat, self.means, self.fisher, lr, importance).reshape(())
