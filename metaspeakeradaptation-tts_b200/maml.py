"""MAML / first-order MAML driver: mirror of msa_tts/maml.py on the CUDA hot path.

``MAML._metatrain_step(items_b)`` is the body of the reference's loop over meta-batches (maml.py:36-105):
for each speaker adapt ``n_inner_train`` steps on the "train" split, evaluate on the "test" split, take the
gradient of the test loss w.r.t. the adapted weights (first-order, ``track_higher_grads=False``: maml.py:73-74),
average over speakers with weights 1/N (94-98), clip (101-103) and take the outer step (105).
The weighted accumulation is fused into the backward pass of the test split (its GEMM epilogues write
``meta_grad += w * g``), speakers are sharded over ranks and ONE allreduce joins them (parallel.py).
"""
from __future__ import annotations

from typing import Dict

import torch

from .metatrainer import MetaTrainer


class MAML(MetaTrainer):
    def __init__(self, **params):
        super().__init__(**params)
        self._bwd_event = None
        if params.get("track_higher_grads", False):
            # maml.py:70-71.  The target is restated and pinned on the CPU (oracle.meta.maml2_task, tests/test_oracle.py); a
            # finite-difference Hessian-vector product over the first-order kernels was measured and is NOT offered: in float32 it
            # misses the exact meta-gradient by 20-100 % at every step size (profiles/r02_second_order_fd.txt)
            raise NotImplementedError("second-order MAML needs double-backward through the fused kernels "
                                      "(SURVEY.md 8f item 2); use track_higher_grads=False (FOMAML)")

    def _metatrain_step(self, items_b: Dict[str, Dict[str, tuple]]) -> dict:
        """One meta-step on one meta-batch {speaker: {"train": batch, "test": batch}} (maml.py:36-105)."""
        eng = self.engine
        speakers = list(items_b.keys())
        N = len(speakers)
        mine = self.shard.my_tasks(N)
        n_inner = self.params["n_inner_train"]
        losses, mcds = [], []
        if not mine:
            self.meta_grad.zero_()
        # every batch of this rank's tasks goes to the copy stream up front: H2D of the later tasks overlaps the earlier tasks' passes
        staged = self._stage([items_b[speakers[i]][k] for i in mine for k in ("train", "test")])
        train = {i: staged[2 * n] for n, i in enumerate(mine)}
        test = {i: staged[2 * n + 1] for n, i in enumerate(mine)}
        j = 0
        for group in self._group_plan(mine, train):
            if len(group) == 1:
                self._adapt(group[0], train[group[0]], n_inner)
                slots = [(self.fast, self.task_grad, self.task_bn)]
            else:       # first inner step of the whole group as one pass (the recurrences share their per-step hand-offs)
                self._adapt_group(group, train, n_inner)
                slots = [self._slot(k) for k in range(len(group))]
            tshape = {(tuple(test[i][1].shape), test[i][3].shape[2]) for i in group}
            if len(group) > 1 and len(tshape) == 1 and self.params.get("group_tasks", True):
                # the test-split passes of the group as ONE grouped pass with per-task weights: the recurrences run task by task (every
                # task has its own adapted weights), everything between them overlaps across the tasks; the task gradients land in the
                # slots' gradient buffers and are mixed with weights 1/N afterwards (maml.py:73-74, 94-98)
                bds = [self._unpack_batch(test[i])[0] for i in group]
                B, L = bds[0]["inputs"].shape
                T = bds[0]["melspecs"].shape[2]
                masks = [self._masks(i, n_inner, B, T, L, slot=k) for k, i in enumerate(group)]
                lg = eng.forward_group([s_[0] for s_ in slots], [s_[2] for s_ in slots], bds, masks)
                for k in range(len(group)):
                    mcds.append(eng.mcd_group(k, bds[k]["melspec_lengths"]))
                    losses.append(lg[k:k + 1])
                eng.backward_group([s_[0] for s_ in slots], [s_[1] for s_ in slots])
                for k in range(len(group)):
                    eng.axpy(self.meta_grad, slots[k][1], 1.0 / N, init=(j == 0))
                    j += 1
                continue
            for k, i in enumerate(group):
                fast, _, bn = slots[k]
                inputs, _ = self._unpack_batch(test[i])
                B, L = inputs["inputs"].shape
                T = inputs["melspecs"].shape[2]
                _, loss = eng.forward(fast, bn, inputs, self._masks(i, n_inner, B, T, L), outputs=False)
                mcds.append(eng.mcd(inputs["melspec_lengths"]))      # maml.py:78-82, on the device: no copy of the mels, no host sync
                # task_grads = autograd.grad(loss_test, fmodel.parameters(time=-1)); mix_grad weight 1/N (maml.py:73-74, 94-98)
                last = self.shard.world > 1 and j == len(mine) - 1
                if last:      # the allreduce of everything but the encoder gradients starts while this pass finishes (parallel.py)
                    if self._bwd_event is None:
                        self._bwd_event = torch.cuda.Event()
                        self._bwd_event.record()      # creates the underlying cudaEvent_t (torch creates it lazily)
                    eng.backward(fast, self.meta_grad, accumulate=(j > 0), scale=1.0 / N, decoder_done=self._bwd_event)
                    self.shard.allreduce_tail_async(self.meta_grad, eng.encoder_prefix(), self._bwd_event)
                else:
                    eng.backward(fast, self.meta_grad, accumulate=(j > 0), scale=1.0 / N)
                losses.append(loss)
                j += 1
        sumsq = self._outer_update()
        local = torch.cat(losses) if losses else torch.zeros(0, device=self.device)
        mcd = torch.cat(mcds) if mcds else torch.zeros(0, device=self.device)
        return {"loss_test": local, "mcd": mcd, "task_index": mine, "grad_sumsq": sumsq}
