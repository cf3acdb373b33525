"""Reptile driver: mirror of msa_tts/reptile.py on the CUDA hot path.

Per speaker: adapt ``n_inner_train`` SGD steps from theta_0, task "gradient" = -(theta_T - theta_0)
(reptile.py:42, 73-77).  Two outer-loop semantics are provided:
  * ``reptile_sequential=True`` -- the DEFAULT on one GPU: the reference's literal behaviour -- an outer step after EACH
    speaker, each starting from the already updated weights, ``step_global`` advancing once per speaker
    (reptile.py:37-39, 82-89).  Cannot shard; world size must be 1.
  * ``reptile_sequential=False`` (opt-in; BASELINE.json config 3; the default when the run is sharded over several GPUs,
    with a warning): all speakers of the meta-batch start from the same theta_0, the deltas are averaged with weights 1/N
    and joined by ONE allreduce -- the only form that shards.
With one speaker per meta-batch the two coincide.
"""
from __future__ import annotations

from typing import Dict

import torch

from .metatrainer import MetaTrainer


class Reptile(MetaTrainer):
    def __init__(self, **params):
        super().__init__(**params)
        seq = params.get("reptile_sequential", None)
        if seq is None:
            seq = self.shard.world == 1
            if not seq and self.shard.rank == 0:
                import warnings
                warnings.warn("Reptile on several GPUs runs the BATCHED variant (all speakers of a meta-batch start from the same "
                              "weights, one averaged outer step); the reference's per-speaker sequential outer steps cannot be "
                              "sharded.  Set reptile_sequential explicitly to silence this.")
        self.sequential = bool(seq)
        if self.sequential and self.shard.world > 1:
            raise ValueError("sequential Reptile (the reference's literal semantics) cannot be sharded")

    def _eval_test(self, i: int, task, n_inner: int, fast=None, bn=None) -> torch.Tensor:
        """reptile.py:58-70: test loss of the adapted weights, no gradient."""
        fast = self.fast if fast is None else fast
        bn = self.task_bn if bn is None else bn
        inputs, _ = self._unpack_batch(task["test"])
        B, L = inputs["inputs"].shape
        T = inputs["melspecs"].shape[2]
        _, loss = self.engine.forward(fast, bn, inputs, self._masks(i, n_inner, B, T, L), outputs=False)
        self._mcds.append(self.engine.mcd(inputs["melspec_lengths"]))       # reptile.py:62-66, on the device
        return loss

    def _metatrain_step(self, items_b: Dict[str, Dict[str, tuple]], eval_test: bool = True) -> dict:
        eng = self.engine
        speakers = list(items_b.keys())
        N = len(speakers)
        n_inner = self.params["n_inner_train"]
        losses = []
        self._mcds = []
        if self.sequential:
            for i, spk in enumerate(speakers):
                self._adapt(i, items_b[spk]["train"], n_inner)
                if eval_test:
                    losses.append(self._eval_test(i, items_b[spk], n_inner))
                eng.reptile_delta(self.meta_grad, self.fast, self.theta, 1.0, init=True)       # reptile.py:75-77
                sumsq = self._outer_update()                                                     # reptile.py:82-89
            return {"loss_test": torch.cat(losses) if losses else None, "mcd": torch.cat(self._mcds) if self._mcds else None,
                    "task_index": list(range(N)), "grad_sumsq": sumsq}
        mine = self.shard.my_tasks(N)
        if not mine:
            self.meta_grad.zero_()
        train = {i: items_b[speakers[i]]["train"] for i in mine}
        j = 0
        for group in self._group_plan(mine, train):
            if len(group) == 1:
                self._adapt(group[0], train[group[0]], n_inner)
                slots = [(self.fast, self.task_grad, self.task_bn)]
            else:       # the first inner step of the group as one pass from theta_0, the other steps task by task
                self._adapt_group(group, train, n_inner)
                slots = [self._slot(k) for k in range(len(group))]
            tests = [items_b[speakers[i]]["test"] for i in group]
            if eval_test and len(group) > 1 and len({(tuple(b[1].shape), b[3].shape[2]) for b in tests}) == 1:
                # reptile.py:58-70 for the whole group: one grouped forward with per-task weights (recurrences task by task, the rest
                # overlapped across the tasks)
                bds = [self._unpack_batch(b)[0] for b in tests]
                B, L = bds[0]["inputs"].shape
                T = bds[0]["melspecs"].shape[2]
                masks = [self._masks(i, n_inner, B, T, L, slot=k) for k, i in enumerate(group)]
                lg = eng.forward_group([s_[0] for s_ in slots], [s_[2] for s_ in slots], bds, masks)
                for k in range(len(group)):
                    losses.append(lg[k:k + 1])
                    self._mcds.append(eng.mcd_group(k, bds[k]["melspec_lengths"]))
                done_test = True
            else:
                done_test = False
            for k, i in enumerate(group):
                fast, _, bn = slots[k]
                if eval_test and not done_test:
                    losses.append(self._eval_test(i, items_b[speakers[i]], n_inner, fast, bn))
                eng.reptile_delta(self.meta_grad, fast, self.theta, 1.0 / N, init=(j == 0))
                j += 1
        sumsq = self._outer_update()
        return {"loss_test": torch.cat(losses) if losses else None, "mcd": torch.cat(self._mcds) if self._mcds else None,
                "task_index": mine, "grad_sumsq": sumsq}
