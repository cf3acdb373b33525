"""MetaTrainer: host-side mirror of msa_tts/metatrainer.py for the hot path.

Keeps the reference's call surface for the part that is in scope -- model / criterion /
inner + outer optimizer set-up from the ``params`` dict (metatrainer.py:33-56, 81-93),
``_unpack_batch`` (95-117), checkpoint save / load with the reference's ``state_dict`` keys
(119-122, 138-146) -- and owns what the B200 design adds: the flat parameter / gradient /
optimizer-state buffers, the per-task fast weights and private BatchNorm statistics
(SURVEY.md Q18), the dropout-mask stream keyed by (meta-step, task, pass) so that sharded and
unsharded runs agree, and the task-to-rank sharding with ONE allreduce.

Out of scope here: data loaders, audio, TensorBoard, plotting.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from .engine import Engine, batch_to_device
from .helpers import optimizer_hparams
from .parallel import ShardInfo


def plan_task_groups(mine, shapes, n_inner: int, stateful_inner: bool, enabled: bool, group_size) -> List[List[int]]:
    """Which tasks of a meta-batch share grouped passes (include/msa_b200.h "task groups").  Every task of a meta-batch starts from
    the same theta (maml.py:38-41), so the train-split passes of tasks with equal (B, T, L) run as one pass whose recurrences hand
    their data over once per step for all rows; later inner steps and the test passes of the same group run as grouped passes with
    per-task weights.  ``mine``: task indices of this rank; ``shapes[i]``: ((B, L), T) of task i's train batch; ``group_size(n, B)``:
    most tasks one grouped pass may carry.  A stateful inner optimizer (momentum / Adam) with more than one inner step keeps its
    state per task and takes the plain path.  Returns lists of task indices in task order; singletons take the plain path."""
    if n_inner < 1 or not enabled or (stateful_inner and n_inner > 1):
        return [[i] for i in mine]
    by_shape: Dict[tuple, List[int]] = {}
    for i in mine:
        by_shape.setdefault(shapes[i], []).append(i)
    plan = []
    for shape, idx in by_shape.items():
        gmax = max(1, int(group_size(len(idx), shape[0][0])))
        for k in range(0, len(idx), gmax):
            plan.append(idx[k:k + gmax])
    plan.sort(key=lambda g: g[0])
    return plan


class _Staged:
    """A batch tuple whose device copy is in flight on the copy stream; indexable like the tuple (shapes for the group plan)."""

    def __init__(self, host, dev, event):
        self.host, self.dev, self.event = host, dev, event

    def __getitem__(self, i):
        return self.host[i]


class MetaTrainer:
    def __init__(self, **params):
        self.params = params
        self.model_params = params["model"]
        crit = params.get("criterion", {"criterion_type": "Tacotron2Loss", "reduction": "none", "pos_weight": 10.0})
        if crit["criterion_type"] != "Tacotron2Loss":
            raise RuntimeError(f"Criterion {crit} not defined.")            # metatrainer.py:88-89
        self.shard = ShardInfo.from_env(params.get("distributed", None))
        self.device = torch.device(f"cuda:{self.shard.local_rank}")
        torch.cuda.set_device(self.device)
        self.speaker_emb_type = self.model_params["speaker_emb_type"]
        self.engine = Engine(self.model_params, self.device, reduction=crit["reduction"], pos_weight=crit["pos_weight"],
                             gemm_tf32=params.get("gemm_tf32", 0))
        self.layout = self.engine.layout
        # inner / outer optimizers: name + eval'ed string hyper-parameters, like helpers.get_optimizer (helpers.py:20-26)
        self.inner = optimizer_hparams(params["optim_inner"])
        self.outer = optimizer_hparams(params["optim_outer"])
        if self.inner["name"] not in ("SGD", "Adam"):
            raise NotImplementedError("inner optimizer: the SGD and Adam rules are implemented (SURVEY.md 8f item 3)")
        if self.inner["name"] == "Adam" and (self.inner.get("amsgrad", False) or self.inner.get("maximize", False)):
            raise NotImplementedError("inner Adam: amsgrad / maximize are not implemented")
        if self.outer["name"] not in ("SGD", "Adam"):
            raise NotImplementedError("outer optimizer: SGD and Adam are implemented")
        # flat state
        from .synth import init_params
        self.theta = self.engine.flat_from_dict(init_params(self.model_params, params.get("init_seed", 0)))
        self.base_bn = self.engine.new_bn_stats()      # never updated by MAML/Reptile
        self.meta_grad = self.engine.new_flat()
        self.fast = self.engine.new_flat()
        self.task_grad = self.engine.new_flat()
        self.task_bn = self.engine.new_bn_stats()
        self.inner_buf = self.engine.new_flat() if (self.inner["name"] == "SGD" and self.inner.get("momentum", 0.0)) else None
        # inner Adam: moment buffers, reset for every task (higher builds the optimizer state inside innerloop_ctx)
        self.inner_m = self.engine.new_flat() if self.inner["name"] == "Adam" else None
        self.inner_v = self.engine.new_flat() if self.inner["name"] == "Adam" else None
        self.outer_m = self.engine.new_flat() if (self.outer["name"] == "Adam" or self.outer.get("momentum", 0.0)) else None
        self.outer_v = self.engine.new_flat() if self.outer["name"] == "Adam" else None
        self.sumsq = torch.zeros(1, device=self.device)
        self.step_global = 0
        self._outer_steps = 0            # outer optimizer steps taken so far: survives run()'s reset of step_global (maml.py:20), like
        #                                  the state of the reference's torch optimizer does
        self.mask_seed = int(params.get("dataset_random_seed", 1234))
        self._mask_bufs: Dict[tuple, torch.Tensor] = {}
        self._slots: list = []
        self._copy_stream = None
        self.injected_masks = None       # parity tests: {(task_index, pass_index): reference-layout mask dict}
        if params.get("finetune", False):
            self._load_checkpoint()

    # ---- data ------------------------------------------------------------------------------------
    def _unpack_batch(self, batch_items):
        """metatrainer.py:95-117.  A batch that ``_stage`` has already put on the copy stream is only waited for."""
        if isinstance(batch_items, _Staged):
            torch.cuda.current_stream().wait_event(batch_items.event)
            return batch_items.dev, batch_items.dev["stop"]
        d = batch_to_device(batch_items, self.device, self.speaker_emb_type, non_blocking=True)
        stop = d["stop"]
        return d, stop

    def _stage(self, batches: List[tuple]) -> List["_Staged"]:
        """Host -> device copies of the coming batches on a SIDE stream: with pinned host tensors
        (``MetaCollator(pin_memory=True)``) the copy engine moves the batches of the later tasks while the compute stream works on the
        earlier ones; every consumer waits for its batch's own event.  Batches that already live on the device pass through."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream()
        out = []
        with torch.cuda.stream(self._copy_stream):
            for b in batches:
                if isinstance(b, _Staged):
                    out.append(b)
                    continue
                d = batch_to_device(b, self.device, self.speaker_emb_type, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                for t in d.values():
                    t.record_stream(cur)
                out.append(_Staged(b, d, ev))
        return out

    def _masks(self, task_index: int, pass_index: int, B: int, T: int, L: int, slot: int = 0) -> torch.Tensor:
        """Dropout keep-masks for one pass, keyed by (meta-step, task, pass) -- independent of sharding and of task grouping
        (``slot`` only selects the buffer: the passes of a group need their masks alive at the same time)."""
        if self.injected_masks is not None:
            return self.engine.pack_masks(self.injected_masks[(task_index, pass_index)], B, T, L)
        key = (B, T, L, slot)
        if key not in self._mask_bufs:
            self._mask_bufs[key] = torch.empty(self.engine.mask_bytes(B, T, L), dtype=torch.uint8, device=self.device)
        seed = (self.mask_seed * 1000003 + self.step_global) * 1000003 + task_index * 64 + pass_index
        return self.engine.generate_masks(B, T, L, seed, self._mask_bufs[key])

    # ---- inner loop (higher.innerloop_ctx + diffopt.step, maml.py:40-54) ------------------------------
    def _adapt(self, task_index: int, batch, n_inner: int, pass0: int = 0):
        """n_inner x (forward, backward, functional SGD / Adam step) on the train split, starting from theta.  The first step reads
        theta itself and writes the updated weights to ``fast`` (the functional update is out of place), so the 121 MB
        ``fast <- theta`` copy of a literal ``innerloop_ctx`` never happens.  ``pass0``: first pass index of the dropout-mask stream."""
        eng = self.engine
        self.task_bn.copy_(self.base_bn)
        inputs, _ = self._unpack_batch(batch)
        B, L = inputs["inputs"].shape
        T = inputs["melspecs"].shape[2]
        h = self.inner
        losses = []
        if h["name"] == "Adam":
            self.inner_m.zero_()
            self.inner_v.zero_()
        if n_inner == 0:
            self.fast.copy_(self.theta)
        for it in range(n_inner):
            src = self.theta if it == 0 else self.fast
            _, loss = eng.forward(src, self.task_bn, inputs, self._masks(task_index, pass0 + it, B, T, L), outputs=False)
            eng.backward(src, self.task_grad)
            if h["name"] == "Adam":
                eng.adam_step(src, self.task_grad, self.fast, self.inner_m, self.inner_v, lr=h["lr"], step=it + 1,
                              betas=h.get("betas", (0.9, 0.999)), eps=h.get("eps", 1e-8), weight_decay=h.get("weight_decay", 0.0))
            else:
                eng.sgd_step(src, self.task_grad, p_out=self.fast, lr=h["lr"], momentum=h.get("momentum", 0.0),
                             dampening=h.get("dampening", 0.0), weight_decay=h.get("weight_decay", 0.0),
                             nesterov=h.get("nesterov", False), buf=self.inner_buf, first_step=(it == 0))
            losses.append(loss)
        return losses

    # ---- grouped first inner step ------------------------------------------------------------------------------
    def _group_plan(self, mine: List[int], batches: Dict[int, tuple], n_inner: Optional[int] = None) -> List[List[int]]:
        """Tasks whose inner steps can share grouped passes (``plan_task_groups``)."""
        h = self.inner
        stateful = (h["name"] == "Adam") or bool(h.get("momentum", 0.0))
        shapes = {i: (tuple(batches[i][1].shape), batches[i][3].shape[2]) for i in mine}
        n_inner = self.params["n_inner_train"] if n_inner is None else n_inner
        return plan_task_groups(mine, shapes, n_inner, stateful, self.params.get("group_tasks", True), self.engine.group_size)

    def _slot(self, k: int):
        """Per-slot fast weights / gradient / BatchNorm buffers of a group (slot 0 = the buffers of the plain path)."""
        while len(self._slots) <= k:
            if not self._slots:
                self._slots.append((self.fast, self.task_grad, self.task_bn))
            else:
                self._slots.append((self.engine.new_flat(), self.engine.new_flat(), self.engine.new_bn_stats()))
        return self._slots[k]

    def _inner_step(self, src, grad, dst, it: int) -> None:
        eng, h = self.engine, self.inner
        if h["name"] == "Adam":
            eng.adam_step(src, grad, dst, self.inner_m, self.inner_v, lr=h["lr"], step=it + 1,
                          betas=h.get("betas", (0.9, 0.999)), eps=h.get("eps", 1e-8), weight_decay=h.get("weight_decay", 0.0))
        else:
            eng.sgd_step(src, grad, p_out=dst, lr=h["lr"], momentum=h.get("momentum", 0.0),
                         dampening=h.get("dampening", 0.0), weight_decay=h.get("weight_decay", 0.0),
                         nesterov=h.get("nesterov", False), buf=self.inner_buf, first_step=(it == 0))

    def _adapt_group(self, group: List[int], batches: Dict[int, tuple], n_inner: int, pass0: int = 0):
        """First inner step of all tasks of ``group`` as ONE grouped pass from theta, then the remaining inner steps task by task.
        Afterwards slot k holds the adapted weights and the private BatchNorm statistics of task group[k]."""
        eng = self.engine
        slots = [self._slot(k) for k in range(len(group))]
        bds = []
        for k, i in enumerate(group):
            slots[k][2].copy_(self.base_bn)
            bds.append(self._unpack_batch(batches[i])[0])
        B, L = bds[0]["inputs"].shape
        T = bds[0]["melspecs"].shape[2]
        masks = [self._masks(i, pass0, B, T, L, slot=k) for k, i in enumerate(group)]
        losses = eng.forward_group(self.theta, [s[2] for s in slots], bds, masks)
        eng.backward_group(self.theta, [s[1] for s in slots])
        for k in range(len(group)):
            if self.inner["name"] == "Adam":
                self.inner_m.zero_()
                self.inner_v.zero_()
            self._inner_step(self.theta, slots[k][1], slots[k][0], 0)
        out = [[losses[k:k + 1]] for k in range(len(group))]
        # the later inner steps have per-task weights: still ONE grouped pass per step (recurrences task by task, everything between
        # them overlapped across the tasks), then the per-task functional update
        for it in range(1, n_inner):
            masks = [self._masks(i, pass0 + it, B, T, L, slot=k) for k, i in enumerate(group)]
            lg = eng.forward_group([s_[0] for s_ in slots], [s_[2] for s_ in slots], bds, masks)
            eng.backward_group([s_[0] for s_ in slots], [s_[1] for s_ in slots])
            for k in range(len(group)):
                fast, grad, _ = slots[k]
                self._inner_step(fast, grad, fast, it)
                out[k].append(lg[k:k + 1])
        return out

    # ---- epoch loops (maml.py:19-36, reptile.py:19-36) -------------------------------------------------------------------
    def _metatrain(self, epoch: int, dataloader_metatrain=None) -> List[dict]:
        """maml.py:33-108 / reptile.py:33-105: ``_metatrain(epoch)`` iterates ``self.dataloader_metatrain`` like the reference; an
        iterable of meta-batches may be passed instead (the reference's DataLoader construction is out of scope)."""
        dl = dataloader_metatrain if dataloader_metatrain is not None else getattr(self, "dataloader_metatrain", None)
        if dl is None:
            raise RuntimeError("_metatrain: set self.dataloader_metatrain (an iterable of {speaker: {train, test}} dicts)")
        logs = [self._metatrain_step(items_b) for items_b in dl]
        self.engine.abort_flush()
        return logs

    def run(self, dataloader_metatrain=None, n_epochs: Optional[int] = None, dataloader_metatest=None) -> List[dict]:
        """maml.py:19-31 / reptile.py:19-31: per epoch one pass over the meta-train loader, a checkpoint every
        ``ckpt_save_epoch_interval`` epochs and a meta-test every ``metatest_epoch_interval`` epochs (both optional here: a params
        dict without the key, or without a meta-test loader, skips that part).  Returns the meta-train logs; the meta-test logs of
        the last meta-test are kept in ``self.last_metatest``."""
        if dataloader_metatrain is not None:
            self.dataloader_metatrain = dataloader_metatrain
        if dataloader_metatest is not None:
            self.dataloader_metatest = dataloader_metatest
        n_epochs = int(self.params.get("n_epochs", 1)) if n_epochs is None else n_epochs
        self.step_global = 0
        out = []
        for epoch in range(1, n_epochs + 1):
            out += self._metatrain(epoch)
            k = self.params.get("ckpt_save_epoch_interval", 0)
            if k and epoch % k == 0:
                self._save_checkpoint()
            k = self.params.get("metatest_epoch_interval", 0)
            if k and epoch % k == 0 and getattr(self, "dataloader_metatest", None) is not None:
                self.last_metatest = self._metatest(epoch)
        return out

    # ---- meta-test (maml.py:115-179, reptile.py:108-172, baseline.py:299-361) --------------------------------------
    METATEST_PASS0 = 32      # dropout-mask stream of the meta-test passes (pass indices 32 ..; meta-training uses 0 .. n_inner_train)

    def _metatest_step(self, items_b: Dict[str, Dict[str, tuple]], return_outputs: bool = False) -> dict:
        """Meta-test on one meta-batch {speaker: {"train": batch, "test": batch}}: per speaker ``n_inner_test`` inner steps on the
        "train" split starting from theta (the model stays in train mode, maml.py:116), then the loss and the MCD of the ADAPTED
        weights on the "test" split, no gradient (maml.py:139-149, 167-169).  Nothing persistent moves: theta, the base BatchNorm
        statistics (higher's functional copy owns its buffers), the outer optimizer state and ``step_global`` stay as they were.
        Speakers are sharded over the ranks like in meta-training; there is no collective (every rank logs its own speakers).
        ``return_outputs``: also return every speaker's test-pass outputs ``[out_post, out_inner, out_stop, out_attn]`` (what the
        reference plots, maml.py:151-165); those passes then run speaker by speaker."""
        eng = self.engine
        speakers = list(items_b.keys())
        mine = self.shard.my_tasks(len(speakers))
        n_inner = int(self.params.get("n_inner_test", self.params["n_inner_train"]))
        p0 = self.METATEST_PASS0
        losses, mcds, outs = [], [], []
        train = {i: items_b[speakers[i]]["train"] for i in mine}
        plan = [[i] for i in mine] if return_outputs else self._group_plan(mine, train, n_inner)
        for group in plan:
            if len(group) == 1:
                self._adapt(group[0], train[group[0]], n_inner, pass0=p0)
                slots = [(self.fast, self.task_grad, self.task_bn)]
            else:
                self._adapt_group(group, train, n_inner, pass0=p0)
                slots = [self._slot(k) for k in range(len(group))]
            tests = [items_b[speakers[i]]["test"] for i in group]
            if len(group) > 1 and len({(tuple(b[1].shape), b[3].shape[2]) for b in tests}) == 1:
                bds = [self._unpack_batch(b)[0] for b in tests]
                B, L = bds[0]["inputs"].shape
                T = bds[0]["melspecs"].shape[2]
                masks = [self._masks(i, p0 + n_inner, B, T, L, slot=k) for k, i in enumerate(group)]
                lg = eng.forward_group([s_[0] for s_ in slots], [s_[2] for s_ in slots], bds, masks)
                for k in range(len(group)):
                    losses.append(lg[k:k + 1])
                    mcds.append(eng.mcd_group(k, bds[k]["melspec_lengths"]))
                continue
            for k, i in enumerate(group):
                fast, _, bn = slots[k]
                inputs, _ = self._unpack_batch(tests[k])
                B, L = inputs["inputs"].shape
                T = inputs["melspecs"].shape[2]
                out, loss = eng.forward(fast, bn, inputs, self._masks(i, p0 + n_inner, B, T, L), outputs=return_outputs)
                losses.append(loss)
                mcds.append(eng.mcd(inputs["melspec_lengths"]))
                if return_outputs:
                    outs.append(out)
        eng.abort_poll()
        log = {"loss_test": torch.cat(losses) if losses else torch.zeros(0, device=self.device),
               "mcd": torch.cat(mcds) if mcds else torch.zeros(0, device=self.device),
               "task_index": mine, "speakers": [speakers[i] for i in mine]}
        if return_outputs:
            log["outputs"] = outs
        return log

    def _metatest(self, epoch: int, dataloader_metatest=None) -> List[dict]:
        """maml.py:115-179: ``_metatest(epoch)`` over ``self.dataloader_metatest`` (or the iterable passed in); one log dict per
        meta-batch with the per-speaker test losses and MCDs as device tensors (plots and TensorBoard are out of scope)."""
        dl = dataloader_metatest if dataloader_metatest is not None else getattr(self, "dataloader_metatest", None)
        if dl is None:
            raise RuntimeError("_metatest: set self.dataloader_metatest (an iterable of {speaker: {train, test}} dicts)")
        logs = [self._metatest_step(items_b) for items_b in dl]
        self.engine.abort_flush()
        return logs

    # ---- outer update (maml.py:94-105 / reptile.py:82-89) ----------------------------------------------
    def _outer_update(self) -> torch.Tensor:
        """allreduce(sum) of the locally weighted meta-gradient, grad-norm, clip, optimizer step -- all on flat buffers."""
        eng = self.engine
        self.shard.allreduce_sum(self.meta_grad, eng.encoder_prefix())       # two parts; the first may already be in flight (parallel.py)
        eng.sumsq(self.meta_grad, self.sumsq)                     # apply_grad's norm / clip_grad_norm_'s total norm
        # a persistent kernel that gave up polling leaves garbage gradients: poison the norm on the device (the update kernels below
        # skip on a non-finite norm, theta and the optimizer state stay intact) and let the host find out without a stall
        eng.abort_guard(self.sumsq)
        thr = float(self.params["grad_clip_thresh"]) if self.params.get("clip_grad_norm", False) else 0.0
        o = self.outer
        if o["name"] == "SGD":
            eng.clip_sgd(self.theta, self.meta_grad, self.sumsq, lr=o["lr"], max_norm=thr, momentum=o.get("momentum", 0.0),
                         dampening=o.get("dampening", 0.0), weight_decay=o.get("weight_decay", 0.0), nesterov=o.get("nesterov", False),
                         buf=self.outer_m, first_step=(self._outer_steps == 0))
        else:
            eng.clip_adam(self.theta, self.meta_grad, self.outer_m, self.outer_v, self.sumsq, lr=o["lr"], step=self._outer_steps + 1,
                          betas=o.get("betas", (0.9, 0.999)), eps=o.get("eps", 1e-8), weight_decay=o.get("weight_decay", 0.0),
                          max_norm=thr)
        self.step_global += 1
        self._outer_steps += 1
        eng.abort_poll()
        return self.sumsq

    # ---- checkpoints (metatrainer.py:119-122, 138-146) --------------------------------------------------
    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {k: v.detach().cpu().clone() for k, v in self.engine.dict_from_flat(self.theta).items()}
        for k, v in self.engine.bn_dict(self.base_bn).items():
            sd[k] = v.detach().cpu().clone()
        for name in self.layout.bn_names:
            sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        P = {n: sd[n] for n in self.layout.names()}
        self.theta.copy_(self.engine.flat_from_dict(P))

    def _save_checkpoint(self, path: Optional[str] = None) -> str:
        k = self.step_global // 100
        path = path or os.path.join(self.params.get("output_path", "."), f"checkpoint_{k}.pt")
        if self.shard.rank == 0:
            torch.save(self.state_dict(), path)
        return path

    def _load_checkpoint(self) -> None:
        """metatrainer.py:138-146: parameters only (the reference copies ``named_parameters()``, so the BatchNorm running statistics of
        the checkpoint are NOT loaded); a tensor that is missing or has another shape is reported and keeps its initial value."""
        print(f"Loading checkpoint from  {self.params['finetune_checkpoint_path']}")
        sd = torch.load(self.params["finetune_checkpoint_path"], map_location="cpu")
        P = {n: v.detach().cpu() for n, v in self.engine.dict_from_flat(self.theta).items()}
        for n in self.layout.names():
            if n in sd and tuple(sd[n].shape) == tuple(P[n].shape):
                P[n] = sd[n].detach().float().cpu()
            else:
                print(f"Could not load weights for {n}")
        self.theta.copy_(self.engine.flat_from_dict(P))
