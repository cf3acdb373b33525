"""JointTrainer: host-side mirror of msa_tts/baseline.py for the hot path (BASELINE.json configs[0]: the plain supervised
forward + loss + backward + optimizer step of ``Tacotron2NV``).

Keeps the reference's call surface for what is in scope -- model / criterion / optimizer set-up from the ``params`` dictionary
(baseline.py:24-105), ``_unpack_batch`` (107-129), ``run`` / ``_train`` / ``_test`` (181-296), checkpoints with the reference's
``state_dict`` keys (131-134, 151-158) -- on top of the drop-in ``Tacotron2NV`` (model.py) and the fused training step of
``continual.train_step``.  The higher-based ``_metatest`` of the reference class (baseline.py:299-361) is the same computation as
``MetaTrainer._metatest`` (metatrainer.py here); a joint model is meta-tested by loading its checkpoint into ``MAML`` with
``finetune=True``.

Out of scope: the reference's data loaders / audio front end (pass any iterable of batch tuples, e.g. from
``data.Collator``), TensorBoard, plots, the YAML dump of the parameters.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from .continual import train_step
from .engine import batch_to_device
from .model import Tacotron2NV


class JointTrainer:
    def __init__(self, **params):
        self.params = params
        crit = params.get("criterion", {"criterion_type": "Tacotron2Loss", "reduction": "none", "pos_weight": 10.0})
        if crit["criterion_type"] != "Tacotron2Loss":
            raise RuntimeError(f"Criterion {crit} not defined.")                     # baseline.py:97-98
        mp = dict(params["model"])
        for k in ("freeze_charemb", "freeze_encoder", "freeze_decoder"):             # baseline.py:52-55
            if k in params:
                mp[k] = params[k]
        if params.get("model_name", "Tacotron2NV") != "Tacotron2NV":
            raise NotImplementedError                                                # baseline.py:59-62
        self.model = Tacotron2NV(mp, device=params.get("device", None), criterion=crit, gemm_tf32=params.get("gemm_tf32", 0),
                                 init_seed=params.get("init_seed", 0))
        self.device = self.model.engine.device
        self.model.to(self.device)                                                   # baseline.py:64 (a no-op here)
        self.speaker_emb_type = mp["speaker_emb_type"]
        self.optim = params["optim"]            # evaluated like helpers.get_optimizer by train_step (SGD / Adam rules of torch.optim)
        self.step_global = 0
        self.best_test_loss = 100000000.0
        if params.get("finetune", False):
            self._load_checkpoint()

    # ---- data -----------------------------------------------------------------------------------------------------------
    def _unpack_batch(self, batch_items):
        """baseline.py:107-129: the model's keyword inputs on the device + the stop labels."""
        d = batch_to_device(batch_items, self.device, self.speaker_emb_type, non_blocking=True)
        return d, d["stop"]

    # ---- epochs ---------------------------------------------------------------------------------------------------------
    def _train(self, epoch: int, dataloader_train=None) -> List[dict]:
        """baseline.py:195-237: per batch forward, loss, backward, optimizer step (the reference clips the STALE gradients of the
        previous step right before ``zero_grad()``, which changes nothing -- SURVEY.md Q12 -- so there is no clip here either).
        Returns one log dict per step: loss and MCD as device tensors (no host sync inside the loop)."""
        dl = dataloader_train if dataloader_train is not None else getattr(self, "dataloader_train", None)
        if dl is None:
            raise RuntimeError("JointTrainer._train: set self.dataloader_train (an iterable of batch tuples)")
        self.model.train()
        logs = []
        for batch in dl:
            log = train_step(self.model, batch, self.optim)
            self.model.count_bn_batches()
            log["step"] = self.step_global
            logs.append(log)
            self.step_global += 1
        self.model.engine.check_abort()
        return logs

    def _test(self, epoch: int, dataloader_test=None) -> dict:
        """baseline.py:254-296: the test loader in TRAIN mode (dropout on, BatchNorm batch statistics, running statistics moving --
        the reference calls ``self.model.train()`` there) without gradients; mean loss and mean MCD over the batches, and
        ``checkpoint_best.pt`` whenever the mean loss improves."""
        dl = dataloader_test if dataloader_test is not None else getattr(self, "dataloader_test", None)
        if dl is None:
            raise RuntimeError("JointTrainer._test: set self.dataloader_test (an iterable of batch tuples)")
        self.model.train()
        eng, m = self.model.engine, self.model
        losses, mcds = [], []
        for batch in dl:
            bd, _ = self._unpack_batch(batch)
            B, L = bd["inputs"].shape
            T = bd["melspecs"].shape[2]
            _, loss = eng.forward(m.flat, m.bn_flat, bd, m._masks(B, T, L), outputs=False)
            mcds.append(eng.mcd(bd["melspec_lengths"]))
            losses.append(loss)
            m.count_bn_batches()
        if not losses:
            return {"loss": None, "mcd": None, "best": False}
        loss_total = float(torch.cat(losses).mean())              # the one host sync of the epoch
        mcd_total = float(torch.cat(mcds).mean())
        eng.check_abort()
        best = loss_total < self.best_test_loss
        if best:
            self.best_test_loss = loss_total
            self._save_checkpoint(os.path.join(self.params.get("output_path", "."), "checkpoint_best.pt"))
        return {"loss": loss_total, "mcd": mcd_total, "best": best, "step": self.step_global}

    def run(self, dataloader_train=None, dataloader_test=None, n_epochs: Optional[int] = None) -> List[dict]:
        """baseline.py:181-193: per epoch ``_train``, ``_test``, a checkpoint every ``ckpt_save_epoch_interval`` epochs.  Returns the
        per-epoch test logs (``self.train_logs`` keeps the per-step training logs of the last epoch)."""
        if dataloader_train is not None:
            self.dataloader_train = dataloader_train
        if dataloader_test is not None:
            self.dataloader_test = dataloader_test
        n_epochs = int(self.params.get("n_epochs", 1)) if n_epochs is None else n_epochs
        self.step_global = 0
        self.best_test_loss = 100000000.0
        out = []
        for epoch in range(1, n_epochs + 1):
            self.train_logs = self._train(epoch)
            if getattr(self, "dataloader_test", None) is not None:
                out.append(self._test(epoch))
            k = self.params.get("ckpt_save_epoch_interval", 0)
            if k and epoch % k == 0:
                self._save_checkpoint()
        return out

    # ---- checkpoints (baseline.py:131-134, 151-158) ---------------------------------------------------------------------
    def _save_checkpoint(self, path: Optional[str] = None) -> str:
        k = self.step_global // 100
        path = path or os.path.join(self.params.get("output_path", "."), f"checkpoint_{k}.pt")
        torch.save({k_: v.detach().cpu().clone() for k_, v in self.model.state_dict().items()}, path)
        return path

    def _load_checkpoint(self) -> None:
        """baseline.py:151-158: parameters only; a tensor that is missing or has another shape is reported and skipped."""
        print(f"Loading checkpoint from  {self.params['finetune_checkpoint_path']}")
        ckpt = torch.load(self.params["finetune_checkpoint_path"], map_location="cpu")
        own = dict(self.model.named_parameters())
        with torch.no_grad():
            for name, p in own.items():
                if name in ckpt and tuple(ckpt[name].shape) == tuple(p.shape):
                    p.copy_(ckpt[name])
                else:
                    print(f"Could not load weights for {name}")
