"""Task sharding across GPUs: one process per GPU, speaker tasks split by index, ONE allreduce per meta-step.

Tasks of a meta-batch are independent given theta_0 (maml.py:38-41 vs 94-105): task i runs on rank i % W, each
rank accumulates sum_local w_i g_i into its flat fp32 meta-gradient buffer, and a single NCCL allreduce(SUM) of
that buffer (121 MB at default dims) over NVLink / NVSwitch yields the averaged meta-gradient on every rank; the
clip + outer-optimizer update is then replicated (identical on every rank, no broadcast).  The reference has no
distributed code at all -- this is the B200 design, not a port.

The allreduce is issued in TWO parts, always in the same order on every rank: the tail of the flat buffer (everything but the
encoder: ~90 % of the bytes) and then the encoder prefix.  A backward pass produces the encoder gradients last, so the rank whose
last task runs as a plain pass starts the tail part on a side stream as soon as the pass reports those gradients final
(``msa_backward_mark_event``) and it overlaps the encoder BiLSTM recurrence + encoder convolutions of that pass.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass
class ShardInfo:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    group: Optional[object] = None
    _side: Optional[object] = None          # side stream of the early (tail) part of the allreduce
    _tail_done: bool = False                # the tail part of this meta-step's allreduce is already in flight

    @staticmethod
    def from_env(spec=None) -> "ShardInfo":
        if isinstance(spec, ShardInfo):
            return spec
        if dist.is_available() and dist.is_initialized():
            return ShardInfo(dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", 0)))
        return ShardInfo()

    def my_tasks(self, n_tasks: int) -> List[int]:
        """Indices of the tasks of a meta-batch this rank adapts (task i -> rank i % W)."""
        return [i for i in range(n_tasks) if i % self.world == self.rank]

    def allreduce_tail_async(self, flat: torch.Tensor, prefix: int, ready: "torch.cuda.Event") -> None:
        """Start the allreduce of ``flat[prefix:]`` on a side stream once ``ready`` fires (the backward pass that is still running on
        the current stream has finished writing that part); ``allreduce_sum`` completes the job."""
        if self.world == 1:
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=flat.device)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            dist.all_reduce(flat[prefix:], op=dist.ReduceOp.SUM, group=self.group)
        self._tail_done = True

    def allreduce_sum(self, flat: torch.Tensor, prefix: int = 0) -> None:
        """allreduce(SUM) of a flat buffer.  With ``prefix`` > 0 as two collectives -- ``flat[prefix:]`` (skipped if
        ``allreduce_tail_async`` already issued it), then ``flat[:prefix]`` -- so that every rank issues the same sequence whether or
        not it could start the first part early."""
        if self.world == 1:
            return
        if prefix <= 0 or prefix >= flat.numel():
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return
        if not self._tail_done:
            dist.all_reduce(flat[prefix:], op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(flat[:prefix], op=dist.ReduceOp.SUM, group=self.group)
        if self._tail_done:
            torch.cuda.current_stream().wait_stream(self._side)
            self._tail_done = False

    def allreduce_max(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)

    def gather_scalars(self, local: torch.Tensor, n_tasks: int) -> torch.Tensor:
        """Per-task scalars (losses) from every rank, in task order; a tiny collective used only for logging."""
        if self.world == 1:
            return local
        full = torch.zeros(n_tasks, dtype=local.dtype, device=local.device)
        idx = self.my_tasks(n_tasks)
        if idx:
            full[torch.tensor(idx, device=local.device)] = local
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
        return full

    def barrier(self) -> None:
        if self.world > 1:
            dist.barrier(group=self.group)
