"""Task sharding across GPUs: one process per GPU, speaker tasks split by index, ONE allreduce per meta-step.

Tasks of a meta-batch are independent given theta_0 (maml.py:38-41 vs 94-105): task i runs on rank i % W, each
rank accumulates sum_local w_i g_i into its flat fp32 meta-gradient buffer, and a single NCCL allreduce(SUM) of
that buffer (121 MB at default dims) over NVLink / NVSwitch yields the averaged meta-gradient on every rank; the
clip + outer-optimizer update is then replicated (identical on every rank, no broadcast).  The reference has no
distributed code at all -- this is the B200 design, not a port.
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

    def allreduce_sum(self, flat: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

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
