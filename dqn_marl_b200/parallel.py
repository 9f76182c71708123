"""Data-parallel plumbing (one process per GPU, torch.distributed).

The path shards naturally: envs and replay are rank-local, every rank holds a full replica of the
Q-network, and the ONLY exchange step of a learn call is the all-reduce of the flat 8,157,093-float gradient
(NCCL over NVLink on GPUs; gloo in the CPU tests of this host-side logic)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def env_shard(rank: int, world: int, n_envs_total: int) -> Tuple[int, int]:
    """Contiguous partition of the global env ids: (first id, count) of this rank.  The first id is the rank's
    ``env_id_base`` (Philox counter word 0), so a sharded run draws exactly what a single-GPU run of the same global
    batch would."""
    base, rem = divmod(n_envs_total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum-all-reduce then divide: the gradient of the global batch mean when every rank used a batch of equal
    size (dqn_agent.py:151 is a mean over the batch)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
