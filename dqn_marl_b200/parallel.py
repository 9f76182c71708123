"""Data-parallel plumbing (one process per GPU, torch.distributed).

The path shards naturally: envs and replay are rank-local, every rank holds a full replica of the
Q-network, and the ONLY exchange step of a learn call is the all-reduce of the flat 8,157,093-float gradient
(NCCL over NVLink on GPUs; gloo in the CPU tests of this host-side logic, tests/test_parallel_cpu.py).  Everything here
is device-agnostic and is what the product calls: `VecDQNAgent.grad_step` -> `allreduce_overlapped`, `VecDQNAgent.__init__`
-> `min_over_ranks`, `VecDQNAgent.ready_to_learn` -> `learn_gate_open`, bench.py -> `max_over_ranks`."""
from __future__ import annotations

import os
from typing import Callable, Tuple

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


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def max_over_ranks(value: float, device, group=None) -> float:
    """Timings of a multi-GPU run are the MAX over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if _active(group):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def min_over_ranks(value: int, device, group=None) -> int:
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    if _active(group):
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def learn_gate_open(pushes: int, min_push: int, capacity: int, batch_size: int, steps: int, warmup_steps: int) -> bool:
    """`len(memory) > batch_size` (train_dqn.py:117-118) and the warm-up test (dqn_agent.py:128) on quantities EVERY rank agrees
    on: the number of remember_batch() calls, the smallest shard's transitions per call (`min_over_ranks` at construction) and
    the replicated learn-step counter.  A rank-local len(memory) opens one step apart on ranks whose env shards differ by one
    env (n_envs % world != 0); the learn step is a collective, so the run would hang in NCCL."""
    filled = min(pushes * min_push, capacity)
    return filled > batch_size and steps >= warmup_steps


def allreduce_overlapped(flat_g: torch.Tensor, head: int, part1: Callable[[], object], part2: Callable[[], object], group=None):
    """The one exchange step of a data-parallel learn call, overlapped with the backward.

    `part1()` must leave flat_g[head:] final (loss + backward of fc3 / fc2 / fc1: 99 % of the floats), `part2()` then fills
    flat_g[:head] (the convolution layers).  The sum-all-reduce of the tail is issued right after part 1 and runs on the
    communicator's stream while part 2 computes; the small head follows.  Returns part1()'s result; on return flat_g holds the
    SUM over ranks (the caller scales by 1/world, `dqn_agent.py:151` being a mean over the batch)."""
    out = part1()
    w1 = dist.all_reduce(flat_g[head:], op=dist.ReduceOp.SUM, group=group, async_op=True)
    part2()
    w2 = dist.all_reduce(flat_g[:head], op=dist.ReduceOp.SUM, group=group, async_op=True)
    w1.wait()
    w2.wait()
    return out
