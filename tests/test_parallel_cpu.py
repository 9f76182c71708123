"""N > 1 host-side logic on CPU: world_size-2 gloo process group driving the SAME helpers the product calls on NCCL
(dqn_marl_b200/parallel.py): env sharding, the overlapped two-part gradient all-reduce of VecDQNAgent.grad_step, the
rank-consistent learn gate, min / max over ranks.  The CUDA half is tests/test_multi_gpu.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dqn_marl_b200.parallel import (allreduce_overlapped, env_shard, learn_gate_open, max_over_ranks, min_over_ranks,
                                        rank_world)
    assert rank_world() == (rank, world, rank)
    first, count = env_shard(rank, world, 4097)
    # the exchange step of VecDQNAgent.grad_step: part 1 fills the tail of the flat gradient, part 2 the head; afterwards every
    # rank holds the SUM over ranks, and part 1's return value (the loss) comes back
    n, head = 1000, 37
    g = torch.zeros(n)
    order = []

    def part1():
        g[head:] = float(rank + 1)
        order.append(1)
        return "loss"

    def part2():
        assert order == [1]
        g[:head] = 10.0 * (rank + 1)
        order.append(2)

    out = allreduce_overlapped(g, head, part1, part2)
    assert out == "loss" and order == [1, 2]
    # the gate opens on the same call count on both ranks although their shards differ by one env (2049 vs 2048 envs)
    min_push = min_over_ranks(count, "cpu")
    opened = next(p for p in range(1, 100) if learn_gate_open(p, min_push, 1 << 20, 4096, 0, 0))
    local_opened = next(p for p in range(1, 100) if p * count > 4096)
    mx = max_over_ranks(10.0 + rank, "cpu")
    q.put((rank, first, count, g[head:].unique().tolist(), g[:head].unique().tolist(), min_push, opened, local_opened, mx))
    dist.destroy_process_group()


def test_world2_gloo_shard_allreduce_and_gate():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, t0, h0, mp0, o0, lo0, m0), (r1, f1, c1, t1, h1, mp1, o1, lo1, m1) = res
    assert (f0, c0) == (0, 2049) and (f1, c1) == (2049, 2048)          # contiguous, covers 4097 ids exactly once
    assert t0 == t1 == [3.0] and h0 == h1 == [30.0]                      # sum over ranks, identical everywhere
    assert mp0 == mp1 == 2048 and o0 == o1 == 3                          # 3 * 2048 > 4096 on BOTH ranks ...
    assert (lo0, lo1) == (2, 3)                                          # ... whereas rank-local counts open one call apart
    assert m0 == m1 == 11.0


def test_learn_gate_semantics():
    from dqn_marl_b200.parallel import learn_gate_open
    assert not learn_gate_open(1, 32, 50000, 32, 0, 0)                   # len(memory) > batch_size is strict (train_dqn.py:117)
    assert learn_gate_open(2, 32, 50000, 32, 0, 0)
    assert not learn_gate_open(100, 32, 50000, 32, 0, 1000)              # default warmup_steps = 1000 never opens (dqn_agent.py:80,128)
    assert not learn_gate_open(10 ** 6, 64, 4096, 4096, 0, 0)            # a ring that cannot hold more than one batch


def test_env_shard_partitions():
    from dqn_marl_b200.parallel import env_shard
    for world in (1, 2, 3, 4, 8):
        for n in (8, 4096, 16384, 1001):
            ids = []
            for r in range(world):
                first, count = env_shard(r, world, n)
                ids += list(range(first, first + count))
            assert ids == list(range(n))
