"""N > 1 host-side logic on CPU: world_size-2 gloo process group (env sharding, gradient averaging, max-over-ranks)."""
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
    from dqn_marl_b200.parallel import allreduce_mean_, env_shard, max_over_ranks, rank_world
    assert rank_world() == (rank, world, rank)
    first, count = env_shard(rank, world, 4097)
    # every rank holds a different "gradient"; the mean must be identical everywhere afterwards
    g = torch.full((1000,), float(rank + 1))
    allreduce_mean_(g)
    mx = max_over_ranks(10.0 + rank, "cpu")
    q.put((rank, first, count, g[0].item(), g.std().item(), mx))
    dist.destroy_process_group()


def test_world2_gloo_shard_and_allreduce():
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
    (r0, f0, c0, g0, s0, m0), (r1, f1, c1, g1, s1, m1) = res
    assert (f0, c0) == (0, 2049) and (f1, c1) == (2049, 2048)          # contiguous, covers 4097 ids exactly once
    assert g0 == g1 == 1.5 and s0 == s1 == 0.0
    assert m0 == m1 == 11.0


def test_env_shard_partitions():
    from dqn_marl_b200.parallel import env_shard
    for world in (1, 2, 3, 4, 8):
        for n in (8, 4096, 16384, 1001):
            ids = []
            for r in range(world):
                first, count = env_shard(r, world, n)
                ids += list(range(first, first + count))
            assert ids == list(range(n))
