"""Data-parallel correctness on hardware — run under torchrun with 2 (or more) ranks, one GPU each, by
tests/test_multi_gpu.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dp_check.py

(1) the two-part backward + overlapped NCCL all-reduce + grad_scale of VecDQNAgent.grad_step equals the single-rank gradient
    of the GLOBAL batch (the gradient of dqn_agent.py:151's batch mean); (2) after K learn steps on
    rank-specific batches every replica holds bit-identical weights, Adam state and target network."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dqn_marl_b200.agents import qnet_params as qp
    from dqn_marl_b200.agents.dqn_agent import VecDQNAgent
    from dqn_marl_b200.agents.qnet import QNet

    out = {}
    for precision, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        Bl = 256                                        # per rank
        B = Bl * world
        torch.manual_seed(100 + rank)                   # different initial weights per rank: the constructor must broadcast rank 0's
        agent = VecDQNAgent(dev, dict(batch_size=Bl, precision=precision, dropout="train", seed=5, epsilon=0.3), n_envs=64)
        g = torch.Generator().manual_seed(7)            # the same global batch on every rank
        full = dict(states=((torch.rand((B, 11, 11, 6), generator=g) < 0.3).float() * torch.rand((B, 11, 11, 6), generator=g)).to(dev),
                    actions=torch.randint(0, 5, (B,), generator=g).to(dev), rewards=(torch.randn(B, generator=g) * 0.3).to(dev),
                    next_states=(torch.rand((B, 11, 11, 6), generator=g) < 0.3).float().to(dev),
                    dones=(torch.rand(B, generator=g) < 0.1).to(torch.uint8).to(dev))
        mask_on = (torch.rand((B, 512), generator=g) >= 0.2).to(torch.uint8).to(dev)
        mask_tg = (torch.rand((B, 512), generator=g) >= 0.2).to(torch.uint8).to(dev)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        mine = {k: v[sl].contiguous() for k, v in full.items()}
        hp = agent._hparams()
        hp.adam_step = 1
        loss = agent.grad_step(mine, hp, mask_on[sl].contiguous(), mask_tg[sl].contiguous())
        torch.cuda.synchronize(dev)
        g_dp = (agent.net.flat_g * agent._grad_scale()).clone()
        losses = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss.clone())
        # single-rank gradient of the global batch, same weights
        ref = QNet(dev, max_batch=B)
        ref.flat_p.copy_(agent.net.flat_p); ref.flat_t.copy_(agent.net.flat_t); ref.params_changed()
        ref.set_precision(precision)
        ref_loss = ref.td_backward(full, hp, mask_on, mask_tg)
        torch.cuda.synchronize(dev)
        a, b = g_dp.double(), ref.flat_g.double()
        rel = float((a - b).norm() / b.norm())
        per_tensor = {}
        for k, name in enumerate(qp.NAMES):
            x, y = a[qp.OFFSETS[k]:qp.OFFSETS[k] + qp.NUMEL[k]], b[qp.OFFSETS[k]:qp.OFFSETS[k] + qp.NUMEL[k]]
            per_tensor[name] = float((x - y).norm() / (y.norm() + 1e-30))
        loss_rel = abs(float(torch.stack(losses).mean()) - float(ref_loss)) / abs(float(ref_loss))
        assert rel <= tol, (precision, "gradient of the global batch", rel, per_tensor)
        assert max(per_tensor.values()) <= 10 * tol, (precision, per_tensor)
        assert loss_rel <= tol, (precision, "loss", loss_rel)

        # K learn steps on rank-specific batches: replicas stay bit-identical
        for step in range(6):
            gk = torch.Generator().manual_seed(1000 * (rank + 1) + step)
            bk = dict(states=(torch.rand((Bl, 11, 11, 6), generator=gk) < 0.3).float().to(dev), actions=torch.randint(0, 5, (Bl,), generator=gk).to(dev),
                      rewards=torch.randn(Bl, generator=gk).to(dev), next_states=(torch.rand((Bl, 11, 11, 6), generator=gk) < 0.3).float().to(dev),
                      dones=(torch.rand(Bl, generator=gk) < 0.1).to(torch.uint8).to(dev))
            agent.learn_on(bk)
            if step == 3:
                agent.update_target_network()
        torch.cuda.synchronize(dev)
        ident = True
        for name in ("flat_p", "flat_t", "flat_m", "flat_v"):
            t = getattr(agent.net, name)
            root = t.clone()
            dist.broadcast(root, src=0)
            ident = ident and bool(torch.equal(root.view(torch.int32), t.view(torch.int32)))
        flag = torch.tensor([int(ident)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1, (precision, "replicas diverged")
        out[precision] = {"grad_rel_err": rel, "loss_rel_err": loss_rel, "replicas_identical": True,
                          "param_checksum": float(agent.net.flat_p.double().sum().item())}
    if rank == 0:
        print("DP_CHECK " + json.dumps({"world": world, **out}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
