"""Batched DQN training loop: the reference's `runners/train_dqn.py` episode loop (act -> env.step -> remember ->
learn every step once len(memory) > batch_size, train_dqn.py:98-125) over thousands of envs per GPU, with every
tensor resident on the device and no host synchronisation inside the loop.

    python -m dqn_marl_b200.runners.train_dqn_vec --envs 4096 --people 150 --steps 200 --batch 1024
    torchrun --nproc-per-node 8 -m dqn_marl_b200.runners.train_dqn_vec ...      (env shards + gradient all-reduce)
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass

import torch

from ..agents.dqn_agent import VecDQNAgent
from ..envs.vec_env import VecEvacuationEnv
from ..layout import Layout
from ..parallel import env_shard, rank_world


@dataclass
class LoopStats:
    env_steps: int = 0
    learn_steps: int = 0
    last_loss: float = float("nan")


class VecTrainer:
    """One rank's share of the loop.  target_sync_every counts learn steps (the reference syncs every 50 episodes,
    train_dqn.py:124-125; with thousands of asynchronous episodes per step a step count is the batched analogue)."""

    def __init__(self, layout: Layout, n_envs: int, people: int, device, agent_cfg: dict, env_id_base: int = 0, seed: int = 0,
                 replay_capacity: int = 1 << 18, target_sync_every: int = 200, strict_reference: bool = False, process_group=None,
                 log_dir: str = None, overlap: bool = False):
        """log_dir: write RewardTracker / PerformanceRecorder-format episode logs there (runners/episode_log.py).  Logging needs
        the finished episodes' final state, so the env then runs WITHOUT in-kernel auto-reset: finished envs are re-spawned by
        a masked reset after the transition was stored — the stored next_state is the terminal observation, as in the reference
        (train_dqn.py:104-107), instead of the first observation of the next episode."""
        cfg = dict(agent_cfg)
        cfg.setdefault("memory_size", replay_capacity)
        cfg.setdefault("seed", seed)
        self.recorder = None
        # overlap: env step + replay push run on a second stream while sample + learn run on the caller's stream (the env kernel is
        # fp64 / ALU work, the learner tensor-core work).  The learn step then samples the ring as it was BEFORE this step's push
        # (the sample kernel is ordered before the push), i.e. it learns on data one env step older than the sequential loop.
        self.overlap = bool(overlap) and log_dir is None
        self.env = VecEvacuationEnv(layout, n_envs, people, device=device, seed=seed, env_id_base=env_id_base,
                                    strict_reference=strict_reference, auto_reset=log_dir is None)
        if log_dir is not None:
            from .episode_log import EpisodeRecorder
            self.recorder = EpisodeRecorder(n_envs, people, self.env.device, save_dir=log_dir)
        self.agent = VecDQNAgent(self.env.device, cfg, n_envs, layout.n_robots, env_id_base, process_group)
        dev, E, R = self.env.device, n_envs, layout.n_robots
        self.obs = [torch.zeros((E, R, 11, 11, 6), dtype=torch.float32, device=dev) for _ in range(2)]
        self.reward = torch.zeros((E,), dtype=torch.float64, device=dev)
        self.done = torch.zeros((E,), dtype=torch.uint8, device=dev)
        self.cur = 0
        self.target_sync_every = target_sync_every
        self.stats = LoopStats()
        self.obs[0].copy_(self.env.reset())
        if self.overlap:
            self._es = torch.cuda.Stream(device=dev)
            self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()
            torch.cuda.current_stream(dev).synchronize()

    def _step_overlapped(self, learn: bool):
        a, e = self.agent, self.env
        main = torch.cuda.current_stream(e.device)
        o, o2 = self.obs[self.cur], self.obs[self.cur ^ 1]
        main.wait_event(self._join)                                               # the previous env step has written `o`
        actions = a.act_batch(o, training=True)
        ready = learn and a.ready_to_learn()
        batch = a.sample_batch() if ready else None                               # reads the ring before this step's push
        self._fork.record(main)
        self._es.wait_event(self._fork)
        with torch.cuda.stream(self._es):
            e.step_into(actions, o2, self.reward, self.done)
            a.remember_batch(o, actions, self.reward, o2, self.done)
            self._join.record(self._es)
        self.cur ^= 1
        self.stats.env_steps += 1
        loss = None
        if ready:
            loss = a.learn_on(batch)
            self.stats.learn_steps += 1
            if self.stats.learn_steps % self.target_sync_every == 0:
                a.update_target_network()
        return loss

    def join(self):
        """Wait for everything the loop has enqueued (both streams)."""
        if self.overlap:
            torch.cuda.current_stream(self.env.device).wait_event(self._join)
        torch.cuda.current_stream(self.env.device).synchronize()

    def close(self):
        self.join()
        self.env.close(); self.agent.net.close(); self.agent.memory.close()

    def step(self, learn: bool = True):
        if self.overlap:
            return self._step_overlapped(learn)
        a, e = self.agent, self.env
        o, o2 = self.obs[self.cur], self.obs[self.cur ^ 1]
        actions = a.act_batch(o, training=True)                                   # dqn_agent.py:101
        e.step_into(actions, o2, self.reward, self.done)                          # evacuation_env.py:122
        a.remember_batch(o, actions, self.reward, o2, self.done)                  # dqn_agent.py:97
        if self.recorder is not None:
            self.recorder.record_step(self.reward, self.done, e)
            e.reset(env_mask=self.done, obs_out=o2)                               # evacuation_env.py:61 for the finished envs
        self.cur ^= 1
        self.stats.env_steps += 1
        loss = None
        if learn and a.ready_to_learn():                                          # train_dqn.py:117-118, rank-consistent
            loss = a.learn_device()
            self.stats.learn_steps += 1
            if self.stats.learn_steps % self.target_sync_every == 0:
                a.update_target_network()
        return loss


    # -- checkpoint of the whole loop ----------------------------------------------------------------
    def state_dict(self) -> dict:
        torch.cuda.synchronize(self.env.device)
        return {"env": self.env.state_dict(), "agent": self.agent.training_state_dict(), "cur": self.cur,
                "obs": self.obs[self.cur].cpu(), "stats": vars(self.stats).copy()}

    def load_state_dict(self, sd: dict):
        self.env.load_state_dict(sd["env"])
        self.agent.load_training_state_dict(sd["agent"])
        self.cur = int(sd["cur"])
        self.obs[self.cur].copy_(sd["obs"])
        for k, v in sd["stats"].items():
            setattr(self.stats, k, v)

    def save(self, path: str):
        torch.save(self.state_dict(), path)

    def load(self, path: str):
        self.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="global number of envs (sharded over ranks)")
    ap.add_argument("--people", type=int, default=150)
    ap.add_argument("--grid", type=int, nargs=2, default=[36, 30])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--log-dir", default=None, help="write reward_logs/reward_data.json, episode_data.csv, training_performance.csv here")
    ap.add_argument("--save", default=None, help="write a resumable checkpoint of env + replay + agent here at the end")
    ap.add_argument("--resume", default=None, help="continue from a checkpoint written by --save")
    args = ap.parse_args()
    rank, world, local = rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    first, count = env_shard(rank, world, args.envs)
    L, W = args.grid
    layout = Layout.reference_room(L, W) if (L, W) == (36, 30) else Layout.synthetic(L, W, seed=2024)
    torch.manual_seed(args.seed)
    tr = VecTrainer(layout, count, args.people, torch.device("cuda", local),
                    dict(batch_size=args.batch, learning_rate=1e-4, gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995),
                    env_id_base=first, seed=args.seed, log_dir=args.log_dir if rank == 0 else None)
    if args.resume:
        tr.load(args.resume if world == 1 else f"{args.resume}.rank{rank}")
    for t in range(args.steps):
        loss = tr.step()
        if rank == 0 and loss is not None and (t % 20 == 0 or t == args.steps - 1):
            sc = tr.env.scalars
            print(f"step {t:5d}  loss {loss.item():12.4f}  eps {tr.agent.epsilon:.4f}  evac/env {sc[:, 6].float().mean().item():.2f}"
                  f"  dead/env {sc[:, 7].float().mean().item():.2f}", flush=True)
    if tr.recorder is not None:
        tr.recorder.save_data()
    if args.save:
        tr.save(args.save if world == 1 else f"{args.save}.rank{rank}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
