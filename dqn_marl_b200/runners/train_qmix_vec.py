"""Batched two-robot QMIX-lite loop: the reference's `runners/train_qmix.py` (:26-134) over thousands of `EvacuationEnvMulti`
instances per GPU.

What the reference does per env step (train_qmix.py:66-118), kept as is: two independent `DQNAgent`s act on their own robot's
window (:67-68); the JOINT transition (both windows, both actions, the shared scalar reward, done) goes into one replay
(:71); a learn step samples joint transitions (:77), mixes the two chosen Q-values through a monotonic network
`q_tot = relu([q1, q2] |W1| + b1) |W2| + b2` (:39-54, :92-96), bootstraps through the target networks and the target mixer
(:99-104), takes ONE mse loss (:107), clips each agent's and the mixer's gradient norm at 1.0 (:110-112), steps three Adam
optimizers (:113) and, with probability 0.01, hard-syncs both target networks and the target mixer (:116-118).

Here: the env batch is `VecEvacuationEnv` with a 2-robot layout; the joint replay is two device rings (one per robot) driven
with the same seed and draw counter, so both return the same sampled rows; each agent's Q-network runs on the CUDA kernels and
its output is an autograd node (`mq_qnet_backward`), so the 39-weight mixer stays a torch module and `loss.backward()`
fills both agents' gradient buffers; clip + Adam per agent is the fused `mq_qnet_clip_adam`.

    python -m dqn_marl_b200.runners.train_qmix_vec --envs 1024 --steps 200 --batch 1024
"""
from __future__ import annotations

import argparse
import random

import torch

from ..agents.dqn_agent import VecDQNAgent
from ..envs.vec_env import VecEvacuationEnv
from ..layout import Layout


class MixingNetwork(torch.nn.Module):
    """train_qmix.py:39-54 — monotonic mixer: absolute values of the weights, biases free."""

    def __init__(self, n_agents: int = 2, embed_dim: int = 32):
        super().__init__()
        self.fc1_weight = torch.nn.Parameter(torch.randn(n_agents, embed_dim))
        self.fc1_bias = torch.nn.Parameter(torch.zeros(embed_dim))
        self.fc2_weight = torch.nn.Parameter(torch.randn(embed_dim, 1))
        self.fc2_bias = torch.nn.Parameter(torch.zeros(1))

    def forward(self, q_vals: torch.Tensor) -> torch.Tensor:      # (batch, n_agents) -> (batch,)
        hidden = torch.relu(torch.matmul(q_vals, torch.abs(self.fc1_weight)) + self.fc1_bias)
        return (torch.matmul(hidden, torch.abs(self.fc2_weight)) + self.fc2_bias).squeeze(-1)


class VecQmixTrainer:
    def __init__(self, layout: Layout, n_envs: int, people: int, device, agent_cfg: dict, seed: int = 0, replay_capacity: int = 1 << 16,
                 target_sync_prob: float = 0.01, strict_reference: bool = False, mixer_lr: float = 1e-3):
        assert layout.n_robots == 2, "QMIX-lite mixes two robots (evacuation_env_multi.py:27)"
        self.env = VecEvacuationEnv(layout, n_envs, people, device=device, seed=seed, strict_reference=strict_reference, auto_reset=True)
        dev = self.env.device
        cfg = dict(agent_cfg)
        cfg.setdefault("memory_size", replay_capacity)
        cfg["seed"] = seed                                       # BOTH rings draw the same rows: one joint replay
        self.agents = [VecDQNAgent(dev, cfg, n_envs, 1, env_id_base=k * n_envs) for k in range(2)]      # distinct epsilon-greedy streams
        for k, a in enumerate(self.agents):
            a.memory.seed = seed
        self.gamma = cfg.get("gamma", 0.99)
        self.batch_size = self.agents[0].batch_size
        self.mixing, self.target_mixing = MixingNetwork().to(dev), MixingNetwork().to(dev)
        self.target_mixing.load_state_dict(self.mixing.state_dict())
        self.mix_optimizer = torch.optim.Adam(self.mixing.parameters(), lr=mixer_lr)
        self.target_sync_prob = target_sync_prob
        self.rng = random.Random(seed)
        E = n_envs
        self.obs = [torch.zeros((E, 2, 11, 11, 6), dtype=torch.float32, device=dev) for _ in range(2)]
        self.reward = torch.zeros((E,), dtype=torch.float64, device=dev)
        self.done = torch.zeros((E,), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((E, 2), dtype=torch.int32, device=dev)
        self.cur = 0
        self.learn_steps = 0
        self.target_syncs = 0
        self.obs[0].copy_(self.env.reset())

    def learn(self) -> torch.Tensor:
        a1, a2 = self.agents
        B = self.batch_size
        b1, b2 = a1.memory.sample(B), a2.memory.sample(B)                     # same (seed, draw id, size): the same joint rows
        with torch.enable_grad():
            q1 = a1.q_network(b1["states"]).gather(1, b1["actions"].unsqueeze(1)).squeeze(1)        # train_qmix.py:92-93
            q2 = a2.q_network(b2["states"]).gather(1, b2["actions"].unsqueeze(1)).squeeze(1)
            q_tot = self.mixing(torch.stack([q1, q2], dim=1))                                         # :95-96
            with torch.no_grad():                                                                     # :99-104
                nq1 = a1.target_network(b1["next_states"]).max(1)[0]
                nq2 = a2.target_network(b2["next_states"]).max(1)[0]
                y = b1["rewards"] + self.gamma * self.target_mixing(torch.stack([nq1, nq2], dim=1)) * (b1["dones"] == 0)
            loss = torch.nn.functional.mse_loss(q_tot, y)                                             # :107
            self.mix_optimizer.zero_grad()
            for a in self.agents:
                a.optimizer.zero_grad()
            loss.backward()                                        # one backward: mq_qnet_backward runs once per agent
        torch.nn.utils.clip_grad_norm_(self.mixing.parameters(), 1.0)                                 # :112
        self.mix_optimizer.step()
        for a in self.agents:                                      # clip_grad_norm_(q_network.parameters(), 1.0) + Adam step, fused (:110-113)
            a._adam_t += 1
            a.net.clip_adam(a._hparams(clip=1.0))
        self.learn_steps += 1
        if self.rng.random() < self.target_sync_prob:              # :116-118
            for a in self.agents:
                a.update_target_network()
            self.target_mixing.load_state_dict(self.mixing.state_dict())
            self.target_syncs += 1
        return loss.detach()

    def step(self, learn: bool = True):
        o, o2 = self.obs[self.cur], self.obs[self.cur ^ 1]
        for k, a in enumerate(self.agents):                                                            # :67-68
            self.actions[:, k] = a.act_batch(o[:, k].contiguous().unsqueeze(1), training=True).view(-1)
        self.env.step_into(self.actions, o2, self.reward, self.done)                                   # :69
        for k, a in enumerate(self.agents):                                                            # :71 — the joint transition, one ring per robot
            a.remember_batch(o[:, k].contiguous(), self.actions[:, k].contiguous(), self.reward, o2[:, k].contiguous(), self.done)
        self.cur ^= 1
        if learn and len(self.agents[0].memory) >= self.batch_size:                                    # :76
            return self.learn()
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--people", type=int, default=150)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    tr = VecQmixTrainer(Layout.reference_room(n_robots=2), args.envs, args.people, "cuda",
                        dict(batch_size=args.batch, learning_rate=1e-4, gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995),
                        seed=args.seed)
    for t in range(args.steps):
        loss = tr.step()
        if loss is not None and (t % 20 == 0 or t == args.steps - 1):
            print(f"step {t:5d}  qmix loss {loss.item():12.4f}  target syncs {tr.target_syncs}", flush=True)


if __name__ == "__main__":
    main()
