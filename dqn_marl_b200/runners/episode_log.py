"""Episode logs of the batched runner in the file formats of the reference's `RewardTracker` and `PerformanceRecorder`
(reference Louvre_Evacuation/utils/reward_visualizer.py:97-122 `save_data`: reward_data.json + episode_data.csv;
utils/visualization.py:24-39,55-57 + runners/train_dqn.py:204-206: training_performance.csv), so that the reference's own
plotting / analysis code (`RewardTracker.load_data`, reward_visualizer.py:124-139) reads a batched run.

The reference records one episode at a time on the host.  Here thousands of envs finish episodes on different steps, on the
device: every step appends one row per env (done flag, episode return, length, evacuated, dead, mean / min health of the
living) to a device-side staging buffer with a handful of elementwise torch ops — no host synchronisation — and the buffer
is copied to the host once every `flush_every` steps, where finished episodes are appended in (step, env) order.
"""
from __future__ import annotations

import csv
import json
import os
from collections import deque

import numpy as np
import torch


class EpisodeRecorder:
    WINDOW = 100                      # reward_visualizer.py:32 window_size

    def __init__(self, n_envs: int, num_people: int, device, save_dir: str = "dqn_results", flush_every: int = 64):
        self.E, self.N, self.dev = int(n_envs), int(num_people), torch.device(device)
        self.save_dir = save_dir
        self.flush_every = int(flush_every)
        self.ep_return = torch.zeros(self.E, dtype=torch.float64, device=self.dev)
        self.ep_len = torch.zeros(self.E, dtype=torch.int32, device=self.dev)
        # staging: [flush_every][E] rows of (done, return, steps, evacuated, dead, avg_health, min_health) + mean step reward
        self._stage = torch.zeros((self.flush_every, 7, self.E), dtype=torch.float64, device=self.dev)
        self._stage_step_reward = torch.zeros(self.flush_every, dtype=torch.float64, device=self.dev)
        self._rows = 0
        # RewardTracker fields (reward_visualizer.py:24-37)
        self.episode_rewards, self.episode_steps = [], []
        self.episode_evacuation_rates, self.episode_death_rates = [], []
        self.step_rewards = []
        self.recent_rewards = deque(maxlen=self.WINDOW)
        self.total_episodes = 0
        self.total_steps = 0
        # PerformanceRecorder.episode_data (visualization.py:28-39)
        self.episode_data = []

    @torch.no_grad()
    def record_step(self, reward: torch.Tensor, done: torch.Tensor, env) -> None:
        """Call after env.step(auto_reset=False) and BEFORE the masked reset: `env` still holds the finished episodes' state."""
        self.ep_return += reward
        self.ep_len += 1
        N = self.N
        alive = (env.flags[:, :N] & 2) == 0                               # not dead (get_performance_metrics, evacuation_env.py:296-297)
        h = env.health[:, :N]
        n_alive = alive.sum(1)
        row = self._stage[self._rows]
        row[0] = done.to(torch.float64)
        row[1] = self.ep_return
        row[2] = self.ep_len.to(torch.float64)
        row[3] = env.scalars[:, 6].to(torch.float64)                      # MQ_S_EVAC
        row[4] = env.scalars[:, 7].to(torch.float64)                      # MQ_S_DEAD
        row[5] = torch.where(alive, h, torch.zeros_like(h)).sum(1) / n_alive.clamp(min=1).to(torch.float64)
        row[5] = torch.where(n_alive > 0, row[5], torch.full_like(row[5], float("nan")))
        row[6] = torch.where(alive, h, torch.full_like(h, 100.0)).amin(1)  # min(..., default=100)
        self._stage_step_reward[self._rows] = reward.mean()
        keep = (done == 0)
        self.ep_return *= keep.to(torch.float64)
        self.ep_len *= keep.to(torch.int32)
        self._rows += 1
        if self._rows == self.flush_every:
            self.flush()

    def flush(self) -> None:
        if self._rows == 0:
            return
        stage = self._stage[:self._rows].cpu().numpy()
        self.step_rewards.extend(float(v) for v in self._stage_step_reward[:self._rows].cpu().numpy())
        for t in range(self._rows):
            for e in np.nonzero(stage[t, 0])[0]:
                ret, steps, evac, dead = float(stage[t, 1, e]), int(stage[t, 2, e]), int(stage[t, 3, e]), int(stage[t, 4, e])
                self.episode_rewards.append(ret)
                self.episode_steps.append(steps)
                self.episode_evacuation_rates.append(evac / self.N)
                self.episode_death_rates.append(dead / self.N)
                self.recent_rewards.append(ret)
                self.total_steps += steps
                self.episode_data.append({
                    "episode": self.total_episodes, "total_reward": ret, "evacuated": evac, "dead": dead,
                    "remaining": self.N - evac - dead, "evacuation_rate": evac / self.N, "death_rate": dead / self.N,
                    "avg_health": float(stage[t, 5, e]), "min_health": float(stage[t, 6, e]), "total_steps": steps})
                self.total_episodes += 1
        self._rows = 0

    # reward_visualizer.py:60-76
    def get_statistics(self) -> dict:
        if not self.episode_rewards:
            return {}
        r = np.asarray(self.episode_rewards)
        return {"total_episodes": self.total_episodes, "total_steps": self.total_steps, "avg_reward": float(r.mean()),
                "max_reward": float(r.max()), "min_reward": float(r.min()), "std_reward": float(r.std()),
                "recent_avg_reward": float(np.mean(self.recent_rewards)) if self.recent_rewards else 0,
                "avg_evacuation_rate": float(np.mean(self.episode_evacuation_rates)),
                "avg_death_rate": float(np.mean(self.episode_death_rates)),
                "avg_steps_per_episode": float(np.mean(self.episode_steps))}

    def save_data(self) -> None:
        """reward_visualizer.py:97-122 (<save_dir>/reward_logs/reward_data.json, episode_data.csv) and train_dqn.py:204-206
        (<save_dir>/training_performance.csv).  `step_rewards` holds the mean reward over the env batch per step."""
        self.flush()
        logs = os.path.join(self.save_dir, "reward_logs")
        os.makedirs(logs, exist_ok=True)
        data = {"episode_rewards": self.episode_rewards, "episode_steps": self.episode_steps,
                "episode_evacuation_rates": self.episode_evacuation_rates, "episode_death_rates": self.episode_death_rates,
                "step_rewards": self.step_rewards, "statistics": self.get_statistics()}
        with open(os.path.join(logs, "reward_data.json"), "w", encoding="utf-8") as f:
            json.dump(data, f, ensure_ascii=False, indent=2)
        with open(os.path.join(logs, "episode_data.csv"), "w", newline="", encoding="utf-8") as f:
            w = csv.writer(f)
            w.writerow(["episode", "reward", "steps", "evacuation_rate", "death_rate"])
            for k in range(len(self.episode_rewards)):
                w.writerow([k, self.episode_rewards[k], self.episode_steps[k], self.episode_evacuation_rates[k], self.episode_death_rates[k]])
        cols = ["episode", "total_reward", "evacuated", "dead", "remaining", "evacuation_rate", "death_rate", "avg_health",
                "min_health", "total_steps"]
        with open(os.path.join(self.save_dir, "training_performance.csv"), "w", newline="", encoding="utf-8") as f:
            w = csv.DictWriter(f, fieldnames=cols)
            w.writeheader()
            w.writerows(self.episode_data)
