"""EvacuationEnv / EvacuationEnvMulti — drop-in for the reference's single-environment classes.

Same constructor, attributes and return types as the reference
(Louvre_Evacuation/envs/evacuation_env.py:13-309, evacuation_env_multi.py:16-89) so that its runners
(`runners/train_dqn.py`, `train_double_dqn.py`, `train_qmix.py`, `evaluate_*.py`) can drive it unchanged;
underneath it is a batch of ONE env on the fused CUDA kernels (`VecEvacuationEnv`).  For throughput use
`VecEvacuationEnv` directly — this facade synchronises every step because the reference API hands back
Python scalars and numpy arrays.

Randomness: the reference draws from the global `random` / `np.random` streams; this build uses keyed
Philox draws.  The key is taken from `random.getrandbits(63)` at construction unless `seed=` is
given, so `random.seed(s)` before building the env still makes a run reproducible.
"""
from __future__ import annotations

import random
from typing import List, Optional

import numpy as np
import torch

from ..layout import CELL_VALID, Layout
from .vec_env import S_CUR_STEP, S_DEAD, S_EVAC, S_FIRE_STEP, S_RPX, S_RPY, VecEvacuationEnv


class _Person:
    """Read-only view with the attribute names of the reference's Person (people.py:8-23).  The values are read on access
    from the env's host mirror of the device state (one pinned copy per step, `EvacuationEnv._pull`), so a step costs no
    per-person Python work; `trajectory` (people.py:21, evacuation_env.py:80,135: one {'pos', 'step'} entry per step) is
    materialised from the per-step position records when somebody reads it."""
    __slots__ = ("id", "_i", "_env", "_traj", "_traj_upto")

    def __init__(self, pid, env=None):
        self.id = pid
        self._i = pid - 1
        self._env = env
        self._traj = []
        self._traj_upto = 0

    @property
    def pos(self):
        v = int(self._env._h_pos_np[self._i])
        return ((v & 0xFFFF) + 0.5, ((v >> 16) & 0xFFFF) + 0.5)

    @property
    def health(self):
        return float(self._env._h_health_np[self._i])

    @property
    def move_accumulator(self):
        return float(self._env._h_acc_np[self._i])

    @property
    def savety(self):
        return bool(self._env._h_flags_np[self._i] & 1)

    @property
    def dead(self):
        return bool(self._env._h_flags_np[self._i] & 2)

    @property
    def speed(self):
        """people.py:38-44"""
        h = self.health
        return 0.4 if h < 20 else 0.3 + 0.7 * h / 100

    @property
    def trajectory(self):
        rec = self._env._traj_rec
        if self._traj_upto > len(rec):            # the env was reset since the last read
            self._traj, self._traj_upto = [], 0
        for pos, label in rec[self._traj_upto:]:
            v = int(pos[self._i])
            self._traj.append({"pos": ((v & 0xFFFF) + 0.5, ((v >> 16) & 0xFFFF) + 0.5), "step": label})
        self._traj_upto = len(rec)
        return self._traj


class _FireModelView:
    """`env.map.fire_model` / `env.fire_model`: update() advances the (shared) device fire step
    (fire_model.py:63-67); get_max_danger evaluates the host-side schedule (fire_model.py:143-188)."""

    def __init__(self, env, schedule):
        self._env, self._sched = env, schedule

    def update(self):
        sc = self._env._vec.scalars
        sc[0, S_FIRE_STEP] = torch.clamp(sc[0, S_FIRE_STEP] + 1, max=self._sched.max_steps)
        self._env._h_sc_np[S_FIRE_STEP] = min(int(self._env._h_sc_np[S_FIRE_STEP]) + 1, self._sched.max_steps)

    def get_max_danger(self, position):
        step = int(self._env._h_sc_np[S_FIRE_STEP])
        return float(self._sched.danger_field(step, np.float64(position[0]), np.float64(position[1])))


class _MapView:
    """`env.map` with the attributes runners and tests touch (map.py:38-204)."""

    def __init__(self, env, layout: Layout):
        self._env, self._lay = env, layout
        self.Length, self.Width = layout.L, layout.W
        self.Exit = [tuple(e) for e in layout.exits]
        self.Barrier = list(layout.barriers)
        self.space = layout.space
        self.robot_range = tuple(layout.robot_range)
        self.barrier_list = [tuple(int(v) for v in c) for c in np.argwhere(layout.barrier_mask != 0)]
        self.fire_model = _FireModelView(env, layout.fire)

    def Check_Valid(self, x, y):
        x, y = int(x), int(y)
        if x >= self.Length + 1 or x <= 0 or y >= self.Width + 1 or y <= 0:
            return False
        return bool(self._lay.cellinfo[x, y] & CELL_VALID)

    def get_fire_danger(self, pos):
        return self.fire_model.get_max_danger(pos)

    def getDeltaP(self, P1, P2):
        return self.space[int(P1[0])][int(P1[1])] - self.space[int(P2[0])][int(P2[1])]

    # robot_position / robot_positions are writable in the reference (evaluate_strategies.py:83 parks the
    # robot at [1000, 1000]); writes go to the device state.
    @property
    def robot_positions(self):
        e = self._env
        return [[int(a), int(b)] for a, b in e._h_robots_np[:e._vec.n_robots]]

    @robot_positions.setter
    def robot_positions(self, value):
        e = self._env
        v = e._vec
        for r, p in enumerate(value[:v.n_robots]):
            v.robots[0, r, 0], v.robots[0, r, 1] = int(p[0]), int(p[1])
            e._h_robots_np[r, 0], e._h_robots_np[r, 1] = int(p[0]), int(p[1])

    @property
    def robot_position(self):
        sc = self._env._h_sc_np
        return [int(sc[S_RPX]), int(sc[S_RPY])]

    @robot_position.setter
    def robot_position(self, value):
        sc = self._env._vec.scalars
        sc[0, S_RPX], sc[0, S_RPY] = int(value[0]), int(value[1])
        self._env._h_sc_np[S_RPX], self._env._h_sc_np[S_RPY] = int(value[0]), int(value[1])


class _PeopleView:
    """`env.people`: list of person views + rmap as float64 like the reference (people.py:158-163)."""

    def __init__(self, env, n):
        self._env = env
        self.list = [_Person(i + 1, env) for i in range(n)]
        self.tot = n

    @property
    def rmap(self):
        return self._env._vec.rmap_bytes()[0].cpu().numpy().astype(np.float64)


class LazyInfo(dict):
    """step() info dict; the O(N) lists of evacuation_env.py:160-170 are built on first access — from copies of THIS step's
    person arrays, so an info dict that is read later still describes the step that returned it."""

    _LAZY = ("people_positions", "health_values", "evacuation_status")

    def __init__(self, env, eager):
        super().__init__(eager)
        self._snap = (env._traj_rec[-1][0], env._h_health_np.copy(), env._h_flags_np.copy())

    def __missing__(self, key):
        if key in self._LAZY:
            pos, health, flags = self._snap
            self["people_positions"] = [((int(v) & 0xFFFF) + 0.5, ((int(v) >> 16) & 0xFFFF) + 0.5) for v in pos]
            self["health_values"] = [float(h) for h in health]
            self["evacuation_status"] = [bool(f & 1) for f in flags]
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        return key in self._LAZY or dict.__contains__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default


class EvacuationEnv:
    """Single robot (evacuation_env.py:13)."""
    EVAC_REWARD: float = 50.0
    DEATH_PENALTY: float = 200.0
    DEATH_ACC_PENALTY: float = 0.5
    ALIVE_BONUS: float = 1.0
    _N_ROBOTS = 1

    def __init__(self, width=36, height=30, fire_zones=None, exit_location=None, num_people=150,
                 device="cuda", seed: Optional[int] = None):
        self.width, self.height, self.num_people = width, height, num_people
        self.time_per_step = 0.5
        self.max_simulation_time = 600
        self.max_steps = int(self.max_simulation_time / self.time_per_step)
        if exit_location is None:
            exit_location = [36, 15]
        if fire_zones is None:
            fire_zones = {(18, 14), (19, 14), (20, 14), (18, 15), (19, 15), (20, 15), (18, 16), (19, 16), (20, 16)}
        self.exit_location = exit_location
        self.fire_zones = fire_zones          # ignored by the simulation, as in the reference (quirk Q9)
        self.state_size = (11, 11, 6)
        self.action_size = 5
        self.num_robots = self._N_ROBOTS
        if seed is None:
            seed = random.getrandbits(63)
        self._layout = Layout.reference_room(width, height, exit_location, n_robots=self._N_ROBOTS)
        self._vec = VecEvacuationEnv(self._layout, 1, num_people, device=device, seed=seed, strict_reference=True,
                                     auto_reset=False, max_steps=self.max_steps, reward_coefs=self._coefs())
        self._coefs_sent = self._coefs()
        dev = self._vec.device
        self._obs64 = torch.zeros((1, self._N_ROBOTS, 11, 11, 6), dtype=torch.float64, device=dev)
        self._act = torch.zeros((1, self._N_ROBOTS), dtype=torch.int32, device=dev)
        # Host mirror of everything a step hands back (pinned): the step enqueues the action upload, the kernel and these
        # copies on one stream and synchronises ONCE; person / map views read the mirror.
        v, N = self._vec, num_people

        def pinned(t):
            return torch.empty(t.shape, dtype=t.dtype).pin_memory()
        self._mirror = [(pinned(v.pos[0, :N]), v.pos[0, :N]), (pinned(v.health[0, :N]), v.health[0, :N]),
                        (pinned(v.acc[0, :N]), v.acc[0, :N]), (pinned(v.flags[0, :N]), v.flags[0, :N]),
                        (pinned(v.robots[0]), v.robots[0]), (pinned(v.scalars[0]), v.scalars[0]),
                        (pinned(self._obs64[0]), self._obs64[0]), (pinned(v.reward), v.reward), (pinned(v.done), v.done)]
        (self._h_pos_np, self._h_health_np, self._h_acc_np, self._h_flags_np, self._h_robots_np, self._h_sc_np, self._h_obs_np,
         self._h_rew_np, self._h_done_np) = [h.numpy() for h, _ in self._mirror]
        self._h_pos_np = self._h_pos_np.view(np.uint32)
        self._h_act = torch.zeros((1, self._N_ROBOTS), dtype=torch.int32).pin_memory()
        self._h_act_np = self._h_act.numpy()
        self._traj_rec = []
        self.map = _MapView(self, self._layout)
        self.fire_model = _FireModelView(self, self._layout.obs_fire)
        self.people = _PeopleView(self, num_people)
        self.reset()

    # class attributes are mutated at runtime by overnight_experiments.py:69-70
    def _coefs(self):
        c = type(self)
        return (float(c.EVAC_REWARD), float(c.DEATH_PENALTY), float(c.DEATH_ACC_PENALTY), float(c.ALIVE_BONUS))

    def _sync_coefs(self):
        c = self._coefs()
        if c != self._coefs_sent:
            self._vec.set_reward_coefs(*c)
            self._coefs_sent = c

    def _pull(self, step_label=None, reset=False):
        """Device -> host mirror: nine small asynchronous copies into pinned memory and ONE synchronisation (the reference API
        hands back Python scalars and numpy arrays, so a step has to wait for the device exactly once)."""
        for h, d in self._mirror:
            h.copy_(d, non_blocking=True)
        torch.cuda.current_stream(self._vec.device).synchronize()
        sc = self._h_sc_np
        self.current_step = int(sc[S_CUR_STEP])
        self.time = self.current_step * self.time_per_step
        self._evac, self._dead = int(sc[S_EVAC]), int(sc[S_DEAD])
        if reset:
            self._traj_rec = [(self._h_pos_np.copy(), 0)]              # evacuation_env.py:80
        else:
            self._traj_rec.append((self._h_pos_np.copy(), step_label))     # evacuation_env.py:135

    def _state(self):
        o = self._h_obs_np.copy()
        return o[0] if self._N_ROBOTS == 1 else [o[r] for r in range(self._N_ROBOTS)]

    def reset(self):
        """evacuation_env.py:61-82"""
        self._vec.reset(obs64=self._obs64)
        self.robot_direction = 1
        self.prev_evacuated = 0
        self.prev_dead = 0
        self._pull(reset=True)
        self.robot_trajectory = [(tuple(self.map.robot_position), 0)]
        return self._state()

    def _get_state(self):
        return self._state()

    def step(self, action):
        """evacuation_env.py:122-172"""
        self._sync_coefs()
        step_label = self.current_step
        acts = action if self._N_ROBOTS > 1 else [action]
        for r, a in enumerate(acts):
            self._h_act_np[0, r] = int(a) if isinstance(a, (int, np.integer)) or hasattr(a, "__int__") else -1
        self._act.copy_(self._h_act, non_blocking=True)
        self._vec.step(self._act, obs64=self._obs64)
        self._pull(step_label=step_label)
        reward = float(self._h_rew_np[0])
        done = bool(self._h_done_np[0])
        if self._N_ROBOTS == 1:
            self.robot_trajectory.append((tuple(self.map.robot_position), step_label))
        else:
            for p in self.map.robot_positions:
                self.robot_trajectory.append((tuple(p), step_label))
        return self._state(), reward, done, self._info()

    def _info(self):
        return LazyInfo(self, {
            "robot_position": tuple(self.map.robot_position),
            "fire_spread": [],
            "evacuation_rate": self._evac / self.num_people,
            "death_rate": self._dead / self.num_people,
            "current_step": self.current_step,
            "simulation_time": self.time,
        })

    def get_performance_metrics(self):
        """evacuation_env.py:290-309"""
        alive = [float(h) for h in self._h_health_np[(self._h_flags_np & 2) == 0]]
        return {
            "evacuated": self._evac, "dead": self._dead, "remaining": self.num_people - self._evac - self._dead,
            "evacuation_rate": self._evac / self.num_people, "death_rate": self._dead / self.num_people,
            "avg_health": np.mean(alive) if alive else float("nan"), "min_health": min(alive, default=100),
            "total_steps": self.current_step, "total_time": self.time,
        }


class EvacuationEnvMulti(EvacuationEnv):
    """Two robots starting at (10,15) and (20,15); list-of-states API (evacuation_env_multi.py:16-89)."""
    _N_ROBOTS = 2

    def reset(self) -> List[np.ndarray]:
        super().reset()
        self.robot_trajectory = [(tuple(p), 0) for p in self.map.robot_positions]
        return self._state()

    def _get_joint_state(self):
        return self._state()

    def step(self, actions: List[int]):
        assert len(actions) == self.num_robots, "one action per robot"
        return super().step(actions)

    def _info(self):
        return {
            "robot_positions": [tuple(p) for p in self.map.robot_positions],
            "evacuation_rate": self._evac / self.num_people, "death_rate": self._dead / self.num_people,
            "current_step": self.current_step, "simulation_time": self.time,
        }
