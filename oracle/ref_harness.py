"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path.

Drives the UNMODIFIED Python reference (imported from /root/reference, never
copied) with *keyed* random draws so that its trajectories can be compared bit
for bit with the C restatement (oracle/env_oracle.c) and the CUDA kernels.

How: ``Louvre_Evacuation/envs/people.py`` does ``import random`` /
``import numpy as np`` at module scope (people.py:1-2); we rebind those two
module globals to proxy objects that forward everything except the four draw
sites (people.py:69-75, :186-190, :239, :290), which are answered from
``oracle/keyed_draws.py``.  The proxy looks at the caller's frame to learn which
person / direction the draw belongs to.

This file only works where /root/reference exists (the build container).  The
golden vectors it produces are committed under tests/golden/ by
``oracle/make_golden.py``; nothing on the GPU box imports this module.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.dont_write_bytecode = True
_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
from keyed_draws import Draws  # noqa: E402

REFERENCE_ROOT = os.environ.get("MARL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Louvre_Evacuation", "envs"))


def _import_reference():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import Louvre_Evacuation.envs.people as people_mod
    import Louvre_Evacuation.envs.map as map_mod
    import Louvre_Evacuation.envs.fire_model as fire_mod
    import Louvre_Evacuation.envs.evacuation_env as env_mod
    import Louvre_Evacuation.envs.evacuation_env_multi as envm_mod
    return people_mod, map_mod, fire_mod, env_mod, envm_mod


class _RandomProxy:
    """Stands in for the ``random`` module inside people.py."""

    def __init__(self, real, draws: Draws):
        self._real = real
        self._d = draws
        self._spawn_person = -1
        self._spawn_calls = 0
        self._spawn_x = 0
        self.dims = (0, 0)

    def __getattr__(self, name):
        return getattr(self._real, name)

    # people.py:290 — called from find_best_direction(x, y) <- run()
    def uniform(self, a, b):
        dire = sys._getframe(1).f_locals["dire"]
        p = sys._getframe(2).f_locals["p"]
        u = self._d.noise_u(p.id - 1, dire)
        return a + (b - a) * u

    # people.py:239 — only movers[0] matters for state (losers bump thmap only)
    def shuffle(self, movers):
        if len(movers) < 2:
            return
        best = min(range(len(movers)), key=lambda k: (self._d.prio(movers[k][0].id - 1), movers[k][0].id - 1))
        movers[0], movers[best] = movers[best], movers[0]

    # people.py:186-190 — randint(1, L-2) then randint(1, W-2), repeated while the cell is invalid
    def randint(self, a, b):
        i = sys._getframe(1).f_locals["i"]
        if i != self._spawn_person:
            self._spawn_person = i
            self._spawn_calls = 0
        attempt, axis = divmod(self._spawn_calls, 2)
        self._spawn_calls += 1
        L, W = self.dims
        x, y = self._d.spawn_xy(i, attempt, L, W)
        return x if axis == 0 else y

    def new_episode(self):
        self._spawn_person = -1
        self._spawn_calls = 0


class _NpRandomProxy:
    def __init__(self, real, draws: Draws):
        self._real = real
        self._d = draws

    def __getattr__(self, name):
        return getattr(self._real, name)

    # people.py:69-75 — called from Person.update_health(self, danger_level)
    def uniform(self, lo, hi):
        person = sys._getframe(1).f_locals["self"]
        u = self._d.health_u(person.id - 1)
        # must be np.float64 (not float): see SURVEY.md Appendix A on sum() at evacuation_env.py:245
        return np.float64(lo + (hi - lo) * u)


class _NpProxy:
    def __init__(self, real, draws: Draws):
        self._real = real
        self.random = _NpRandomProxy(real.random, draws)

    def __getattr__(self, name):
        return getattr(self._real, name)


class RefEnv:
    """One reference environment under keyed draws.

    kind = 'single' -> EvacuationEnv, 'multi' -> EvacuationEnvMulti (2 robots).
    """

    def __init__(self, kind="single", width=36, height=30, exit_location=None, num_people=150,
                 seed=0, env_id=0, layout=None):
        """layout = dict(exits=[(x, y), ...], barriers=[((x0, y0), (x1, y1)), ...], fire_first_only=bool) replaces the
        room hard-wired in EvacuationEnv.__init__ (evacuation_env.py:42-53) by the reference's own
        ``Map(L, W, exits, [Init_Barrier(A, B), ...])`` (map.py:38-79: multi-exit Dijkstra, one fire source per
        barrier) — SURVEY.md §8(c).  fire_first_only keeps only the first barrier's source (what
        dqn_marl_b200.layout.Layout.synthetic builds) by overriding map.fire_model and re-running Init_Potential."""
        import random as _random
        people_mod, map_mod, fire_mod, env_mod, envm_mod = _import_reference()
        self.mods = (people_mod, map_mod, fire_mod, env_mod, envm_mod)
        self.draws = Draws(seed, env_id)
        self._rp = _RandomProxy(_random, self.draws)
        self._rp.dims = (width, height)
        self._np = _NpProxy(np, self.draws)
        self.kind = kind
        self.L, self.W, self.N = width, height, num_people
        self.R = 2 if kind == "multi" else 1
        self.tick = 0
        self.episode = 0
        self._install()
        try:
            cls = envm_mod.EvacuationEnvMulti if kind == "multi" else env_mod.EvacuationEnv
            self.draws.episode = 0
            self._rp.new_episode()
            self.env = cls(width, height, None, exit_location, num_people)   # ctor spawns episode 0
            if layout is not None:
                assert kind == "single"
                exits = [list(map(int, e)) for e in layout["exits"]]
                bars = [map_mod.Init_Barrier(tuple(A), tuple(B)) for (A, B) in layout["barriers"]]
                m = map_mod.Map(width, height, exits, bars)
                if layout.get("fire_first_only"):
                    (A, B) = bars[0]
                    m.fire_model = fire_mod.FireSpreadModel([fire_mod.FireSource(
                        center=((A[0] + B[0]) / 2, (A[1] + B[1]) / 2), size=(2, 2), temp_max=900, co_max=1800)])
                    m.Init_Potential()
                self.env.map = m
                self.env.exit_location = exits[0]
                self.draws.episode = 0
                self._rp.new_episode()
                self.env.reset()                                             # episode 0 again, on the new map
        finally:
            self._uninstall()
        self.last_obs = None

    # --- proxy plumbing -------------------------------------------------
    def _install(self):
        pm = self.mods[0]
        self._saved = (pm.random, pm.np)
        pm.random = self._rp
        pm.np = self._np

    def _uninstall(self):
        pm = self.mods[0]
        pm.random, pm.np = self._saved

    # --- driver -----------------------------------------------------------
    def reset(self):
        self.episode += 1
        self.draws.episode = self.episode
        self._rp.new_episode()
        self._install()
        try:
            obs = self.env.reset()
        finally:
            self._uninstall()
        return self._obs(obs)

    def step(self, action):
        self.draws.tick = self.tick
        self._install()
        try:
            obs, reward, done, info = self.env.step(action)
        finally:
            self._uninstall()
        self.tick += 1
        return self._obs(obs), float(reward), bool(done), info

    def observe(self):
        """Current observation without stepping (evacuation_env.py:84 / evacuation_env_multi.py:44)."""
        e = self.env
        return self._obs(e._get_joint_state() if self.kind == "multi" else e._get_state())

    def _obs(self, obs):
        o = np.stack(obs) if self.kind == "multi" else np.asarray(obs)[None]
        return np.ascontiguousarray(o, dtype=np.float64)      # (R, 11, 11, 6)

    # --- state extraction -------------------------------------------------
    def snapshot(self):
        e = self.env
        pl = e.people.list
        px = np.array([int(p.pos[0]) for p in pl], dtype=np.int16)
        py = np.array([int(p.pos[1]) for p in pl], dtype=np.int16)
        health = np.array([float(p.health) for p in pl], dtype=np.float64)
        acc = np.array([float(p.move_accumulator) for p in pl], dtype=np.float64)
        flags = np.array([(1 if p.savety else 0) | (2 if p.dead else 0) for p in pl], dtype=np.uint8)
        rmap = (np.asarray(e.people.rmap) != 0).astype(np.uint8)        # (L+2, W+2)
        robots = np.array(e.map.robot_positions, dtype=np.int16).reshape(-1, 2)
        return dict(px=px, py=py, health=health, acc=acc, flags=flags, rmap=rmap, robots=robots,
                    robot_obs=np.array(e.map.robot_position, dtype=np.int16),
                    fire_step=np.int32(e.fire_model.progressive_model.current_step),
                    cur_step=np.int32(e.current_step))

    def layout_tables(self, steps=range(0, 181), pad=6, box=None):
        """Static tables the GPU build consumes, evaluated with the reference's own code:
        space (map.py:127-148), danger at cell centres per fire step (people.py:205 via
        fire_model.py:143-188) and at integer coordinates on a padded box (evacuation_env.py:106)."""
        e = self.env
        L, W = self.L, self.W
        space = np.array(e.map.space, dtype=np.float64)
        barrier = np.zeros((L + 2, W + 2), dtype=np.uint8)
        for (bx, by) in e.map.barrier_list:
            barrier[bx, by] = 1
        steps = list(steps)
        # box = (x0, y0, x1, y1): evaluate only there (large grids); tables are still full-size, 0 elsewhere
        bx0, by0, bx1, by1 = box if box is not None else (-pad, -pad, L + 2 + pad, W + 2 + pad)
        ctr = np.zeros((len(steps), L + 2, W + 2), dtype=np.float64)
        integ = np.zeros((len(steps), L + 2 + 2 * pad, W + 2 + 2 * pad), dtype=np.float64)
        for which, fm in (("map", e.map.fire_model), ("env", e.fire_model)):
            pm = fm.progressive_model
            saved = (pm.current_step, pm.current_fire_sources)
            try:
                for k, s in enumerate(steps):
                    pm.current_step = s
                    pm.current_fire_sources = pm._interpolate_fire_sources()
                    if which == "map":
                        for x in range(max(0, bx0), min(L + 2, bx1)):
                            for y in range(max(0, by0), min(W + 2, by1)):
                                ctr[k, x, y] = fm.get_max_danger((x + 0.5, y + 0.5))
                    else:
                        for x in range(max(-pad, bx0), min(L + 2 + pad, bx1)):
                            for y in range(max(-pad, by0), min(W + 2 + pad, by1)):
                                integ[k, x + pad, y + pad] = fm.get_max_danger((x, y))
            finally:
                pm.current_step, pm.current_fire_sources = saved
        return dict(space=space, barrier=barrier, danger_ctr=ctr, danger_int=integ,
                    steps=np.array(steps, dtype=np.int32), pad=np.int32(pad))
