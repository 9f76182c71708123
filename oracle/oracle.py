"""ORACLE / TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/env_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc.so")
_lib = None

OBS_SIZE = 726
MAXR = 4


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "env_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, u64, u32, dbl = C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_double
        L.orc_env_new.restype = vp
        L.orc_env_new.argtypes = [i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp,
                                  u64, u32, i32, i32, i32]
        L.orc_env_free.argtypes = [vp]
        L.orc_env_set_coefs.argtypes = [vp, dbl, dbl, dbl, dbl]
        L.orc_env_set_robot.argtypes = [vp, i32, i32, i32]
        L.orc_env_set_fire_step.argtypes = [vp, i32]
        L.orc_env_reset.argtypes = [vp, vp, vp]
        L.orc_env_step.argtypes = [vp, vp, vp, vp, vp]
        L.orc_env_export.argtypes = [vp] * 9
        L.orc_batch_step.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32]
        L.orc_batch_reset.argtypes = [vp, i32, vp, i32]
        L.orc_pairwise_sum.restype = dbl
        L.orc_pairwise_sum.argtypes = [vp, C.c_long]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class LayoutTables:
    """The static tables the oracle consumes (reference form: space + barrier list mask).
    Built either from a product ``Layout`` (tables only) or from a golden fixture."""

    def __init__(self, L, W, space, barrier, exits, obs_exit, ctr_box, danger_ctr, int_box, danger_int,
                 robot_range=(15, 30), robot_starts=((15, 15),), reset_obs_center=(15, 15)):
        self.L, self.W = int(L), int(W)
        self.space = np.ascontiguousarray(space, dtype=np.float64)
        self.barrier = np.ascontiguousarray(barrier, dtype=np.uint8)
        self.exits = np.ascontiguousarray(np.asarray(exits).reshape(-1, 2), dtype=np.int32)
        self.obs_exit = np.asarray(obs_exit, dtype=np.int32)
        self.ctr_box = np.asarray(ctr_box, dtype=np.int32)
        self.danger_ctr = np.ascontiguousarray(danger_ctr, dtype=np.float64)
        self.int_box = np.asarray(int_box, dtype=np.int32)
        self.danger_int = np.ascontiguousarray(danger_int, dtype=np.float64)
        self.robot_range = np.asarray(robot_range, dtype=np.int32)
        rs = np.zeros((MAXR, 2), dtype=np.int32)
        rs[:] = np.asarray(robot_starts[0], dtype=np.int32)
        rs[:len(robot_starts)] = np.asarray(robot_starts, dtype=np.int32)
        self.robot_starts = rs
        self.n_robots = len(robot_starts)
        self.reset_obs_center = np.asarray(reset_obs_center, dtype=np.int32)

    @classmethod
    def from_golden(cls, g, meta):
        """Tables EVALUATED BY THE REFERENCE (tests/golden/layout_*.npz written by oracle/make_golden.py layout_file: Map.space,
        barrier_list, both fire models per fire step) — nothing of the product's table builders is involved.  Needs a fixture
        whose ``steps`` are 0..S-1 on the full grid (no box)."""
        L, W = int(meta["width"]), int(meta["height"])
        steps = [int(s) for s in g["steps"]]
        assert steps == list(range(len(steps))) and "box" not in g, "from_golden needs contiguous fire steps on the full grid"
        pad = int(g["pad"])
        spec = meta.get("layout")
        exits = spec["exits"] if spec else [meta["exit"]]
        return cls(L, W, g["space"], g["barrier"], exits, exits[0], (0, 0, L + 2, W + 2), g["danger_ctr"],
                   (-pad, -pad, L + 2 + 2 * pad, W + 2 + 2 * pad), g["danger_int"])

    @classmethod
    def from_layout(cls, lay):
        return cls(lay.L, lay.W, lay.space, lay.barrier_mask, lay.exits, lay.obs_exit, lay.ctr_box, lay.danger_ctr,
                   lay.int_box, lay.danger_int, lay.robot_range, lay.robot_starts, lay.reset_obs_center)


class OracleEnv:
    def __init__(self, tables: LayoutTables, n_people, n_robots=1, seed=0, env_id=0, max_steps=1200,
                 reset_robots=None, reset_fire=0):
        self.t = tables
        self.N, self.R = int(n_people), int(n_robots)
        if reset_robots is None:
            reset_robots = 0 if n_robots == 1 else 1
        t = tables
        self.h = lib().orc_env_new(t.L, t.W, self.N, self.R, _p(t.space), _p(t.barrier), len(t.exits), _p(t.exits),
                                   _p(t.obs_exit), t.danger_ctr.shape[0], _p(t.ctr_box), _p(t.danger_ctr),
                                   _p(t.int_box), _p(t.danger_int), _p(t.robot_range), _p(t.robot_starts),
                                   _p(t.reset_obs_center), int(seed), int(env_id), int(max_steps),
                                   int(reset_robots), int(reset_fire))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_env_free(self.h)
            self.h = None

    def set_coefs(self, a, b, c, d):
        lib().orc_env_set_coefs(self.h, a, b, c, d)

    def set_robot(self, r, x, y):
        lib().orc_env_set_robot(self.h, r, x, y)

    def set_fire_step(self, s):
        lib().orc_env_set_fire_step(self.h, s)

    def reset(self, inject=None):
        obs = np.zeros((self.R, 11, 11, 6), dtype=np.float64)
        inj = None if inject is None else np.ascontiguousarray(inject, dtype=np.int16)
        lib().orc_env_reset(self.h, _p(inj), _p(obs))
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(np.atleast_1d(actions), dtype=np.int32)
        obs = np.zeros((self.R, 11, 11, 6), dtype=np.float64)
        r = np.zeros(1, dtype=np.float64)
        d = np.zeros(1, dtype=np.uint8)
        lib().orc_env_step(self.h, _p(a), _p(obs), _p(r), _p(d))
        return obs, float(r[0]), bool(d[0])

    def snapshot(self):
        t = self.t
        px = np.zeros(self.N, np.int16); py = np.zeros(self.N, np.int16)
        health = np.zeros(self.N, np.float64); acc = np.zeros(self.N, np.float64)
        flags = np.zeros(self.N, np.uint8)
        rmap = np.zeros((t.L + 2, t.W + 2), np.uint8)
        robots = np.zeros((MAXR, 2), np.int32); sc = np.zeros(8, np.int32)
        lib().orc_env_export(self.h, _p(px), _p(py), _p(health), _p(acc), _p(flags), _p(rmap), _p(robots), _p(sc))
        return dict(px=px, py=py, health=health, acc=acc, flags=flags, rmap=rmap, robots=robots[:self.R].astype(np.int16),
                    fire_step=np.int32(sc[0]), cur_step=np.int32(sc[1]), scalars=sc)


class OracleBatch:
    """n independent oracle envs stepped on ``threads`` host threads (bench cpu_baseline)."""

    def __init__(self, tables, n_envs, n_people, n_robots=1, seed=0, env_id_base=0, threads=1, **kw):
        self.envs = [OracleEnv(tables, n_people, n_robots, seed, env_id_base + k, **kw) for k in range(n_envs)]
        self.ptrs = (C.c_void_p * n_envs)(*[e.h for e in self.envs])
        self.n, self.R, self.threads = n_envs, n_robots, threads

    def reset(self):
        obs = np.zeros((self.n, self.R, 11, 11, 6), dtype=np.float32)
        lib().orc_batch_reset(self.ptrs, self.n, _p(obs), self.threads)
        return obs

    def step(self, actions, auto_reset=True):
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n, self.R)
        obs = np.zeros((self.n, self.R, 11, 11, 6), dtype=np.float32)
        r = np.zeros(self.n, np.float64); d = np.zeros(self.n, np.uint8)
        lib().orc_batch_step(self.ptrs, self.n, _p(a), int(auto_reset), _p(obs), _p(r), _p(d), self.threads)
        return obs, r, d


def pairwise_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().orc_pairwise_sum(_p(a), len(a))
