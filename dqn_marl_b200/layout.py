"""Static per-layout tables shared by every env instance of a batch.

Mirrors the *init-time* part of the reference's ``Map`` (reference
Louvre_Evacuation/envs/map.py:38-79 grid + barrier list, :127-148 floor field,
:93-113 exit test) and turns it into flat tables the CUDA kernels index:

  space      f64 [G]      static floor field, inf = wall / unreachable (map.py:148)
  dp5        f64 [G][8]   (space[c] - space[n_d]) * 5.0 per MoveTO direction d
                          (people.py:270,288; map.py:11-19), -inf when n_d fails
                          Check_Valid (map.py:85-92) -> never selected
  cellinfo   u8  [G]      bit0 Check_Valid, bit1 obs channel 3 (evacuation_env.py:109),
                          bit2 obs channel 4 (:113), bit3 checkSavefy (map.py:93-113)
  danger_ctr f64 [S][bw][bh]  danger at cell centres (people.py:205) on a bounding box
  danger_int f64 [S][iw][ih]  danger at integer coordinates (evacuation_env.py:106)

G = (L+2)*(W+2), cell index = x*(W+2)+y (the reference indexes space[x][y]).
The Dijkstra itself runs in the native library (csrc/floor_field.cpp).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .fire import FireSchedule

# map.py:11-19 — RIGHT, UP, LEFT, DOWN, then the four diagonals
MOVE_TO = ((1, 0), (0, -1), (-1, 0), (0, 1), (1, -1), (-1, -1), (-1, 1), (1, 1))

CELL_VALID = 1
CELL_OBS_BLOCKED = 2
CELL_OBS_EXIT = 4
CELL_EVACUATES = 8

MAX_ROBOTS = 4


def init_barrier(A, B):
    """map.py:25-33 — normalise a rectangle given by two corners."""
    if A[0] > B[0]:
        A, B = B, A
    x1, y1 = A[0], A[1]
    x2, y2 = B[0], B[1]
    if y1 < y2:
        return ((x1, y1), (x2, y2))
    return ((x1, y2), (x2, y1))


def fire_add_term(fire: FireSchedule, L: int, W: int) -> np.ndarray:
    """map.py:143-147 — 200 * danger(i,j)^2 at fire step 0, integer coordinates.  The reference squares an np.float64 scalar
    (``danger ** 2`` -> libm pow), so do exactly that per cell."""
    xs, ys = np.meshgrid(np.arange(L + 2), np.arange(W + 2), indexing="ij")
    d0 = fire.danger_field(0, xs, ys)
    add = np.zeros((L + 2, W + 2), dtype=np.float64)
    for (i, j) in zip(*np.nonzero(d0)):
        add[i, j] = 200 * (np.float64(d0[i, j]) ** 2)
    return add


def danger_tables(fire: FireSchedule, obs_fire: FireSchedule, L: int, W: int):
    """Danger at cell centres (people.py:205, the map's fire model) and at integer coordinates (evacuation_env.py:106, the
    env's own model) for every fire step, on their bounding boxes -> (ctr_box, danger_ctr, int_box, danger_int)."""
    S = fire.max_steps + 1
    bx0, by0, bx1, by1 = fire.bounding_box()
    cx0, cy0 = max(bx0, 0), max(by0, 0)
    cx1, cy1 = min(bx1, L + 2), min(by1, W + 2)
    cw, ch = max(cx1 - cx0, 1), max(cy1 - cy0, 1)
    gx, gy = np.meshgrid(np.arange(cx0, cx0 + cw), np.arange(cy0, cy0 + ch), indexing="ij")
    danger_ctr = np.stack([fire.danger_field(s, gx + 0.5, gy + 0.5) for s in range(S)])
    ox0, oy0, ox1, oy1 = obs_fire.bounding_box()
    ix0 = max(ox0, -6)
    ix1 = min(ox1, L + 8)
    iy0 = max(oy0, -6)
    iy1 = min(oy1, W + 8)
    iw, ih = max(ix1 - ix0, 1), max(iy1 - iy0, 1)
    gx, gy = np.meshgrid(np.arange(ix0, ix0 + iw), np.arange(iy0, iy0 + ih), indexing="ij")
    danger_int = np.stack([obs_fire.danger_field(s, gx, gy) for s in range(S)])
    return (cx0, cy0, cw, ch), danger_ctr, (ix0, iy0, iw, ih), danger_int


@dataclass
class Layout:
    L: int
    W: int
    exits: List[Tuple[int, int]]
    barriers: List[Tuple[Tuple[int, int], Tuple[int, int]]]
    robot_range: Tuple[int, int] = (15, 30)                 # map.py:75
    robot_starts: Sequence[Tuple[int, int]] = ((15, 15),)   # map.py:76 / evacuation_env_multi.py:27
    reset_obs_center: Tuple[int, int] = (15, 15)            # evacuation_env.py:64 (quirk Q7)
    obs_exit: Optional[Tuple[int, int]] = None              # evacuation_env.py:113 — exit_location
    fire: Optional[FireSchedule] = None
    obs_fire: Optional[FireSchedule] = None                 # the env's own model (quirk Q8)
    # derived
    space: np.ndarray = field(default=None, repr=False)
    dp5: np.ndarray = field(default=None, repr=False)
    cellinfo: np.ndarray = field(default=None, repr=False)
    barrier_mask: np.ndarray = field(default=None, repr=False)
    ctr_box: Tuple[int, int, int, int] = (0, 0, 0, 0)
    int_box: Tuple[int, int, int, int] = (0, 0, 0, 0)
    danger_ctr: np.ndarray = field(default=None, repr=False)
    danger_int: np.ndarray = field(default=None, repr=False)

    @property
    def G(self) -> int:
        return (self.L + 2) * (self.W + 2)

    @property
    def n_robots(self) -> int:
        return len(self.robot_starts)

    # ------------------------------------------------------------------
    @classmethod
    def reference_room(cls, width=36, height=30, exit_location=None, n_robots=1, floor_field=None):
        """The geometry hard-wired in EvacuationEnv.__init__ (evacuation_env.py:33-53): one
        exit, barrier (18,14)-(20,16) with a fire source at its centre; 2-robot variant
        starts at (10,15),(20,15) (evacuation_env_multi.py:27)."""
        if exit_location is None:
            exit_location = [36, 15]
        ex = (int(exit_location[0]), int(exit_location[1]))
        bar = [init_barrier((18, 14), (20, 16))]
        starts = ((15, 15),) if n_robots == 1 else ((10, 15), (20, 15)) + ((15, 15),) * (n_robots - 2)
        lay = cls(L=int(width), W=int(height), exits=[ex], barriers=bar, robot_starts=starts, obs_exit=ex)
        lay.build(floor_field)
        return lay

    @classmethod
    def synthetic(cls, L, W, n_exits=1, wall_fill=0.10, seed=2024, n_robots=1, floor_field=None):
        """Deterministic synthetic 'gallery' layout for the large benchmark shapes:
        random wall rectangles up to ``wall_fill`` of the area, the robot band x in [15,30] and the
        fire barrier neighbourhood kept free, exits on the edges (inside 1..L x 1..W, map.py:66-71)."""
        rng = np.random.Generator(np.random.PCG64(seed))
        occ = np.zeros((L + 2, W + 2), dtype=bool)
        bars = [init_barrier((18, 14), (20, 16))]
        occ[18:21, 14:17] = True
        target = wall_fill * L * W
        tries = 0
        while occ[1:L + 1, 1:W + 1].sum() < target and tries < 100000:
            tries += 1
            w = int(rng.integers(2, max(3, L // 12)))
            h = int(rng.integers(2, max(3, W // 12)))
            x0 = int(rng.integers(2, L - w - 1))
            y0 = int(rng.integers(2, W - h - 1))
            # keep the robot band, a corridor ring along the outer wall and previous rooms' margins free
            if x0 <= 31 and x0 + w >= 14:
                continue
            if occ[x0 - 1:x0 + w + 1, y0 - 1:y0 + h + 1].any():
                continue
            occ[x0:x0 + w, y0:y0 + h] = True
            bars.append(((x0, y0), (x0 + w - 1, y0 + h - 1)))
        exits = []
        cand = [(L, W // 2), (1, W // 2), (L // 2, 1), (L // 2, W), (L, W // 4), (1, 3 * W // 4), (L // 4, W), (3 * L // 4, 1)]
        for k in range(n_exits):
            exits.append(cand[k % len(cand)])
        starts = ((15, 15),) if n_robots == 1 else ((16, 15), (28, 15)) + ((22, 20),) * (n_robots - 2)
        lay = cls(L=L, W=W, exits=exits, barriers=bars, robot_starts=starts, obs_exit=exits[0])
        # only the first barrier burns (the reference would create one source per barrier, map.py:58-65)
        lay.fire = FireSchedule([((19.0, 15.0), (2, 2), 0.4)])
        lay.obs_fire = FireSchedule([((19, 15), (2, 2), 0.4)])
        lay.build(floor_field)
        return lay

    # ------------------------------------------------------------------
    def build(self, floor_field=None):
        L, W = self.L, self.W
        if self.fire is None:
            # map.py:58-65 — one source per barrier at its centre, size (2,2); default intensity 0.4
            self.fire = FireSchedule([(((A[0] + B[0]) / 2, (A[1] + B[1]) / 2), (2, 2), 0.4) for (A, B) in self.barriers])
        if self.obs_fire is None:
            # evacuation_env.py:46-53 — the env's own model: one source at (19,15)
            self.obs_fire = FireSchedule([((19, 15), (2, 2), 0.4)])
        if self.obs_exit is None:
            self.obs_exit = self.exits[0]

        # map.py:44-57 — wall ring + barrier rectangles; barrier_list membership
        wall = np.zeros((L + 2, W + 2), dtype=np.uint8)
        wall[0, :] = wall[L + 1, :] = 1
        wall[:, 0] = wall[:, W + 1] = 1
        for (A, B) in self.barriers:
            wall[A[0]:B[0] + 1, A[1]:B[1] + 1] = 1
        barrier = wall.copy()
        ex, ey = self.exits[0]
        wall[ex, ey] = 0                       # map.py:67 — only Exit[0] is forced open
        if ex == L:
            wall[ex + 1, ey] = 0               # map.py:68-71 (cell stays outside Check_Valid's bounds)
        if ey == W:
            wall[ex, ey + 1] = 0
        barrier[ex, ey] = 0                    # map.py:72-73
        self.barrier_mask = barrier

        add = fire_add_term(self.fire, L, W)

        if floor_field is None:
            from . import _lib
            floor_field = _lib.floor_field
        self.space = floor_field(L, W, wall, np.asarray(self.exits, dtype=np.int32), add)

        valid = np.isfinite(self.space)
        valid[0, :] = valid[L + 1, :] = False
        valid[:, 0] = valid[:, W + 1] = False
        info = np.zeros((L + 2, W + 2), dtype=np.uint8)
        info[valid] |= CELL_VALID
        info[(~valid) | (barrier != 0)] |= CELL_OBS_BLOCKED
        ox, oy = self.obs_exit
        if 0 <= ox <= L + 1 and 0 <= oy <= W + 1:
            info[ox, oy] |= CELL_OBS_EXIT
        for (exx, eyy) in self.exits:          # map.py:109-112 — Chebyshev distance <= 1 of any exit
            info[max(exx - 1, 0):exx + 2, max(eyy - 1, 0):eyy + 2] |= CELL_EVACUATES
        self.cellinfo = info

        dp5 = np.full((L + 2, W + 2, 8), -np.inf, dtype=np.float64)
        sp = self.space
        for d, (dx, dy) in enumerate(MOVE_TO):
            src = valid[1:L + 1, 1:W + 1]
            nb = valid[1 + dx:L + 1 + dx, 1 + dy:W + 1 + dy]
            ok = src & nb
            with np.errstate(invalid="ignore"):
                delta = (sp[1:L + 1, 1:W + 1] - sp[1 + dx:L + 1 + dx, 1 + dy:W + 1 + dy]) * 5.0
            dp5[1:L + 1, 1:W + 1, d] = np.where(ok, delta, -np.inf)
        self.dp5 = dp5

        # danger tables on bounding boxes
        self.ctr_box, self.danger_ctr, self.int_box, self.danger_int = danger_tables(self.fire, self.obs_fire, L, W)
        return self

    # convenience lookups used by the single-env facade and tests --------------------
    def danger_center(self, step, x, y):
        x0, y0, w, h = self.ctr_box
        step = min(int(step), self.danger_ctr.shape[0] - 1)
        if x0 <= x < x0 + w and y0 <= y < y0 + h:
            return float(self.danger_ctr[step, x - x0, y - y0])
        return 0.0

    def danger_integer(self, step, x, y):
        x0, y0, w, h = self.int_box
        step = min(int(step), self.danger_int.shape[0] - 1)
        if x0 <= x < x0 + w and y0 <= y < y0 + h:
            return float(self.danger_int[step, x - x0, y - y0])
        return 0.0


class DeviceLayoutBatch:
    """n layouts of one size whose floor fields and kernel tables are built ON THE DEVICE (per-env random
    layouts): `mq_floor_field_device` (the reference's Map.Init_Potential, map.py:127-148, as a relaxation to the same fixed
    point) -> `mq_layout_tables_device` (dp5 / cellinfo) -> `VecEvacuationEnv(DeviceLayoutBatch, ..., env_layout=...)`, which
    hands the device tables to `mq_env_create_layouts`.  Walls and exits differ per layout; the fire (one schedule, default:
    the reference room's source at (19, 15)), robot band and robot starts are shared, so the danger tables exist once.

        barrier  (n, L+2, W+2) uint8: Map.barrier_list membership — outer ring + barrier rectangles (map.py:43-57)
        exits    (n, max_exits, 2) int32, n_exits (n,) int32: exit cells inside 1..L x 1..W; exits[k, 0] is the env's
                 exit_location (forced open, observation channel 4, reward: map.py:66-73, evacuation_env.py:113,194)
    """

    def __init__(self, L: int, W: int, barrier, exits, n_exits, device="cuda", fire: Optional[FireSchedule] = None,
                 obs_fire: Optional[FireSchedule] = None, robot_range=(15, 30), robot_starts=((15, 15),), reset_obs_center=(15, 15)):
        import torch
        from . import _lib
        self.L, self.W = int(L), int(W)
        dev = torch.device(device)
        self.device = dev
        bar = torch.as_tensor(barrier).to(device=dev, dtype=torch.uint8).clone()
        exits = torch.as_tensor(exits).to(device=dev, dtype=torch.int32).contiguous()
        n_exits = torch.as_tensor(n_exits).to(device=dev, dtype=torch.int32).contiguous()
        n = bar.shape[0]
        assert bar.shape == (n, L + 2, W + 2) and exits.shape[0] == n and exits.shape[2] == 2
        self.n = n
        idx = torch.arange(n, device=dev)
        ex0, ey0 = exits[:, 0, 0].long(), exits[:, 0, 1].long()
        bar[idx, ex0, ey0] = 0                                   # map.py:67,72-73: Exit[0] is opened and leaves barrier_list
        self.barrier = bar
        self.exits, self.n_exits = exits, n_exits
        self.fire = fire or FireSchedule([((19.0, 15.0), (2, 2), 0.4)])
        self.obs_fire = obs_fire or FireSchedule([((19, 15), (2, 2), 0.4)])
        self.robot_range, self.robot_starts, self.reset_obs_center = tuple(robot_range), tuple(robot_starts), tuple(reset_obs_center)
        add = torch.from_numpy(fire_add_term(self.fire, L, W)).to(dev)
        self.space, self.sweeps = _lib.floor_field_device(L, W, bar, exits, n_exits, add.expand(n, L + 2, W + 2).contiguous())
        self.obs_exit = exits[:, 0, :].contiguous()
        self.dp5, self.cellinfo = _lib.layout_tables_device(L, W, self.space, bar, exits, n_exits, self.obs_exit)
        self.ctr_box, ctr, self.int_box, integ = danger_tables(self.fire, self.obs_fire, L, W)
        self.danger_ctr = torch.from_numpy(np.ascontiguousarray(ctr)).to(dev)
        self.danger_int = torch.from_numpy(np.ascontiguousarray(integ)).to(dev)

    @property
    def n_robots(self) -> int:
        return len(self.robot_starts)
