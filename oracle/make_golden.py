"""ORACLE / TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Each trajectory file holds F frames; frame 0 is the state right after the env constructor (which
spawns episode 0, evacuation_env.py:59), every later frame the state after one op:
    op = 0  step(actions[f])     (evacuation_env.py:122)
    op = 1  reset()              (evacuation_env.py:61)
Pin: numpy / torch / python versions of this container are recorded in each file's ``meta``.
"""
from __future__ import annotations

import json
import os
import platform
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from keyed_draws import philox4x32  # noqa: E402
from ref_harness import RefEnv  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
OP_STEP, OP_RESET = 0, 1


def meta(**kw):
    import torch
    d = dict(numpy=np.__version__, torch=torch.__version__, python=platform.python_version(),
             machine=platform.processor() or platform.machine(), reference="LX-530/DQN-MARL @ /root/reference")
    d.update(kw)
    return np.array(json.dumps(d))


def actions_for(seed, t, R, n_act=6):
    """Keyed test actions in 0..5 (5 is invalid on purpose: map.py:180-181 ignores it)."""
    w = philox4x32(0, t, 0, 99, seed)
    return [int((w[r] * n_act) >> 32) for r in range(R)]


def record(kind, W, H, exit_loc, N, seed, n_steps, name, extra_resets=(), layout=None, max_time=None, forced_actions=None):
    R = 2 if kind == "multi" else 1
    ref = RefEnv(kind, W, H, exit_loc, N, seed=seed, layout=layout)
    if max_time is not None:
        # shorter episodes through the reference's own instance attribute: `time_is_up = self.time >= self.max_simulation_time`
        # (evacuation_env.py:27,153) — the time-limit branch of `done`, which no full-length episode of 150 people reaches
        ref.env.max_simulation_time = max_time
        ref.env.max_steps = int(max_time / ref.env.time_per_step)
    frames = []

    def grab(op, act, obs, reward, done):
        s = ref.snapshot()
        s.update(op=np.int8(op), actions=np.array(act, dtype=np.int32), obs=obs, reward=np.float64(reward),
                 done=np.uint8(done), rmap=np.packbits(s["rmap"], axis=None))
        frames.append(s)

    grab(OP_RESET, [0] * R, ref.observe(), 0.0, 0)
    t = 0
    while t < n_steps:
        act = actions_for(seed, t, R)
        if forced_actions and t in forced_actions:            # steer the robot (the actions are stored per frame anyway)
            act = list(forced_actions[t])
        obs, rew, done, _ = ref.step(act if kind == "multi" else act[0])
        grab(OP_STEP, act, obs, rew, done)
        t += 1
        if done or t in extra_resets:
            obs = ref.reset()
            grab(OP_RESET, [0] * R, obs, 0.0, 0)
    out = {k: np.stack([f[k] for f in frames]) for k in frames[0]}
    extra = {} if layout is None else dict(layout=dict(exits=[list(map(int, e)) for e in layout["exits"]],
                                                       barriers=[[list(map(int, A)), list(map(int, B))] for (A, B) in layout["barriers"]],
                                                       fire_first_only=bool(layout.get("fire_first_only", False))))
    if max_time is not None:
        extra["max_steps"] = int(max_time / 0.5)
    out["meta"] = meta(kind=kind, width=W, height=H, exit=list(exit_loc) if exit_loc else [36, 15], n_people=N,
                       seed=seed, n_robots=R, **extra)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **out)
    print(name, "frames", len(frames), "resets", int((out["op"][1:] == OP_RESET).sum()),
          "kB", os.path.getsize(path) // 1024)
    return ref


def layout_file(ref, name, steps, box=None):
    t = ref.layout_tables(steps=steps, box=box)
    probe_in = -np.linspace(0.0, 4.0, 64)
    t["exp_probe_in"] = probe_in
    t["exp_probe_out"] = np.exp(probe_in)       # lets a test tell whether this machine's np.exp matches
    t["meta"] = meta(L=ref.L, W=ref.W)
    if box is not None:
        # large grid: keep only the evaluated box of the danger tables
        pad = int(t["pad"])
        x0, y0, x1, y1 = box
        t["danger_ctr"] = t["danger_ctr"][:, max(0, x0):x1, max(0, y0):y1]
        t["danger_int"] = t["danger_int"][:, x0 + pad:x1 + pad, y0 + pad:y1 + pad]
        t["box"] = np.array(box, dtype=np.int32)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **t)
    print(name, "kB", os.path.getsize(path) // 1024)


def synthetic_specs():
    """Multi-exit / multi-barrier layouts run through the reference's own Map + People:
    (a) the geometry dqn_marl_b200.layout.Layout.synthetic generates (exits, barrier rectangles; only barrier 0 burns), so the
        benchmark layouts' generator is pinned to the reference's floor field, spawn rule, conflict logic and reward;
    (b) a hand-written hall with three exits and four barriers, EVERY barrier burning (map.py:58-65 as is)."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(96, 80, n_exits=3, wall_fill=0.10, seed=7)
    a = dict(exits=[tuple(e) for e in lay.exits], barriers=[(tuple(A), tuple(B)) for (A, B) in lay.barriers], fire_first_only=True)
    b = dict(exits=[(64, 32), (1, 20), (30, 1)],
             barriers=[((18, 14), (20, 16)), ((24, 30), (27, 33)), ((40, 10), (44, 12)), ((8, 40), (12, 44))],
             fire_first_only=False)
    return a, b


def main_synthetic():
    a, b = synthetic_specs()
    ref = record("single", 96, 80, list(a["exits"][0]), 300, 31, 70, "traj_synth_gallery.npz", extra_resets=(40,), layout=a)
    layout_file(ref, "layout_synth_gallery.npz", steps=[0, 3, 40, 180], box=(-6, -6, 50, 46))
    ref = record("single", 64, 48, list(b["exits"][0]), 200, 32, 90, "traj_synth_hall.npz", extra_resets=(55,), layout=b)
    layout_file(ref, "layout_synth_hall.npz", steps=range(0, 101))        # every fire step the trajectory visits


def main_branches():
    """Two trajectories aimed at branches of the reference the other goldens leave cold (found by tracing the reference while
    this script runs: evacuation_env.py:114 and map.py:70-73 were never executed; what stays cold after this is unreachable
    code — evacuation_env.py:218, people.py:79 — or off the path):
    * traj_topexit — 40 x 20 room whose exit (17, 20) lies on the TOP edge, inside the robot's observation window from the
      first frame (observation channel 4 = 1, evacuation_env.py:113-114), opened through a barrier rectangle that covers it
      (`space[ex][ey+1] = 1` for an exit on the top edge and `barrier_list.remove((ex, ey))`, map.py:70-73);
    * traj_westexit_far — west exit, the robot STEERED (+y past the fire barrier, then +x) to the east end of its range, where
      further +x moves are refused (`15 <= x <= 30`, map.py:194) and everybody it influences is more than 20 cells from the
      exit (guidance tier 2.0, evacuation_env.py:207-208)."""
    top = dict(exits=[(17, 20)], barriers=[((18, 12), (20, 14)), ((17, 20), (18, 20))], fire_first_only=True)
    record("single", 40, 20, [17, 20], 80, 811, 60, "traj_topexit.npz", extra_resets=(35,), layout=top)
    record("single", 40, 24, [1, 12], 150, 812, 50, "traj_westexit_far.npz", forced_actions={**{t: [3] for t in range(3)}, **{t: [0] for t in range(3, 18)}})       # +y past the fire barrier, then +x


def main_timelimit():
    # 150 people, episodes cut at 12.5 simulated seconds = 25 steps: `done` by time with people still inside, reset, again
    record("single", 36, 30, None, 150, 4321, 60, "traj_room_timelimit.npz", max_time=12.5)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--synthetic-only" in sys.argv:
        return main_synthetic()
    if "--timelimit-only" in sys.argv:
        return main_timelimit()
    if "--branches-only" in sys.argv:
        return main_branches()
    ref = record("single", 36, 30, None, 150, 1234, 260, "traj_room_single.npz", extra_resets=(7,))
    layout_file(ref, "layout_room.npz", steps=range(0, 181))
    record("multi", 36, 30, None, 150, 99, 120, "traj_room_multi.npz")
    record("single", 36, 30, [36, 15], 20, 5, 200, "traj_room_small.npz")
    record("single", 40, 24, [1, 12], 60, 77, 120, "traj_room_westexit.npz")
    # six people, all of them reach the exit at step 50: the completion bonus branch of _calculate_reward
    # (evacuation_env.py:253-265), done by evacuation, reset, and a second episode
    record("single", 36, 30, None, 6, 100, 90, "traj_room_allevac.npz")
    ref = record("single", 256, 256, [256, 128], 1000, 2024, 24, "traj_big256.npz")
    layout_file(ref, "layout_big256.npz", steps=[0, 5, 24, 90, 180], box=(-6, -6, 50, 46))
    main_synthetic()
    main_timelimit()
    main_branches()


if __name__ == "__main__":
    main()
