"""The oracle (oracle/env_oracle.c) pinned against trajectories recorded from the UNMODIFIED Python
reference (oracle/make_golden.py).  Bit-exact: positions, health, accumulators, flags, rmap, robots,
observations (float64), rewards (float64 bits), dones."""
import numpy as np
import pytest

from util import OP_RESET, OP_STEP, assert_frame_equal, layout_for, load_golden

TRAJS = ["traj_room_single.npz", "traj_room_multi.npz", "traj_room_small.npz", "traj_room_westexit.npz",
         "traj_big256.npz", "traj_room_allevac.npz", "traj_synth_gallery.npz", "traj_synth_hall.npz", "traj_room_timelimit.npz",
         "traj_topexit.npz", "traj_westexit_far.npz"]


@pytest.mark.parametrize("name", TRAJS)
def test_oracle_replays_reference(name):
    from oracle import LayoutTables, OracleEnv
    g = load_golden(name)
    m = g["meta"]
    lay = layout_for(m)
    # traj_room_timelimit: episodes cut by the reference's `time >= max_simulation_time` (evacuation_env.py:153)
    env = OracleEnv(LayoutTables.from_layout(lay), m["n_people"], m["n_robots"], seed=m["seed"], max_steps=m.get("max_steps", 1200))
    F = len(g["op"])
    obs = env.reset()
    assert_frame_equal(g, 0, env.snapshot(), obs, None, None, lay.L, lay.W, name)
    for f in range(1, F):
        if g["op"][f] == OP_STEP:
            obs, r, d = env.step(g["actions"][f])
            assert_frame_equal(g, f, env.snapshot(), obs, r, d, lay.L, lay.W, name)
        else:
            obs = env.reset()
            assert_frame_equal(g, f, env.snapshot(), obs, None, None, lay.L, lay.W, name)


@pytest.mark.parametrize("traj,layout", [("traj_room_single.npz", "layout_room.npz"), ("traj_synth_hall.npz", "layout_synth_hall.npz")])
def test_oracle_replays_reference_on_reference_tables(traj, layout):
    """Same replay with the static tables taken from the reference's own evaluation (Map.space, barrier_list, both fire
    models) instead of the product's Layout: the oracle's parity does not lean on the product's table builders."""
    from oracle import LayoutTables, OracleEnv
    g = load_golden(traj)
    m = g["meta"]
    tabs = LayoutTables.from_golden(load_golden(layout), m)
    env = OracleEnv(tabs, m["n_people"], m["n_robots"], seed=m["seed"])
    L, W = m["width"], m["height"]
    obs = env.reset()
    assert_frame_equal(g, 0, env.snapshot(), obs, None, None, L, W, traj)
    for f in range(1, len(g["op"])):
        if g["op"][f] == OP_STEP:
            if tabs.danger_ctr.shape[0] < 181 and env.snapshot()["fire_step"] + 1 >= tabs.danger_ctr.shape[0]:
                break                      # beyond the fire steps the fixture tabulates
            obs, r, d = env.step(g["actions"][f])
            assert_frame_equal(g, f, env.snapshot(), obs, r, d, L, W, traj)
        else:
            obs = env.reset()
            assert_frame_equal(g, f, env.snapshot(), obs, None, None, L, W, traj)


def test_pairwise_sum_matches_numpy():
    """np.mean at evacuation_env.py:228 is numpy's pairwise add.reduce (third party, numpy 2.3.x)."""
    from oracle import pairwise_sum
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [127, 128, 129, 150, 255, 256, 257, 999, 1000, 1001, 4097, 20000, 30001]:
        a = rng.uniform(0, 300, size=n)
        assert pairwise_sum(a) == np.add.reduce(a), n
        if n:
            assert pairwise_sum(a) / n == np.mean(a), n


def test_keyed_draws_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    from keyed_draws import philox4x32, philox4x32_np, sample_indices
    assert philox4x32(0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    ones = 0xffffffff
    assert philox4x32(ones, ones, ones, ones, (ones << 32) | ones) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, (0x299f31d0 << 32) | 0xa4093822) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    w = philox4x32_np([1, 2, 3], 7, [4, 5, 6], 0, 0xABCDEF0123)
    for k in range(3):
        assert tuple(int(x) for x in w[:, k]) == philox4x32(k + 1, 7, k + 4, 0, 0xABCDEF0123)
    idx = sample_indices(3, 11, 5000, 512)
    assert len(set(idx.tolist())) == 512 and idx.min() >= 0 and idx.max() < 5000
