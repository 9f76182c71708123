"""GPU parity of the fused env kernels (csrc/env.cu) — called through the C-ABI (ctypes) — against
(1) golden trajectories of the unmodified reference and (2) the C oracle on fresh seeded inputs.
Bit-exact: positions, health, accumulators, flags, occupancy, robots, rewards (f64 bits), dones,
observations (the f64 view bit-exact, the f32 view == float32(reference obs))."""
import numpy as np
import pytest
import torch

from util import OP_RESET, OP_STEP, assert_frame_equal, layout_for, load_golden

pytestmark = pytest.mark.gpu

TRAJS = ["traj_room_single.npz", "traj_room_multi.npz", "traj_room_small.npz", "traj_room_westexit.npz",
         "traj_big256.npz", "traj_room_allevac.npz", "traj_synth_gallery.npz", "traj_synth_hall.npz", "traj_room_timelimit.npz",
         "traj_topexit.npz", "traj_westexit_far.npz"]


def _vec(lay, n_envs, N, seed, **kw):
    from dqn_marl_b200.envs import VecEvacuationEnv
    return VecEvacuationEnv(lay, n_envs, N, device="cuda:0", seed=seed, auto_reset=False, **kw)


@pytest.mark.parametrize("name", TRAJS)
def test_gpu_replays_reference_golden(name):
    g = load_golden(name)
    m = g["meta"]
    lay = layout_for(m)
    env = _vec(lay, 1, m["n_people"], m["seed"], max_steps=m.get("max_steps", 1200))     # traj_room_timelimit: done by time
    obs64 = torch.zeros((1, m["n_robots"], 11, 11, 6), dtype=torch.float64, device="cuda:0")
    F = len(g["op"])
    obs = env.reset(obs64=obs64)
    assert_frame_equal(g, 0, env.snapshot(0), obs64[0].cpu().numpy(), None, None, lay.L, lay.W, name)
    assert np.array_equal(obs[0].cpu().numpy(), g["obs"][0].astype(np.float32))
    for f in range(1, F):
        if g["op"][f] == OP_STEP:
            a = torch.tensor(g["actions"][f][None], dtype=torch.int32, device="cuda:0")
            obs, r, d = env.step(a, obs64=obs64)
            assert_frame_equal(g, f, env.snapshot(0), obs64[0].cpu().numpy(), r[0].item(), d[0].item(), lay.L, lay.W, name)
        else:
            obs = env.reset(obs64=obs64)
            assert_frame_equal(g, f, env.snapshot(0), obs64[0].cpu().numpy(), None, None, lay.L, lay.W, name)
        assert np.array_equal(obs[0].cpu().numpy(), g["obs"][f].astype(np.float32)), f"{name} frame {f}: f32 obs"


def _run_vs_oracle(lay, n_envs, N, seed, steps, n_act=6, check_every=1, strict=True, auto_reset=False, inject=None):
    from dqn_marl_b200.envs import VecEvacuationEnv
    from oracle import LayoutTables, OracleEnv
    R = lay.n_robots
    env = VecEvacuationEnv(lay, n_envs, N, device="cuda:0", seed=seed, auto_reset=auto_reset, strict_reference=strict,
                           env_id_base=100)
    tabs = LayoutTables.from_layout(lay)
    kw = {} if strict else dict(reset_robots=1, reset_fire=1)
    orcs = [OracleEnv(tabs, N, R, seed=seed, env_id=100 + k, **kw) for k in range(n_envs)]
    obs64 = torch.zeros((n_envs, R, 11, 11, 6), dtype=torch.float64, device="cuda:0")
    inj_t = None
    if inject is not None:
        inj_t = torch.tensor(inject, dtype=torch.int16, device="cuda:0")
    env.reset(obs64=obs64, inject_spawn=inj_t)
    o_obs = [o.reset(None if inject is None else inject[k]) for k, o in enumerate(orcs)]
    rng = np.random.default_rng(seed)

    def compare(tag, rew=None, done=None, o_rew=None, o_done=None):
        g_obs = obs64.cpu().numpy()
        for k in range(n_envs):
            a, b = env.snapshot(k), orcs[k].snapshot()
            for key in ("px", "py", "flags", "rmap", "robots", "fire_step", "cur_step"):
                assert np.array_equal(a[key], b[key]), f"{tag} env {k}: {key}"
            for key in ("health", "acc"):
                assert np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64)), f"{tag} env {k}: {key}"
            assert np.array_equal(a["scalars"][:8], b["scalars"][:8]), f"{tag} env {k}: scalars {a['scalars']} {b['scalars']}"
            assert np.array_equal(g_obs[k].view(np.uint64), o_obs[k].view(np.uint64)), f"{tag} env {k}: obs"
            if rew is not None:
                assert rew[k].view(np.uint64) == np.float64(o_rew[k]).view(np.uint64), f"{tag} env {k}: reward {rew[k]!r} {o_rew[k]!r}"
                assert bool(done[k]) == bool(o_done[k]), f"{tag} env {k}: done"

    compare("reset")
    for t in range(steps):
        acts = rng.integers(0, n_act, size=(n_envs, R)).astype(np.int32)
        _, r, d = env.step(torch.tensor(acts, device="cuda:0"), obs64=obs64)
        r, d = r.cpu().numpy(), d.cpu().numpy()
        o_rew, o_done = [], []
        for k, o in enumerate(orcs):
            ob, rr, dd = o.step(acts[k])
            if dd and auto_reset:
                ob = o.reset()
            o_obs[k] = ob
            o_rew.append(rr); o_done.append(dd)
        if (t % check_every) == 0 or t == steps - 1:
            compare(f"step {t}", r, d, o_rew, o_done)
        if not auto_reset and d.any():
            mask = torch.tensor(d, dtype=torch.uint8, device="cuda:0")
            env.reset(env_mask=mask, obs64=obs64)
            for k in range(n_envs):
                if d[k]:
                    o_obs[k] = orcs[k].reset()
            compare(f"reset after {t}")
    return env


def test_batch_matches_oracle_room():
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(), n_envs=24, N=150, seed=42, steps=90)


def test_batch_auto_reset_matches_oracle():
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(), n_envs=16, N=40, seed=7, steps=160, auto_reset=True, check_every=3)


def test_batch_non_strict_reset():
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(), n_envs=8, N=30, seed=11, steps=140, auto_reset=True, strict=False, check_every=2)


def test_two_robots_matches_oracle():
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(n_robots=2), n_envs=12, N=150, seed=3, steps=70)


def test_cooperative_scoring_full_ctas_match_oracle():
    """Warp-per-env variant with 28 envs per CTA: in a FULL CTA the warps score the movers of all 28 envs together (chunk
    prefix + round-robin); 60 envs = two cooperative CTAs + a partial one that works per warp.  Single robot, two robots
    (the R > 1 branch reads the other envs' robots from shared memory) and auto-reset (envs of one CTA in different episodes)."""
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(), n_envs=60, N=150, seed=21, steps=60, check_every=4)
    _run_vs_oracle(Layout.reference_room(n_robots=2), n_envs=31, N=150, seed=22, steps=40, check_every=4)
    _run_vs_oracle(Layout.reference_room(), n_envs=58, N=40, seed=23, steps=150, auto_reset=True, check_every=5)


def test_wide_group_serial_sums_with_many_hurt_people():
    """CTA-per-env variant (N > 256) in the small burning room: most people get hurt or die, so the run-based health sum takes
    every path (pure runs of 100.0, sparse hurt people, mostly-hurt batches, dead = +0.0) and the parallel pairwise tree sees
    shrinking n with several leaves; rewards are compared as f64 bits every other step."""
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(), n_envs=5, N=400, seed=31, steps=160, check_every=2)
    _run_vs_oracle(Layout.reference_room(), n_envs=3, N=700, seed=32, steps=120, auto_reset=True, check_every=3)
    # 3000 people in the same room: the 1024-thread variant with the person arrays in global scratch (prefetching chain)
    _run_vs_oracle(Layout.reference_room(), n_envs=2, N=3000, seed=33, steps=90, check_every=3)


def test_synthetic_multi_exit_matches_oracle():
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(96, 80, n_exits=3, seed=5)
    _run_vs_oracle(lay, n_envs=6, N=400, seed=9, steps=50, check_every=2)


def test_c3_shape_matches_oracle():
    """256x256 grid, 1000 people (BASELINE.json configs[2] shape), a few envs of the batch."""
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(256, 256, n_exits=1, seed=2024)
    _run_vs_oracle(lay, n_envs=4, N=1000, seed=2024, steps=40, check_every=4)


def test_edge_cases_one_person_and_crowding():
    """N = 1; and everybody injected onto the same few cells (duplicates: rmap is a flag map, quirks Q2-Q4)."""
    from dqn_marl_b200.layout import Layout
    lay = Layout.reference_room()
    _run_vs_oracle(lay, n_envs=4, N=1, seed=1, steps=60)
    inject = np.zeros((3, 64, 2), dtype=np.int16)
    inject[:, :, 0] = 30 + (np.arange(64) % 3)[None, :]
    inject[:, :, 1] = 14 + (np.arange(64) // 3 % 3)[None, :]
    _run_vs_oracle(lay, n_envs=3, N=64, seed=5, steps=50, inject=inject)


def test_robot_parked_off_map_and_reward_coefs():
    """no-robot policy of evaluate_strategies.py:83 (robot at [1000,1000]) and runtime reward coefficients
    (overnight_experiments.py:69-70)."""
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import Layout
    from oracle import LayoutTables, OracleEnv
    lay = Layout.reference_room()
    env = VecEvacuationEnv(lay, 2, 50, device="cuda:0", seed=8, auto_reset=False)
    orcs = [OracleEnv(LayoutTables.from_layout(lay), 50, 1, seed=8, env_id=k) for k in range(2)]
    env.reset()
    for o in orcs:
        o.reset()
    env.robots[:, 0, 0] = 1000; env.robots[:, 0, 1] = 1000
    env.scalars[:, 8] = 1000; env.scalars[:, 9] = 1000
    env.set_reward_coefs(10.0, 300.0, 1.5, 0.25)
    for o in orcs:
        o.set_robot(0, 1000, 1000)
        o.set_coefs(10.0, 300.0, 1.5, 0.25)
    obs64 = torch.zeros((2, 1, 11, 11, 6), dtype=torch.float64, device="cuda:0")
    for t in range(40):
        a = torch.full((2, 1), t % 5, dtype=torch.int32, device="cuda:0")
        _, r, d = env.step(a, obs64=obs64)
        for k, o in enumerate(orcs):
            ob, rr, dd = o.step([t % 5])
            assert r[k].item() == rr and bool(d[k].item()) == dd
            assert np.array_equal(obs64[k].cpu().numpy(), ob)


def test_full_size_c2_properties():
    """BASELINE.json configs[1] at full size (4096 envs x 150 people): size-independent invariants —
    conservation of people, occupancy bits only on valid cells, health monotone non-increasing,
    saved/dead flags absorbing, determinism of a repeated run."""
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import CELL_VALID, Layout
    lay = Layout.reference_room()
    E, N = 4096, 150

    def run():
        env = VecEvacuationEnv(lay, E, N, device="cuda:0", seed=123, auto_reset=False)
        env.reset()
        g = torch.Generator(device="cuda:0"); g.manual_seed(5)
        prev_h = env.health[:, :N].clone(); prev_f = env.flags[:, :N].clone()
        tot_r = torch.zeros(E, dtype=torch.float64, device="cuda:0")
        for t in range(30):
            a = torch.randint(0, 5, (E, 1), generator=g, device="cuda:0", dtype=torch.int32)
            _, r, d = env.step(a)
            tot_r += r
            h, f = env.health[:, :N], env.flags[:, :N]
            assert (h <= prev_h).all() and (h >= 0).all()
            assert ((f & prev_f) == prev_f).all()
            prev_h, prev_f = h.clone(), f.clone()
        return env, tot_r

    env, r1 = run()
    valid = torch.tensor((lay.cellinfo & CELL_VALID) != 0, device="cuda:0")
    rm = env.rmap_bytes().bool()
    assert not (rm & ~valid[None]).any()
    pos = env.pos[:, :N]
    x, y = (pos & 0xFFFF).long(), (pos >> 16).long()
    assert valid[x, y].all()
    sc = env.scalars
    assert ((env.flags[:, :N] & 1).sum(1) == sc[:, 6]).all() and (((env.flags[:, :N] >> 1) & 1).sum(1) == sc[:, 7]).all()
    _, r2 = run()
    assert torch.equal(r1, r2)


@pytest.mark.parametrize("name", ["traj_room_single.npz", "traj_room_multi.npz"])
def test_reference_facade_replays_golden(name):
    """The drop-in classes (same ctor / reset / step / info surface as evacuation_env.py) on the goldens."""
    from dqn_marl_b200.envs.evacuation_env import EvacuationEnv, EvacuationEnvMulti
    g = load_golden(name)
    m = g["meta"]
    cls = EvacuationEnvMulti if m["kind"] == "multi" else EvacuationEnv
    env = cls(m["width"], m["height"], None, m["exit"], m["n_people"], device="cuda:0", seed=m["seed"])
    assert env.state_size == (11, 11, 6) and env.action_size == 5 and env.max_steps == 1200
    F = min(len(g["op"]), 80)
    kept = None
    for f in range(1, F):
        if g["op"][f] == OP_STEP:
            a = g["actions"][f]
            state, reward, done, info = env.step([int(x) for x in a] if m["kind"] == "multi" else int(a[0]))
            if kept is None and m["kind"] == "single":
                kept = (f, info)                      # read only at the end: must still describe frame f
            assert isinstance(reward, float) and isinstance(done, bool)
            assert reward == g["reward"][f] and done == bool(g["done"][f])
            assert info["current_step"] == int(g["cur_step"][f]) and info["simulation_time"] == 0.5 * int(g["cur_step"][f])
            if m["kind"] == "single":
                assert info["robot_position"] == tuple(int(v) for v in g["robot_obs"][f])
                assert len(info["people_positions"]) == m["n_people"] and "health_values" in info
        else:
            state = env.reset()
        st = np.stack(state) if m["kind"] == "multi" else state[None]
        assert st.dtype == np.float64 and np.array_equal(st, g["obs"][f]), f"{name} frame {f}"
        p = env.people.list
        assert [int(q.pos[0]) for q in p] == g["px"][f].tolist() and [q.health for q in p] == g["health"][f].tolist()
        assert [q.savety for q in p] == [(b & 1) == 1 for b in g["flags"][f]]
    if kept is not None:
        f0, info0 = kept
        assert [int(p[0]) for p in info0["people_positions"]] == g["px"][f0].tolist() and info0["health_values"] == g["health"][f0].tolist()
        assert info0["evacuation_status"] == [(b & 1) == 1 for b in g["flags"][f0]]
    pm = env.get_performance_metrics()
    assert pm["evacuated"] + pm["dead"] + pm["remaining"] == m["n_people"]
    # trajectories are materialised on access: one {'pos', 'step'} entry per step since the last reset (evacuation_env.py:80,135)
    last_reset = max(f for f in range(F) if g["op"][f] != OP_STEP)
    tr = env.people.list[3].trajectory
    assert len(tr) == F - last_reset and tr[0]["step"] == 0 and tr[-1]["pos"] == env.people.list[3].pos
    assert [int(e["pos"][0]) for e in tr] == [int(g["px"][f][3]) for f in range(last_reset, F)]
    assert env.people.list[3].trajectory is tr and len(env.robot_trajectory) >= len(tr)
    assert np.array_equal(env.people.rmap != 0, np.unpackbits(g["rmap"][F - 1])[:(m["width"] + 2) * (m["height"] + 2)].reshape(m["width"] + 2, -1) != 0)
    assert env.map.Check_Valid(5, 5) and not env.map.Check_Valid(19, 15) and not env.map.Check_Valid(0, 3)


def test_large_env_global_scratch_matches_oracle():
    """Envs whose person arrays do not fit shared memory (6000 people): proposal table and per-person arrays live in
    global scratch, only the occupancy bitmap is staged in shared memory."""
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(160, 160, n_exits=4, seed=11)
    _run_vs_oracle(lay, n_envs=2, N=6000, seed=13, steps=14, check_every=2)


def test_c5_shape_matches_oracle():
    """BASELINE.json configs[4] shape: 1024x1024 multi-exit grid, 20000 people per env."""
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(1024, 1024, n_exits=8, wall_fill=0.15, seed=2024)
    _run_vs_oracle(lay, n_envs=2, N=20000, seed=5, steps=6, check_every=1)


def test_host_buffer_step_async_equals_device_step():
    """VecEvacuationEnv.step_async/step_wait (pinned host actions in, pinned host obs/reward/done out, own stream) gives
    exactly the trajectory of the device-resident step()."""
    from dqn_marl_b200.layout import Layout
    lay = Layout.reference_room()
    a = _vec(lay, 64, 150, seed=7)
    b = _vec(lay, 64, 150, seed=7)
    a.reset(); b.reset()
    rng = np.random.default_rng(3)
    for t in range(25):
        acts = torch.from_numpy(rng.integers(0, 5, size=(64, 1)).astype(np.int32))
        ha = acts.pin_memory()
        b.step_async(ha)
        obs, r, d = a.step(acts.cuda())
        ho, hr, hd = b.step_wait()
        assert torch.equal(obs.cpu(), ho) and torch.equal(r.cpu(), hr) and torch.equal(d.cpu(), hd), t
    for k in (0, 63):
        sa, sb = a.snapshot(k), b.snapshot(k)
        for key in ("px", "py", "health", "acc", "flags", "rmap"):
            assert np.array_equal(sa[key], sb[key]), (key, k)


@pytest.mark.gpu
def test_two_live_env_handles_of_one_kernel_variant():
    """The dynamic shared-memory limit is an attribute of the kernel function, not of the env handle: creating a second,
    smaller env of the same variant must not lower it under the first one.  Two handles stepped alternately, both bit-exact."""
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import Layout
    from oracle import LayoutTables, OracleEnv
    lay = Layout.reference_room()
    tabs = LayoutTables.from_layout(lay)
    big = VecEvacuationEnv(lay, 30, 150, device="cuda:0", seed=5, env_id_base=0)
    small = VecEvacuationEnv(lay, 30, 12, device="cuda:0", seed=6, env_id_base=0)
    o_big, o_small = OracleEnv(tabs, 150, 1, seed=5, env_id=3), OracleEnv(tabs, 12, 1, seed=6, env_id=3)
    big.reset(); small.reset(); o_big.reset(); o_small.reset()
    rng = np.random.default_rng(0)
    for t in range(12):
        for env, orc in ((big, o_big), (small, o_small)):
            acts = rng.integers(0, 5, size=(30, 1)).astype(np.int32)
            _, r, d = env.step(torch.tensor(acts, device="cuda:0"))
            _, rr, dd = orc.step(acts[3])
            assert r[3].item() == rr and bool(d[3].item()) == bool(dd), (t, r[3].item(), rr)
            a, b = env.snapshot(3), orc.snapshot()
            assert np.array_equal(a["px"], b["px"]) and np.array_equal(a["health"].view(np.uint64), b["health"].view(np.uint64))


@pytest.mark.gpu
def test_long_runs_with_auto_reset_match_oracle():
    """Several consecutive episodes per env (auto-reset inside the step kernel, envs of one CTA in different episodes and fire
    steps) for every kernel variant: warp-per-env with cooperative scoring, CTA-per-env, global-scratch envs."""
    from dqn_marl_b200.layout import Layout
    _run_vs_oracle(Layout.reference_room(), n_envs=140, N=150, seed=99, steps=900, auto_reset=True, check_every=30)
    _run_vs_oracle(Layout.reference_room(n_robots=2), n_envs=84, N=150, seed=98, steps=700, auto_reset=True, strict=False, check_every=35)
    _run_vs_oracle(Layout.synthetic(256, 256, n_exits=1, seed=2024), n_envs=10, N=1000, seed=77, steps=400, auto_reset=True, check_every=20)
    _run_vs_oracle(Layout.synthetic(96, 80, n_exits=3, seed=5), n_envs=20, N=400, seed=78, steps=500, auto_reset=True, strict=False, check_every=25)
    _run_vs_oracle(Layout.synthetic(160, 160, n_exits=4, seed=11), n_envs=3, N=6000, seed=79, steps=120, check_every=10)
