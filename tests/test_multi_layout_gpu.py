"""several layouts in ONE env batch (`mq_env_create_layouts`, a layout index per env) and the device-side
table pipeline `mq_floor_field_device` -> `mq_layout_tables_device` -> env kernels.  Every env is checked bit for bit
against the C oracle running on ITS layout's host tables (which tests/test_layout.py pins to the reference's Map)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _layouts_36x30():
    """Four rooms of the reference's size that differ in exit, barriers and robot start."""
    from dqn_marl_b200.layout import Layout, init_barrier
    base = Layout.reference_room()
    west = Layout.reference_room(36, 30, [1, 12])
    two = Layout(L=36, W=30, exits=[(36, 15), (18, 1)], barriers=[init_barrier((18, 14), (20, 16)), init_barrier((25, 5), (27, 9))],
                 obs_exit=(36, 15), robot_starts=((16, 14),)).build()
    many = Layout(L=36, W=30, exits=[(20, 30)], barriers=[init_barrier((18, 14), (20, 16)), init_barrier((5, 20), (9, 22)), init_barrier((30, 3), (33, 6))],
                  obs_exit=(20, 30)).build()
    return [base, west, two, many]


def _compare(env, orcs, obs64, tag):
    g_obs = obs64.cpu().numpy()
    for k, o in enumerate(orcs):
        a, b = env.snapshot(k), o.snapshot()
        for key in ("px", "py", "flags", "rmap", "robots", "fire_step", "cur_step"):
            assert np.array_equal(a[key], b[key]), f"{tag} env {k}: {key}"
        for key in ("health", "acc"):
            assert np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64)), f"{tag} env {k}: {key}"
        assert np.array_equal(g_obs[k].view(np.uint64), o._last_obs.view(np.uint64)), f"{tag} env {k}: obs"


def _run_mixed(layouts_for_gpu, host_layouts, env_layout, N, seed, steps, auto_reset, strict):
    from dqn_marl_b200.envs import VecEvacuationEnv
    from oracle import LayoutTables, OracleEnv
    E = len(env_layout)
    env = VecEvacuationEnv(layouts_for_gpu, E, N, device=DEV, seed=seed, env_id_base=40, auto_reset=auto_reset, strict_reference=strict,
                           env_layout=env_layout)
    kw = {} if strict else dict(reset_robots=1, reset_fire=1)
    tabs = [LayoutTables.from_layout(l) for l in host_layouts]
    orcs = [OracleEnv(tabs[env_layout[k]], N, 1, seed=seed, env_id=40 + k, **kw) for k in range(E)]
    obs64 = torch.zeros((E, 1, 11, 11, 6), dtype=torch.float64, device=DEV)
    env.reset(obs64=obs64)
    for o in orcs:
        o._last_obs = o.reset()
    _compare(env, orcs, obs64, "reset")
    rng = np.random.default_rng(seed)
    for t in range(steps):
        acts = rng.integers(0, 6, size=(E, 1)).astype(np.int32)
        _, r, d = env.step(torch.tensor(acts, device=DEV), obs64=obs64)
        r, d = r.cpu().numpy(), d.cpu().numpy()
        for k, o in enumerate(orcs):
            ob, rr, dd = o.step(acts[k])
            o._last_obs = ob
            assert r[k].view(np.uint64) == np.float64(rr).view(np.uint64), f"step {t} env {k}: reward {r[k]!r} {rr!r}"
            assert bool(d[k]) == dd, f"step {t} env {k}: done"
            if dd and auto_reset:
                o._last_obs = o.reset()
        if t % 5 == 4 or t == steps - 1:
            _compare(env, orcs, obs64, f"step {t}")


@pytest.mark.parametrize("N,E,steps", [(150, 37, 60), (20, 12, 260)])
def test_mixed_host_layouts_warp_per_env(N, E, steps):
    """Warp-per-env kernels (28 envs per CTA, cooperative scoring ACROSS envs with different dp5 tables)."""
    lays = _layouts_36x30()
    env_layout = [(7 * k + k // 3) % len(lays) for k in range(E)]
    _run_mixed(lays, lays, env_layout, N, seed=5, steps=steps, auto_reset=(N == 20), strict=(N != 20))


def test_mixed_host_layouts_cta_per_env():
    """One CTA per env (the C3 kernel shape): 96 x 80 galleries generated from different seeds."""
    from dqn_marl_b200.layout import Layout
    lays = [Layout.synthetic(96, 80, n_exits=1 + k, wall_fill=0.10, seed=7 + k) for k in range(3)]
    _run_mixed(lays, lays, [0, 1, 2, 2, 1, 0, 1], 400, seed=9, steps=40, auto_reset=False, strict=True)


def _random_walls(rng, n, L, W):
    from dqn_marl_b200.layout import init_barrier
    walls = np.zeros((n, L + 2, W + 2), dtype=np.uint8)
    walls[:, 0, :] = walls[:, -1, :] = 1
    walls[:, :, 0] = walls[:, :, -1] = 1
    bars, exits = [], []
    for k in range(n):
        bl = [init_barrier((18, 14), (20, 16))]
        for _ in range(int(rng.integers(1, 5))):
            x0, y0 = int(rng.integers(2, L - 5)), int(rng.integers(2, W - 5))
            w, h = int(rng.integers(1, 4)), int(rng.integers(1, 4))
            if x0 <= 31 and x0 + w >= 14:           # keep the robot band free (robots start at (15, 15))
                continue
            bl.append(init_barrier((x0, y0), (x0 + w, y0 + h)))
        for (A, B) in bl:
            walls[k, A[0]:B[0] + 1, A[1]:B[1] + 1] = 1
        ex = [(L, int(rng.integers(2, W - 1))), (1, int(rng.integers(2, W - 1)))][: int(rng.integers(1, 3))]
        for (x, y) in ex:
            walls[k, x, y] = 0                         # exits are walkable cells
        bars.append(bl); exits.append(ex)
    return walls, bars, exits


def test_device_built_layout_batch_drives_the_env():
    """Random walls per layout -> floor fields, dp5 and cellinfo built on the GPU -> env batch.  The device tables equal the
    host Layout's bit for bit, and the trajectories equal the oracle's on the host tables."""
    from dqn_marl_b200.fire import FireSchedule
    from dqn_marl_b200.layout import DeviceLayoutBatch, Layout
    L, W, n = 36, 30, 6
    rng = np.random.default_rng(11)
    walls, bars, exits = _random_walls(rng, n, L, W)
    ex_arr = np.zeros((n, 2, 2), dtype=np.int32)
    n_ex = np.zeros(n, dtype=np.int32)
    host = []
    for k in range(n):
        ex_arr[k, :len(exits[k])] = exits[k]; n_ex[k] = len(exits[k])
        lay = Layout(L=L, W=W, exits=exits[k], barriers=bars[k], obs_exit=exits[k][0])
        lay.fire = FireSchedule([((19.0, 15.0), (2, 2), 0.4)])
        host.append(lay.build())
    batch = DeviceLayoutBatch(L, W, walls, ex_arr, n_ex, device=DEV)
    for k in range(n):
        assert np.array_equal(batch.space[k].cpu().numpy().view(np.uint64), host[k].space.view(np.uint64)), f"layout {k}: space"
        assert np.array_equal(batch.cellinfo[k].cpu().numpy(), host[k].cellinfo), f"layout {k}: cellinfo"
        assert np.array_equal(batch.dp5[k].cpu().numpy().view(np.uint64), host[k].dp5.view(np.uint64)), f"layout {k}: dp5"
    env_layout = [k % n for k in range(20)]
    _run_mixed(batch, host, env_layout, 100, seed=3, steps=50, auto_reset=False, strict=True)


def test_create_layouts_rejects_bad_arguments():
    from dqn_marl_b200 import _lib
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import Layout
    lays = _layouts_36x30()
    with pytest.raises(ValueError):
        VecEvacuationEnv(lays, 4, 10, device=DEV)                                        # no env_layout
    with pytest.raises(ValueError):
        VecEvacuationEnv(lays, 4, 10, device=DEV, env_layout=[0, 1, 2, 9])                # index out of range
    with pytest.raises(ValueError):
        VecEvacuationEnv([lays[0], Layout.reference_room(40, 24, [1, 12])], 2, 10, device=DEV, env_layout=[0, 1])     # other grid size
