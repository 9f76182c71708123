"""Host-side layout tables (floor field, danger tables, cell info) against tables evaluated with the
reference's own Map / FireSpreadModel code (tests/golden/layout_*.npz)."""
import numpy as np

from util import load_golden


def _np_exp_matches(g):
    return np.array_equal(np.exp(g["exp_probe_in"]), g["exp_probe_out"])


def _check_danger(mine, ref, exact):
    if exact:
        assert np.array_equal(mine, ref)
    else:   # another CPU's np.exp: allow 1 ulp on exp-derived entries
        np.testing.assert_allclose(mine, ref, rtol=2.3e-16, atol=0)


def test_room_tables_bit_exact():
    from dqn_marl_b200.layout import CELL_VALID, Layout
    g = load_golden("layout_room.npz")
    lay = Layout.reference_room()
    assert np.array_equal(lay.space.view(np.uint64), g["space"].view(np.uint64))      # map.py:127-148
    assert np.array_equal(lay.barrier_mask, g["barrier"])                                # map.py:43-57
    valid = (lay.cellinfo & CELL_VALID) != 0
    ref_valid = np.isfinite(g["space"]); ref_valid[[0, -1], :] = False; ref_valid[:, [0, -1]] = False
    assert np.array_equal(valid, ref_valid)
    exact = _np_exp_matches(g)
    pad = int(g["pad"])
    L, W = lay.L, lay.W
    x0, y0, w, h = lay.ctr_box
    ix0, iy0, iw, ih = lay.int_box
    for k, s in enumerate(g["steps"]):
        full = np.zeros((L + 2, W + 2)); full[x0:x0 + w, y0:y0 + h] = lay.danger_ctr[s]
        _check_danger(full, g["danger_ctr"][k], exact)                                  # people.py:205
        ref_int = g["danger_int"][k]
        mine = np.zeros_like(ref_int)
        sx0, sx1 = max(ix0, -pad), min(ix0 + iw, L + 2 + pad)
        sy0, sy1 = max(iy0, -pad), min(iy0 + ih, W + 2 + pad)
        mine[sx0 + pad:sx1 + pad, sy0 + pad:sy1 + pad] = lay.danger_int[s, sx0 - ix0:sx1 - ix0, sy0 - iy0:sy1 - iy0]
        _check_danger(mine, ref_int, exact)                                             # evacuation_env.py:106


def test_big256_tables():
    from dqn_marl_b200.layout import Layout
    g = load_golden("layout_big256.npz")
    lay = Layout.reference_room(256, 256, [256, 128])
    assert np.array_equal(lay.space.view(np.uint64), g["space"].view(np.uint64))
    exact = _np_exp_matches(g)
    bx0, by0, bx1, by1 = g["box"]
    x0, y0, w, h = lay.ctr_box
    for k, s in enumerate(g["steps"]):
        ref = g["danger_ctr"][k]                 # cells [0,bx1) x [0,by1)
        mine = np.zeros_like(ref)
        ex, ey = min(x0 + w, bx1), min(y0 + h, by1)
        mine[x0:ex, y0:ey] = lay.danger_ctr[s, :ex - x0, :ey - y0]
        _check_danger(mine, ref, exact)


def _check_synth(layout_name, traj_name, boxed):
    """Multi-exit / multi-barrier tables against the reference's own Map(L, W, exits, barriers) (map.py:38-79,127-148) and its
    fire models evaluated on that map (oracle/make_golden.py main_synthetic)."""
    from dqn_marl_b200.layout import CELL_VALID
    from util import layout_for
    g = load_golden(layout_name)
    lay = layout_for(load_golden(traj_name)["meta"])
    assert np.array_equal(lay.space.view(np.uint64), g["space"].view(np.uint64))          # multi-source Dijkstra + fire term
    assert np.array_equal(lay.barrier_mask, g["barrier"])
    valid = (lay.cellinfo & CELL_VALID) != 0
    ref_valid = np.isfinite(g["space"]); ref_valid[[0, -1], :] = False; ref_valid[:, [0, -1]] = False
    assert np.array_equal(valid, ref_valid)
    exact = _np_exp_matches(g)
    pad = int(g["pad"])
    L, W = lay.L, lay.W
    x0, y0, w, h = lay.ctr_box
    ix0, iy0, iw, ih = lay.int_box
    for k, s in enumerate(g["steps"]):
        full = np.zeros((L + 2, W + 2)); full[x0:x0 + w, y0:y0 + h] = lay.danger_ctr[s]
        full_int = np.zeros((L + 2 + 2 * pad, W + 2 + 2 * pad))
        sx0, sx1 = max(ix0, -pad), min(ix0 + iw, L + 2 + pad)
        sy0, sy1 = max(iy0, -pad), min(iy0 + ih, W + 2 + pad)
        full_int[sx0 + pad:sx1 + pad, sy0 + pad:sy1 + pad] = lay.danger_int[s, sx0 - ix0:sx1 - ix0, sy0 - iy0:sy1 - iy0]
        if boxed:
            bx0, by0, bx1, by1 = g["box"]
            full = full[max(0, bx0):bx1, max(0, by0):by1]
            full_int = full_int[bx0 + pad:bx1 + pad, by0 + pad:by1 + pad]
        _check_danger(full, g["danger_ctr"][k], exact)
        _check_danger(full_int, g["danger_int"][k], exact)


def test_synthetic_gallery_tables_match_reference_map():
    _check_synth("layout_synth_gallery.npz", "traj_synth_gallery.npz", boxed=True)


def test_multi_fire_hall_tables_match_reference_map():
    _check_synth("layout_synth_hall.npz", "traj_synth_hall.npz", boxed=False)


def test_cpp_floor_field_equals_python_restatement():
    """The product's host builder (csrc/floor_field.cpp via mq_floor_field) against the heapq restatement of map.py:127-148
    in oracle/floor_field_py.py, on layouts the goldens do not cover (many exits, 20 % walls, unreachable pockets)."""
    from floor_field_py import floor_field
    from dqn_marl_b200.layout import Layout
    for (L, W, ne, fill, seed) in [(128, 96, 5, 0.20, 3), (64, 200, 8, 0.15, 9), (256, 256, 1, 0.10, 2024)]:
        a = Layout.synthetic(L, W, n_exits=ne, wall_fill=fill, seed=seed)
        b = Layout.synthetic(L, W, n_exits=ne, wall_fill=fill, seed=seed, floor_field=floor_field)
        assert np.array_equal(a.space.view(np.uint64), b.space.view(np.uint64)), (L, W)
        assert np.array_equal(a.dp5.view(np.uint64), b.dp5.view(np.uint64)) and np.array_equal(a.cellinfo, b.cellinfo)


def test_dp5_and_cellinfo_semantics():
    """dp5 = (space[c]-space[n])*5.0 for valid pairs, -inf otherwise; exit neighbourhood evacuates."""
    from dqn_marl_b200.layout import CELL_EVACUATES, CELL_OBS_EXIT, MOVE_TO, Layout
    lay = Layout.reference_room()
    sp = lay.space
    rng = np.random.default_rng(1)
    for _ in range(500):
        x, y, d = int(rng.integers(0, lay.L + 2)), int(rng.integers(0, lay.W + 2)), int(rng.integers(0, 8))
        nx, ny = x + MOVE_TO[d][0], y + MOVE_TO[d][1]
        ok = (1 <= x <= lay.L and 1 <= y <= lay.W and 1 <= nx <= lay.L and 1 <= ny <= lay.W
              and np.isfinite(sp[x, y]) and np.isfinite(sp[nx, ny]))
        if ok:
            assert lay.dp5[x, y, d] == (sp[x, y] - sp[nx, ny]) * 5.0
        else:
            assert lay.dp5[x, y, d] == -np.inf
    assert lay.cellinfo[36, 15] & CELL_OBS_EXIT
    ev = (lay.cellinfo & CELL_EVACUATES) != 0
    assert ev[35:38, 14:17].all() and ev.sum() == 9       # map.py:109-112, quirk Q10


def test_synthetic_layout_is_usable():
    from dqn_marl_b200.layout import CELL_VALID, Layout
    lay = Layout.synthetic(128, 128, n_exits=3, seed=7)
    valid = (lay.cellinfo & CELL_VALID) != 0
    assert 0.80 < valid[1:-1, 1:-1].mean() < 0.95
    assert valid[15:31, 1:129].all() or valid[15:31, 1:129].mean() > 0.97   # robot band kept free of galleries
