"""Shared helpers for the parity tests (golden loading, layouts from golden meta)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
OP_STEP, OP_RESET = 0, 1

_layout_cache = {}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def layout_from_spec(L, W, spec):
    """Product Layout of a multi-exit / multi-barrier geometry recorded in a golden's meta (oracle/make_golden.py
    synthetic_specs): the reference's Map(L, W, exits, barriers) burns every barrier (map.py:58-65) unless fire_first_only."""
    from dqn_marl_b200.fire import FireSchedule
    from dqn_marl_b200.layout import Layout, init_barrier
    exits = [tuple(e) for e in spec["exits"]]
    bars = [init_barrier(tuple(A), tuple(B)) for (A, B) in spec["barriers"]]
    lay = Layout(L=L, W=W, exits=exits, barriers=bars, obs_exit=exits[0])
    if spec.get("fire_first_only"):
        (A, B) = bars[0]
        lay.fire = FireSchedule([(((A[0] + B[0]) / 2, (A[1] + B[1]) / 2), (2, 2), 0.4)])
    return lay.build()


def layout_for(meta):
    from dqn_marl_b200.layout import Layout
    key = (meta["width"], meta["height"], tuple(meta["exit"]), meta["n_robots"], json.dumps(meta.get("layout"), sort_keys=True))
    if key not in _layout_cache:
        if meta.get("layout"):
            _layout_cache[key] = layout_from_spec(meta["width"], meta["height"], meta["layout"])
        else:
            _layout_cache[key] = Layout.reference_room(meta["width"], meta["height"], meta["exit"], n_robots=meta["n_robots"])
    return _layout_cache[key]


def unpack_rmap(bits, L, W):
    return np.unpackbits(bits)[: (L + 2) * (W + 2)].reshape(L + 2, W + 2)


STATE_KEYS = ("px", "py", "health", "acc", "flags", "robots", "fire_step", "cur_step")


def assert_frame_equal(gold, f, snap, obs, reward, done, L, W, tag=""):
    """Bit-exact comparison of one frame of a golden trajectory with an implementation's snapshot."""
    for k in STATE_KEYS:
        a = np.asarray(gold[k][f])
        b = np.asarray(snap[k]).astype(a.dtype).reshape(a.shape)
        assert np.array_equal(a.view(np.uint8) if a.dtype.kind == "f" else a,
                              b.view(np.uint8) if b.dtype.kind == "f" else b), f"{tag} frame {f}: {k} differs"
    assert np.array_equal(unpack_rmap(gold["rmap"][f], L, W), np.asarray(snap["rmap"]).reshape(L + 2, W + 2)), \
        f"{tag} frame {f}: rmap differs"
    if obs is not None:
        g = gold["obs"][f]
        assert np.array_equal(g.view(np.uint64), np.asarray(obs, dtype=np.float64).reshape(g.shape).view(np.uint64)), \
            f"{tag} frame {f}: obs differs"
    if gold["op"][f] == OP_STEP and reward is not None:
        assert np.float64(reward).view(np.uint64) == gold["reward"][f].view(np.uint64), \
            f"{tag} frame {f}: reward {reward!r} != {gold['reward'][f]!r}"
        assert bool(done) == bool(gold["done"][f]), f"{tag} frame {f}: done differs"
