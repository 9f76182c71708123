"""Device floor-field builder against the host builder, which is pinned to the reference's Map.Init_Potential
(map.py:127-148) by tests/test_oracle_golden.py: bit-exact float64 fields for batches of random layouts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _random_layout(rng, L, W, fill, n_exits):
    wall = np.zeros((L + 2, W + 2), dtype=np.uint8)
    wall[0, :] = wall[-1, :] = 1
    wall[:, 0] = wall[:, -1] = 1
    n_rect = int(fill * L * W / 12)
    for _ in range(n_rect):                                   # "gallery" rectangles; may cut pockets off (unreachable cells stay inf)
        x0, y0 = rng.integers(1, L + 1), rng.integers(1, W + 1)
        w, h = rng.integers(1, 6), rng.integers(1, 6)
        wall[x0:min(x0 + w, L + 1), y0:min(y0 + h, W + 1)] = 1
    exits = []
    for k in range(n_exits):
        side = k % 4
        ex, ey = [(L, rng.integers(1, W + 1)), (1, rng.integers(1, W + 1)), (rng.integers(1, L + 1), W), (rng.integers(1, L + 1), 1)][side]
        exits.append((int(ex), int(ey)))
    return wall, exits


@pytest.mark.parametrize("L,W,n,fill,max_exits", [(36, 30, 5, 0.10, 1), (64, 48, 7, 0.20, 3), (200, 256, 3, 0.15, 8), (17, 90, 4, 0.30, 2)])
def test_device_floor_field_equals_host(L, W, n, fill, max_exits):
    from dqn_marl_b200 import _lib
    rng = np.random.default_rng(L * 1000 + W)
    walls, exits_l, adds, refs = [], [], [], []
    ex_arr = np.zeros((n, max_exits, 2), dtype=np.int32)
    n_ex = np.zeros((n,), dtype=np.int32)
    for i in range(n):
        k = int(rng.integers(1, max_exits + 1))
        wall, exits = _random_layout(rng, L, W, fill, k)
        add = rng.random((L + 2, W + 2)) * 200.0                # stands for 200 * danger^2 (map.py:146)
        walls.append(wall); adds.append(add)
        ex_arr[i, :k] = np.asarray(exits, dtype=np.int32); n_ex[i] = k
        refs.append(_lib.floor_field(L, W, wall, np.asarray(exits, dtype=np.int32), add))
    d = "cuda:0"
    out, sweeps = _lib.floor_field_device(L, W, torch.tensor(np.stack(walls), device=d), torch.tensor(ex_arr, device=d), torch.tensor(n_ex, device=d),
                                          torch.tensor(np.stack(adds), device=d))
    got = out.cpu().numpy()
    ref = np.stack(refs)
    assert sweeps > 0
    assert np.array_equal(np.isinf(got), np.isinf(ref)), "reachability differs"
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), "float64 bits differ"
    # and without the additive term
    out2, _ = _lib.floor_field_device(L, W, torch.tensor(np.stack(walls), device=d), torch.tensor(ex_arr, device=d), torch.tensor(n_ex, device=d))
    ref2 = np.stack([_lib.floor_field(L, W, walls[i], ex_arr[i, :n_ex[i]], np.zeros((L + 2, W + 2))) for i in range(n)])
    assert np.array_equal(out2.cpu().numpy().view(np.uint64), ref2.view(np.uint64))
