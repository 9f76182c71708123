"""Randomised parity sweep: env kernels against the C oracle on layouts / populations / seeds drawn at random (fixed
generator), beyond the hand-picked cases of test_env_gpu.py.  Every kernel variant is crossed: warp-per-env (N <= 256),
CTA-per-env (N <= ~5000 people in shared memory) and the global-scratch variant, one and two robots, strict and fresh-run
reset policies, invalid actions (5 is ignored, map.py:180-181).  Bit-exact on positions, health / accumulator bits, occupancy,
f64 observations, rewards and dones at every checked step."""
import numpy as np
import pytest

from test_env_gpu import _run_vs_oracle

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20261019)
    out = []
    for i in range(14):
        kind = ("small", "small", "cta", "cta", "big")[i % 5]
        if kind == "small":
            L, W = int(rng.integers(34, 90)), int(rng.integers(24, 70))
            N, envs, steps = int(rng.integers(1, 257)), int(rng.integers(1, 40)), 70
        elif kind == "cta":
            L, W = int(rng.integers(60, 200)), int(rng.integers(40, 200))
            N, envs, steps = int(rng.integers(257, 2500)), int(rng.integers(1, 6)), 40
        else:
            L, W = int(rng.integers(150, 300)), int(rng.integers(150, 300))
            N, envs, steps = int(rng.integers(5200, 9000)), 2, 10
        out.append(dict(L=L, W=W, N=N, envs=envs, steps=steps, exits=int(rng.integers(1, 9)), fill=float(rng.uniform(0.0, 0.16)),
                        lseed=int(rng.integers(1, 10000)), seed=int(rng.integers(1, 10000)), robots=1 + int(rng.integers(0, 2)),
                        strict=bool(rng.integers(0, 2)), auto=bool(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("c", _cases(), ids=lambda c: f"{c['L']}x{c['W']}_N{c['N']}_E{c['envs']}_R{c['robots']}")
def test_random_configuration_matches_oracle(c):
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(c["L"], c["W"], n_exits=c["exits"], wall_fill=c["fill"], seed=c["lseed"], n_robots=c["robots"])
    _run_vs_oracle(lay, n_envs=c["envs"], N=c["N"], seed=c["seed"], steps=c["steps"], n_act=6, check_every=3,
                   strict=c["strict"], auto_reset=c["auto"])
