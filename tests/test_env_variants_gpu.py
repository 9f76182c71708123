"""The env kernels are compiled in several variants — warps per env (1 / 2 / 4 for small envs, 8 for CTA-per-env, 8 / 16 / 32 for
the global-scratch variant) and envs per CTA (4 / 14 / 28) — of which the library picks one per batch from the env size; the
others stay selectable through `MQ_SMALL_WPE`, `MQ_SMALL_CW`, `MQ_BIG_WPE` (A/B knobs, read at mq_env_create).  Every selectable
variant is held to the same bit-exact comparison with the C oracle; each runs in a subprocess because the knobs are environment
variables."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SMALL = "_run_vs_oracle(Layout.reference_room(n_robots={R}), n_envs={E}, N=150, seed=31, steps=36, check_every=4)"
BIG = "_run_vs_oracle(Layout.synthetic(160, 160, n_exits=4, seed=11), n_envs=2, N=6000, seed=13, steps=8, check_every=2)"


@pytest.mark.parametrize("knobs,call", [
    ({"MQ_SMALL_WPE": "2"}, SMALL.format(R=1, E=9)),
    ({"MQ_SMALL_WPE": "4"}, SMALL.format(R=2, E=7)),
    ({"MQ_SMALL_CW": "4"}, SMALL.format(R=1, E=9)),
    ({"MQ_SMALL_CW": "14"}, SMALL.format(R=2, E=30)),
    ({"MQ_BIG_WPE": "8"}, BIG),
    ({"MQ_BIG_WPE": "16"}, BIG),
], ids=lambda v: "_".join(f"{k}{x}" for k, x in v.items()) if isinstance(v, dict) else None)
def test_selectable_kernel_variant_matches_oracle(knobs, call):
    code = ("import sys; sys.path[:0] = [%r, %r, %r]; from dqn_marl_b200.layout import Layout; from test_env_gpu import _run_vs_oracle; "
            "env = %s; print('variant ok', env.launch_count)" % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), call))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **knobs), capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "variant ok" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]
