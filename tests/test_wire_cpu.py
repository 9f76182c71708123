"""Host half of the compact observation wire format (mq_obs_wire_expand, include/marl_b200.h) without a GPU: records packed
with numpy expand to the dense 11x11x6 windows bit for bit, for any thread count."""
import numpy as np


def test_wire_expand_matches_numpy_packing():
    from dqn_marl_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    n = 5000
    obs = np.zeros((n, 121, 6), np.float32)
    obs[:, :, 1] = rng.random((n, 121)) < 0.3
    obs[:, :, 2] = rng.random((n, 121)).astype(np.float32) * (rng.random((n, 121)) < 0.5)
    obs[:, :, 3] = rng.random((n, 121)) < 0.2
    obs[:, :, 4] = rng.random((n, 121)) < 0.01
    obs[:, 60, 5] = 1                                               # evacuation_env.py:116-117
    wire = np.zeros((n, _lib.MQ_OBS_WIRE_WORDS), np.uint32)
    wire[:, :121] = obs[:, :, 2].view(np.uint32)
    for ch, base in ((1, 121), (3, 125), (4, 129)):
        bits = obs[:, :, ch].astype(np.uint32)
        for c in range(121):
            wire[:, base + (c >> 5)] |= bits[:, c] << np.uint32(c & 31)
    out = np.empty((n, 121, 6), np.float32)
    for threads in (1, 0, 4, 64):
        out[:] = 7
        _lib.check(lib.mq_obs_wire_expand(_lib.ptr(wire), n, _lib.ptr(out), threads), "mq_obs_wire_expand")
        assert np.array_equal(out.view(np.uint32), obs.view(np.uint32)), threads
    # ragged counts and a destination that starts 8 bytes off a 16-byte boundary (the streaming-store path has a head and a
    # tail), repeated calls on the persistent worker pool; the floats around the destination stay untouched
    buf = np.full((n * 726 + 32,), 7, np.float32)
    for off in (0, 2, 1, 5, 14):                                   # 16-byte / 8-byte / odd-float phases of the 64-byte streaming stores
        for m in (1, 7, 8, 9, 129, 1031, n):
            for threads in (1, 3, 16):
                buf[:] = 7
                dst = buf[off:off + m * 726]
                _lib.check(lib.mq_obs_wire_expand(_lib.ptr(wire), m, dst.ctypes.data, threads), "mq_obs_wire_expand")
                assert np.array_equal(dst.view(np.uint32), obs[:m].reshape(-1).view(np.uint32)), (off, m, threads)
                assert (buf[:off] == 7).all() and (buf[off + m * 726:] == 7).all(), (off, m, threads)
    _lib.check(lib.mq_obs_wire_expand(_lib.ptr(wire), 0, _lib.ptr(out), 1), "empty input")


def test_wire_expand_baseline_isa_in_a_subprocess():
    """On a CPU with AVX-512 the library dispatches to the AVX-512 form; MQ_WIRE_ISA=sse2 (read once per process) forces the
    baseline SSE2 form so that both are held to the same bit-exact check on the same machine."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_wire_cpu as t; "
            "t.test_wire_expand_matches_numpy_packing(); print('ok')" % (root, os.path.join(root, "tests")))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MQ_WIRE_ISA="sse2"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
