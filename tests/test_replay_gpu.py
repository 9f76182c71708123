"""Replay ring (csrc/replay.cu) through the C-ABI vs the numpy oracle of the reference's deque replay."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fill(ring, orc, rng, n, t0):
    s = rng.standard_normal((n, 11, 11, 6))
    ns = rng.standard_normal((n, 11, 11, 6))
    a = rng.integers(0, 5, n)
    r = rng.standard_normal(n) * 100
    d = rng.integers(0, 2, n).astype(bool)
    dev = ring.device
    ring.push(torch.tensor(s, dtype=torch.float32, device=dev), torch.tensor(a, dtype=torch.int32, device=dev),
              torch.tensor(r, dtype=torch.float64, device=dev), torch.tensor(ns, dtype=torch.float32, device=dev),
              torch.tensor(d, dtype=torch.uint8, device=dev))
    for k in range(n):
        orc.remember(s[k].astype(np.float32), int(a[k]), float(r[k]), ns[k].astype(np.float32), bool(d[k]))


def _check(out, ref):
    for k in ("states", "next_states", "rewards", "actions", "dones"):
        assert np.array_equal(out[k].cpu().numpy().reshape(ref[k].shape), ref[k]), k


@pytest.mark.parametrize("cap,pushes", [(64, [10, 20, 34]), (100, [37, 37, 37, 37, 100, 1]), (4096, [1500, 1500, 1500, 7])])
def test_ring_fifo_and_sampling(cap, pushes):
    from dqn_marl_b200.replay import ReplayRing
    from replay_oracle import ReplayOracle
    ring = ReplayRing(cap, device="cuda:0", seed=77)
    orc = ReplayOracle(cap, seed=77)
    rng = np.random.default_rng(cap)
    for step, n in enumerate(pushes):
        _fill(ring, orc, rng, n, step)
        assert len(ring) == len(orc)
        size = len(orc)
        # injected indices (replaying what the reference's random.sample picked), with duplicates allowed here
        inj = rng.integers(0, size, size=min(size, 48))
        out = ring.sample(len(inj), inject_idx=torch.tensor(inj))
        _check(out, orc.sample(len(inj), indices=inj))
        # keyed permutation: same indices as the oracle law, all distinct (without replacement)
        B = min(size, 32)
        out = ring.sample(B, draw_id=step, want_idx=True)
        ref = orc.sample(B, draw_id=step)
        assert np.array_equal(out["idx"].cpu().numpy(), ref["idx"]) and len(set(ref["idx"].tolist())) == B
        _check(out, ref)
    # the whole population: a permutation
    out = ring.sample(len(orc), draw_id=99, want_idx=True)
    assert sorted(out["idx"].cpu().tolist()) == list(range(len(orc)))


def test_sample_larger_than_population_raises():
    from dqn_marl_b200 import _lib
    from dqn_marl_b200.replay import ReplayRing
    ring = ReplayRing(16, device="cuda:0")
    with pytest.raises(_lib.MqError):
        ring.sample(4)


def test_full_size_roundtrip_properties():
    """C3-sized push (16384 transitions per step into a 2^20 ring) and a 65536 batch (C5): the sample is a
    set of distinct rows whose payload equals the pushed payload (checksum of checksums)."""
    from dqn_marl_b200.replay import ReplayRing
    dev = torch.device("cuda:0")
    cap, n = 1 << 18, 16384
    ring = ReplayRing(cap, device=dev, seed=5)
    g = torch.Generator(device=dev); g.manual_seed(1)
    for t in range(20):     # wraps the ring once
        tag = torch.arange(t * n, (t + 1) * n, device=dev, dtype=torch.float32)
        s = tag[:, None].expand(n, 726).contiguous()
        ring.push(s, (tag.to(torch.int32) % 5), tag.double() * 0.5, (s + 0.25).contiguous(), (tag.to(torch.int64) % 2).to(torch.uint8))
    assert len(ring) == cap
    out = ring.sample(65536, want_idx=True)
    idx = out["idx"]
    assert idx.unique().numel() == 65536
    tag = out["states"].reshape(65536, 726)[:, 0]
    assert (out["states"].reshape(65536, 726) == tag[:, None]).all()
    assert (out["next_states"].reshape(65536, 726) == tag[:, None] + 0.25).all()
    assert (out["rewards"] == tag * 0.5).all() and (out["actions"] == tag.long() % 5).all()
    oldest_tag = 20 * n - cap
    assert (tag == (oldest_tag + idx).float()).all()
