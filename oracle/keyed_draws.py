"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path.

Keyed random draws (Philox4x32-10) shared by the three implementations of the
Louvre_Evacuation hot path:

  * the unmodified Python reference driven through ``oracle/ref_harness.py``
    (its ``random`` / ``np.random`` module globals are rebound to proxies that
    call the functions below),
  * the plain-C restatement ``oracle/env_oracle.c``,
  * the CUDA kernels in ``dqn_marl_b200/csrc`` (``philox.cuh``).

The reference consumes three interleaved global MT19937 streams whose draw
counts are data dependent, which cannot be replayed in
parallel.  Parity is therefore defined on *keyed* draws: every random decision
of the reference is addressed by (seed, env, tick|episode, person, stream) and
all three implementations evaluate the same pure function for it.

Draw sites in the reference and their keys (counter = (c0, c1, c2, c3)):
  people.py:290  random.uniform(-.1,.1)   c0=env c1=tick    c2=person c3=dir//2      words (2*(dir%2), +1)
  people.py:69-75 np.random.uniform(lo,hi) c0=env c1=tick    c2=person c3=4           words (0,1)
  people.py:239  random.shuffle(movers)    c0=env c1=tick    c2=person c3=4           word 2 = priority (min wins, ties -> lower index)
  people.py:186-190 random.randint spawn   c0=env c1=episode c2=person c3=16+att//2   words (2*(att%2), +1) = (x, y)
  dqn_agent.py:103-104 eps-greedy          c0=env c1=tick    c2=robot  c3=32          words (0,1)=u, word 2=action
  dqn_agent.py:132 random.sample           c0=0   c1=learn_step c2=0    c3=48         4 words = round keys of a Feistel permutation of range(size)
key = (seed & 0xffffffff, seed >> 32).
"""
from __future__ import annotations

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_HEALTH = 4
STREAM_SPAWN = 16
STREAM_AGENT = 32
STREAM_SAMPLE = 48


def philox4x32(c0: int, c1: int, c2: int, c3: int, seed: int):
    """Scalar Philox4x32-10 on Python ints -> 4 uint32 words."""
    k0 = seed & MASK
    k1 = (seed >> 32) & MASK
    c0 &= MASK; c1 &= MASK; c2 &= MASK; c3 &= MASK
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_np(c0, c1, c2, c3, seed: int):
    """Vectorised Philox4x32-10: array-like uint32 counters -> (4, ...) uint32 words."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)])
    c0 = c0.copy(); c1 = c1.copy(); c2 = c2.copy(); c3 = c3.copy()
    k0 = np.uint64(seed & MASK)
    k1 = np.uint64((seed >> 32) & MASK)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for r in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = ((p1 >> s32) ^ c1 ^ k0) & m
        n1 = p1 & m
        n2 = ((p0 >> s32) ^ c3 ^ k1) & m
        n3 = p0 & m
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return np.stack([c0, c1, c2, c3]).astype(np.uint32)


def u53(a: int, b: int) -> float:
    """CPython ``random.random()`` / NumPy legacy ``random_sample`` word->double
    law: ((a>>5)*2^26 + (b>>6)) / 2^53 (exact)."""
    return float(((a >> 5) << 26) | (b >> 6)) * (1.0 / 9007199254740992.0)


def randint32(w: int, lo: int, hi: int) -> int:
    """Inclusive integer in [lo, hi] from one 32-bit word (multiply-high)."""
    return lo + ((w * (hi - lo + 1)) >> 32)


class Draws:
    """Scalar keyed draws for one (seed, env) pair; ``tick``/``episode`` are set by the driver."""

    def __init__(self, seed: int, env: int = 0):
        self.seed = int(seed)
        self.env = int(env)
        self.tick = 0
        self.episode = 0

    def noise_u(self, person: int, dire: int) -> float:
        w = philox4x32(self.env, self.tick, person, dire >> 1, self.seed)
        j = 2 * (dire & 1)
        return u53(w[j], w[j + 1])

    def health_u(self, person: int) -> float:
        w = philox4x32(self.env, self.tick, person, STREAM_HEALTH, self.seed)
        return u53(w[0], w[1])

    def prio(self, person: int) -> int:
        return philox4x32(self.env, self.tick, person, STREAM_HEALTH, self.seed)[2]

    def spawn_xy(self, person: int, attempt: int, L: int, W: int):
        w = philox4x32(self.env, self.episode, person, STREAM_SPAWN + (attempt >> 1), self.seed)
        j = 2 * (attempt & 1)
        return randint32(w[j], 1, L - 2), randint32(w[j + 1], 1, W - 2)

    def agent_u_action(self, robot: int, n_actions: int = 5):
        w = philox4x32(self.env, self.tick, robot, STREAM_AGENT, self.seed)
        return u53(w[0], w[1]), (w[2] * n_actions) >> 32


def _mix32(h: int) -> int:
    h &= MASK
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & MASK
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & MASK
    h ^= h >> 16
    return h


def sample_index(k: int, size: int, rk) -> int:
    """k-th element of a keyed pseudo-random permutation of range(size): a 4-round
    balanced Feistel network on 2*half bits (round keys ``rk`` = the 4 Philox words
    of (0, learn_step, 0, STREAM_SAMPLE)) with cycle-walking back into [0, size).
    Distinct k give distinct indices -> a sample WITHOUT replacement, the law of
    ``random.sample`` at dqn_agent.py:132, computable independently per k."""
    bits = max(2, (size - 1).bit_length())
    half = (bits + 1) >> 1
    mask = (1 << half) - 1
    x = k
    while True:
        l, r = x >> half, x & mask
        for rnd in range(4):
            l, r = r, l ^ (_mix32(r ^ rk[rnd]) & mask)
        x = (l << half) | r
        if x < size:
            return x


def sample_indices(seed: int, learn_step: int, size: int, batch: int) -> np.ndarray:
    rk = philox4x32(0, learn_step, 0, STREAM_SAMPLE, seed)
    return np.array([sample_index(k, size, rk) for k in range(batch)], dtype=np.int64)
