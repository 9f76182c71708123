"""Host-side fire schedule -> danger tables.

The reference evaluates ``ProgressiveFireModel.get_max_danger(pos)`` (reference
Louvre_Evacuation/envs/fire_model.py:143-188) N + 121 times per env step.  It is
a pure function of (position, fire step in 0..max_steps), so this build
tabulates it once per layout on the host and the CUDA kernels only index the
table.  The arithmetic below follows fire_model.py operation by operation in
float64 (numpy ufuncs, the same ``np.sqrt`` / ``np.exp`` the reference calls) so
that the tables are bit-identical to what the reference computes on the same
machine (tests/test_layout.py pins this against tests/golden/).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

# fire_model.py:45-51 — sources that appear later, hard-coded in the reference
DEFAULT_ADDITIONAL = (
    ((25, 20), (5, 5), 0.8),
    ((13, 18), (5, 5), 0.8),
    ((20, 10), (4, 4), 0.7),
    ((14, 22), (4, 4), 0.6),
    ((26, 14), (3, 3), 0.5),
)


@dataclass
class Source:
    center: Tuple[float, float]
    size: Tuple[int, int]
    intensity: float


class FireSchedule:
    """Mirror of ProgressiveFireModel's time interpolation (fire_model.py:7-141).

    ``initial``: [(center, size, intensity)] — the reference builds one per barrier
    with size (2, 2) and intensity 0.4 (map.py:58-65, fire_model.py:22-28).
    """

    def __init__(self, initial: Sequence, additional: Sequence = DEFAULT_ADDITIONAL, max_steps: int = 180):
        self.max_steps = int(max_steps)
        self.initial = [Source(tuple(c), tuple(s), float(i)) for (c, s, i) in initial]
        # fire_model.py:35-42
        self.final = [Source(s.center, (min(8, s.size[0] * 4), min(8, s.size[1] * 4)), min(1.0, s.intensity + 0.4))
                      for s in self.initial]
        self.additional = [Source(tuple(c), tuple(s), float(i)) for (c, s, i) in additional]
        self.base_radius = 5.0      # fire_model.py:55
        self.max_radius = 20.0      # fire_model.py:56
        self.min_danger = 0.05      # fire_model.py:57

    # fire_model.py:69-136
    def sources_at(self, step: int) -> List[Source]:
        finals = self.final + self.additional
        if step >= self.max_steps:
            return list(finals)
        progress = step / self.max_steps
        if progress < 0.2:
            rate = progress * 2
        elif progress < 0.5:
            rate = 0.5 + (progress - 0.2) * 1
        elif progress < 0.8:
            rate = 1.0 + (progress - 0.5) * 0.8
        else:
            rate = 1.3 + (progress - 0.8) * 0.5
        rate = min(rate, 1.0)
        out = []
        for ini, fin in zip(self.initial, self.final):
            sx = ini.size[0] + (fin.size[0] - ini.size[0]) * rate
            sy = ini.size[1] + (fin.size[1] - ini.size[1]) * rate
            inten = ini.intensity + (fin.intensity - ini.intensity) * rate
            out.append(Source(ini.center, (int(sx), int(sy)), inten))
        for i, fin in enumerate(self.additional):
            thr = 0.3 + (i * 0.15)
            if progress >= thr:
                npg = (progress - thr) / (1.0 - thr)
                npg = min(npg, 1.0)
                sx = 1 + (fin.size[0] - 1) * npg
                sy = 1 + (fin.size[1] - 1) * npg
                inten = 0.2 + (fin.intensity - 0.2) * npg
                out.append(Source(fin.center, (int(sx), int(sy)), inten))
        return out

    # fire_model.py:138-141
    def radius_at(self, step: int) -> float:
        progress = min(step / self.max_steps, 1.0)
        return self.base_radius + (self.max_radius - self.base_radius) * progress

    # fire_model.py:143-188, vectorised over positions
    def danger_field(self, step: int, px: np.ndarray, py: np.ndarray) -> np.ndarray:
        step = min(int(step), self.max_steps)
        px = np.asarray(px, dtype=np.float64)
        py = np.asarray(py, dtype=np.float64)
        out = np.zeros(np.broadcast(px, py).shape, dtype=np.float64)
        radius = self.radius_at(step)
        for s in self.sources_at(step):
            dist = np.sqrt((px - s.center[0]) ** 2 + (py - s.center[1]) ** 2)
            core = max(s.size[0], s.size[1]) / 2.0
            base = np.where(dist <= radius * 0.3, s.intensity * 1.0,
                   np.where(dist <= radius * 0.5, s.intensity * 0.8,
                   np.where(dist <= radius * 0.7, s.intensity * 0.6, s.intensity * 0.4)))
            decay = np.exp(-(dist - core) / 6.0)
            d = np.maximum(base * decay, self.min_danger)
            d = np.where(dist <= radius, d, 0.0)
            d = np.where(dist <= core, s.intensity, d)
            out = np.maximum(out, d)
        return out

    def bounding_box(self, margin: float = 0.0):
        """Integer box [x0, x1) x [y0, y1) outside which danger is 0 at every step."""
        srcs = self.final + self.additional
        r = self.max_radius + margin
        x0 = int(np.floor(min(s.center[0] for s in srcs) - r)) - 1
        x1 = int(np.ceil(max(s.center[0] for s in srcs) + r)) + 2
        y0 = int(np.floor(min(s.center[1] for s in srcs) - r)) - 1
        y1 = int(np.ceil(max(s.center[1] for s in srcs) + r)) + 2
        return x0, y0, x1, y1
