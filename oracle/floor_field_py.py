"""ORACLE / TEST INFRASTRUCTURE ONLY — pure-Python restatement of the reference's static floor field.

Map.Init_Potential (reference Louvre_Evacuation/envs/map.py:127-148): multi-source Dijkstra over the 8 MoveTO directions
(map.py:11-19), cost 1.0 for the four axis moves and 1.4 for the diagonals (:137), every exit seeded with 1 (:130-132),
neighbours admitted by Check_Valid (map.py:85-92: inside 1..L x 1..W and not a wall), then `+= add_term` on reached cells
(:143-147, the 200*danger^2 fire term).  Same heap discipline as the reference (tuples (dist, x, y) in a heapq), so equal
floating-point sums are produced in the same order.

Used (1) by tests to cross-check the product's C++ builder (dqn_marl_b200/csrc/floor_field.cpp) independently of the
goldens, and (2) by `bench.py --impl reference`, whose process must not load the product's library.
Signature matches dqn_marl_b200._lib.floor_field so it can be passed as `Layout.build(floor_field=...)`.
"""
import heapq

import numpy as np

MOVE_TO = ((1, 0), (0, -1), (-1, 0), (0, 1), (1, -1), (-1, -1), (-1, 1), (1, 1))      # map.py:11-19


def floor_field(L, W, wall, exits, add_term):
    wall = np.asarray(wall)
    inf = float("inf")
    blocked = [[bool(wall[x, y]) for y in range(W + 2)] for x in range(L + 2)]
    dist = [[inf] * (W + 2) for _ in range(L + 2)]
    heap = []
    for (ex, ey) in np.asarray(exits).reshape(-1, 2).tolist():
        dist[ex][ey] = 1
        heapq.heappush(heap, (1, ex, ey))
    while heap:
        d, x, y = heapq.heappop(heap)
        for i, (dx, dy) in enumerate(MOVE_TO):
            nx, ny = x + dx, y + dy
            if nx >= L + 1 or nx <= 0 or ny >= W + 1 or ny <= 0 or blocked[nx][ny]:      # Check_Valid
                continue
            nd = d + (1.0 if i < 4 else 1.4)
            if nd < dist[nx][ny]:
                dist[nx][ny] = nd
                heapq.heappush(heap, (nd, nx, ny))
    out = np.array(dist, dtype=np.float64)
    add = np.asarray(add_term, dtype=np.float64)
    reached = np.isfinite(out)
    out[reached] = out[reached] + add[reached]
    return out
