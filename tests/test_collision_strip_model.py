"""CPU: the collision grid's x strips (csrc/collide.cuh) take every sweep pair exactly once per shared cell.

The reference pairs two bodies of a 600-unit cell when their x intervals overlap (Simulation.hpp:216-290, the sweep inside a
cell), once per cell they share.  The GPU grid splits a cell's list by strips of COL_STRIP units: a body is entered under
(cell, strip) for every strip its x interval touches, and a pair is taken only in the strip in which the overlap STARTS.
This model replays both rules in fp32 on random scenes (radii from far below to far above the strip width, negative
coordinates, bodies on cell borders) and compares the multisets of (first, second, cell)."""
from collections import Counter

import numpy as np

CELL = np.float32(600.0)
STRIP = np.float32(37.5)


def cell_range(x, y, r):
    f = np.float32
    return (int(f(f(x - r) / CELL)), int(f(f(x + r) / CELL)), int(f(f(y - r) / CELL)), int(f(f(y + r) / CELL)))   # truncation toward zero


def strip_of(v):
    return int(np.floor(np.float32(np.float32(v) / STRIP)))


def sweep_pair(ax, ar, ia, bx, br, ib):
    f = np.float32
    amin, amax, bmin, bmax = f(ax - ar), f(ax + ar), f(bx - br), f(bx + br)
    if max(amin, bmin) > min(amax, bmax):
        return None
    a_first = (amin < bmin) or (amin == bmin and ia < ib)
    return (ia, ib) if a_first else (ib, ia)


def reference_pairs(pos, rad):
    cells = {}
    for i, ((x, y), r) in enumerate(zip(pos, rad)):
        x0, x1, y0, y1 = cell_range(x, y, r)
        for cy in range(y0, y1 + 1):
            for cx in range(x0, x1 + 1):
                cells.setdefault((cx, cy), []).append(i)
    out = Counter()
    for c, members in cells.items():
        for a in range(len(members)):
            for b in range(a + 1, len(members)):
                i, j = members[a], members[b]
                p = sweep_pair(pos[i][0], rad[i], i, pos[j][0], rad[j], j)
                if p:
                    out[(p, c)] += 1
    return out


def strip_pairs(pos, rad):
    f = np.float32
    units = {}
    for i, ((x, y), r) in enumerate(zip(pos, rad)):
        x0, x1, y0, y1 = cell_range(x, y, r)
        s0, s1 = strip_of(f(x - r)), strip_of(f(x + r))
        for cy in range(y0, y1 + 1):
            for cx in range(x0, x1 + 1):
                for st in range(s0, s1 + 1):
                    units.setdefault((cx, cy, st), []).append(i)
    out = Counter()
    for (cx, cy, st), members in units.items():
        for i in members:                      # body i lists the pairs of this unit in which it is `first`
            for j in members:
                if j == i:
                    continue
                p = sweep_pair(pos[i][0], rad[i], i, pos[j][0], rad[j], j)
                if not p or p[0] != i:
                    continue
                if strip_of(max(f(pos[i][0] - rad[i]), f(pos[j][0] - rad[j]))) != st:
                    continue                   # taken in the strip where the overlap starts
                out[(p, (cx, cy))] += 1
    return out


def test_strips_reproduce_the_sweep_pairs_per_cell():
    rng = np.random.default_rng(7)
    for trial in range(12):
        n = int(rng.integers(50, 400))
        span = float(rng.choice([300.0, 1500.0, 5000.0]))
        pos = rng.uniform(-span, span, (n, 2)).astype(np.float32)
        rad = rng.choice([0.5, 5.0, 30.0, 120.0, 400.0], n).astype(np.float32) * rng.uniform(0.5, 1.5, n).astype(np.float32)
        if trial % 3 == 0:                     # bodies exactly on cell and strip borders, coincident intervals
            pos[: n // 4, 0] = np.round(pos[: n // 4, 0] / 37.5) * 37.5
            pos[n // 4: n // 2, 0] = np.round(pos[n // 4: n // 2, 0] / 600.0) * 600.0
            rad[: n // 8] = 37.5
        want, got = reference_pairs(pos, rad), strip_pairs(pos, rad)
        assert want == got, (trial, len(want), len(got), list((want - got).items())[:3], list((got - want).items())[:3])
        assert sum(want.values()) > 0
