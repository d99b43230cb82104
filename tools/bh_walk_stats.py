#!/usr/bin/env python
"""Distribution of the walk length (node visits per target) of a Barnes-Hut scene: the tree is downloaded from the library
and every target's walk is replayed with numpy (the opening test in fp32, as in Quadtree::acc).  The per-thread walk kernel's
time is the LONGEST walk x the latency of one dependent visit, so the maximum and the tail matter, not the mean.
usage: tools/bh_walk_stats.py [n] [theta]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
theta = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
b = ic.reference_disc(n)
with Simulation(b, dt=0.01, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=theta, eps=1.0, rsqrt_mode=capi.RSQRT_REFCOMPAT,
                integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY, collide=1) as s:
    s.step(8)
    s.attract()
    f8, nxt, _depth, leaf = s.bh_nodes()
    pos = s.bodies["pos"][:, :2].astype(np.float32)
x, y, size = f8[:, 0], f8[:, 1], f8[:, 7]
s2 = (size * size).astype(np.float32)
nxt = nxt.astype(np.int64)
t2 = np.float32(theta * theta)
cur = np.zeros(n, dtype=np.int64)
alive = np.ones(n, dtype=bool)
visits = np.zeros(n, dtype=np.int64)
jumps = np.zeros(n, dtype=np.int64)          # visits whose successor is not the adjacent record
while alive.any():
    a = np.nonzero(alive)[0]
    c = cur[a]
    dx = x[c] - pos[a, 0]
    dy = y[c] - pos[a, 1]
    far = s2[c] < (dx * dx + dy * dy).astype(np.float32) * t2
    skip = far | leaf[c]
    nn = np.where(skip, nxt[c], c + 1)
    visits[a] += 1
    jumps[a] += (nn != c + 1) & (nn != 0)
    done = nn == 0
    cur[a] = nn
    alive[a[done]] = False
pct = [50, 90, 99, 99.9, 100]
print(json.dumps({"n": n, "theta": theta, "nodes": int(len(x)), "visits_mean": float(visits.mean()),
                  "visits_pct": dict(zip(map(str, pct), np.percentile(visits, pct).tolist())),
                  "jumps_mean": float(jumps.mean()), "jumps_of_longest": int(jumps[visits.argmax()]),
                  "longest_per_warp_mean": float(visits[: n // 32 * 32].reshape(-1, 32).max(1).mean())}))
