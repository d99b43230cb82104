#!/usr/bin/env python
"""A few steps of the reference's Simulation::step() (uniform_disc(n), Barnes-Hut + clamp + boundary + collide, refcompat)
and an order-sensitive checksum of the final state, cells of the last tree included.  Run under the library's A/B
environment switches (NBODY_BH_LOCAL, NBODY_SORT_LAZY, NBODY_BH_CTA_CLIMB, NBODY_COL_STRIP, NBODY_BH_FUSE_INSERT,
NBODY_BHL_GEOM, NBODY_SORT_COOP) by tests/test_gpu_bh.py: every variant must print the same line.
usage: tools/step_checksum.py [n] [steps] [sort_impl]   (sort_impl 1: one kernel per phase, also in the collision pass)"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sort_impl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
b = ic.reference_disc(n)
r = b["radius"]
r[1 : n // 50] = np.minimum(r[1 : n // 50] * 40.0, 300.0)   # some collisions, some bodies on several strips (not the central mass)
with Simulation(b, dt=0.01, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=1.0, eps=1.0, collide=1, sort_impl=sort_impl,
                rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY, use_graph=0) as s:
    s.step(steps)
    out = s.bodies.copy()
    s.attract()
    f8, nxt, depth, leaf = s.bh_nodes()
    resolved = s.collide_stats()
h = hashlib.sha256()
for a in (out["pos"], out["vel"], out["acc"], f8, nxt, depth, leaf.astype(np.uint8)):
    h.update(np.ascontiguousarray(a).tobytes())
print(f"n={n} steps={steps} cells={len(nxt)} sha256={h.hexdigest()}")
