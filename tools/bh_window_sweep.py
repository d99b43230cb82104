#!/usr/bin/env python
"""Warp-cooperative Barnes-Hut walk: force time against the lane window (NBODY_BH_WALK_WINDOW), and a bit-compare
of the accelerations against window = 1.  usage: tools/bh_window_sweep.py [windows...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402

windows = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 256, 1 << 20]
n = 1 << 20
cases = {
    "octree_theta0.5": (ic.plummer(n, seed=3, dims=3), dict(dt=1e-3, eps=0.01, dims=3, theta=0.5, bh_fix_near_leaves=1)),
    "octree_theta0.3": (ic.plummer(n, seed=3, dims=3), dict(dt=1e-3, eps=0.01, dims=3, theta=0.3, bh_fix_near_leaves=1)),
    "quadtree_theta1.0": (ic.spinning_disc(n, seed=3, scale=3200.0, spin=0.01), dict(dt=0.01, eps=1.0, dims=2, theta=1.0)),
    "quadtree_theta0.5": (ic.spinning_disc(n, seed=3, scale=3200.0, spin=0.01), dict(dt=0.01, eps=1.0, dims=2, theta=0.5)),
}
for name, (b, kw) in cases.items():
    ref = None
    row = {"case": name, "n": n}
    for w in windows:
        os.environ["NBODY_BH_WALK_WINDOW"] = str(w)
        with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, bh_walk=2, use_graph=0, **kw) as s:
            s.attract()
            acc = s.download()["acc"].copy().view(np.uint32)
            if ref is None:
                ref = acc
            same = bool(np.array_equal(ref, acc))
            s.step(2)
            f = []
            for _ in range(5):
                s.profile_next_step(True)
                s.step(1)
                f.append(s.info()["last_force_ms"])
        row[f"w{w}"] = round(float(np.median(f)), 3)
        if not same:
            row[f"w{w}_DIFFERS"] = True
    # the per-thread walk for comparison
    os.environ.pop("NBODY_BH_WALK_WINDOW", None)
    with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, bh_walk=1, use_graph=0, **kw) as s:
        s.step(2)
        f = []
        for _ in range(5):
            s.profile_next_step(True)
            s.step(1)
            f.append(s.info()["last_force_ms"])
    row["per_thread_walk"] = round(float(np.median(f)), 3)
    print(json.dumps(row), flush=True)
