#!/usr/bin/env python
"""Phase times (tree build / walk / integrator, CUDA events inside the library) of a Barnes-Hut step.
usage: tools/bh_phases.py [n] [dims] [theta] [ngpus]   (dims = 3: two-galaxy scene, near leaves included, accurate rsqrt;
ngpus > 1: one process driving that many GPUs, phase times of device 0, plus the wall-clock rate of 200 back-to-back steps)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
dims = int(sys.argv[2]) if len(sys.argv) > 2 else 3
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
ngpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
if dims == 3:
    b = ic.two_galaxy(n, seed=0, dims=3)
    kw = dict(dt=1e-3, eps=0.01, bh_fix_near_leaves=1)
else:
    b = ic.spinning_disc(n, seed=3, scale=100.0 * np.sqrt(n / 1024.0), spin=0.3 / np.sqrt(n / 1024.0))
    kw = dict(dt=0.01, eps=1.0, rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY)
if ngpus > 1:
    kw.update(ngpus=ngpus, device_ids=list(range(ngpus)))
with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, dims=dims, theta=theta, use_graph=0, **kw) as s:
    s.step(3); s.sync()
    if ngpus > 1:
        import time
        t0 = time.perf_counter(); s.step(200); s.sync(); wall = (time.perf_counter() - t0) / 200
        print(json.dumps({"ngpus": ngpus, "wall_ms_per_step": 1e3 * wall}))
    rows = []
    for _ in range(5):
        s.profile_next_step(True); s.step(1)
        i = s.info()
        rows.append((i["last_bh_build_ms"], i["last_force_ms"] - i["last_bh_build_ms"], i["last_integ_ms"]))
    m = np.median(np.array(rows), axis=0)
    print(json.dumps({"n": n, "dims": dims, "theta": theta, "build_ms": float(m[0]), "walk_ms": float(m[1]), "integ_ms": float(m[2]),
                      "bh_nodes": i["bh_nodes"], "visits_per_target": i["last_bh_visits"] / n}))
