#!/usr/bin/env python
"""A few Barnes-Hut steps at one size, for `ncu` launch lists / captures (no timing of its own).
usage: tools/bh_profile.py [n] [dims] [theta] [steps]      (dims = 0: the reference's own scene, uniform_disc(n), with
the collision pass -- the whole Simulation::step())"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dims = int(sys.argv[2]) if len(sys.argv) > 2 else 2
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
if dims == 0:
    b = ic.reference_disc(n)
    dims = 2
    kw = dict(dt=0.01, eps=1.0, rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY, collide=1)
elif dims == 2:
    b = ic.spinning_disc(n, seed=3, scale=100.0 * np.sqrt(n / 1024.0), spin=0.3 / np.sqrt(n / 1024.0))
    b["mass"] = np.random.default_rng(3).uniform(0.1, 3.0, n).astype(np.float32)
    kw = dict(dt=0.01, eps=1.0, rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY)
else:
    b = ic.plummer(n, seed=3, dims=3)
    kw = dict(dt=1e-3, eps=0.01, bh_fix_near_leaves=1)
with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, dims=dims, theta=theta, use_graph=0, **kw) as s:
    for _ in range(steps):
        s.step(1)
    s.sync()
    print(s.info())
