#!/usr/bin/env python
"""BASELINE configs[4] (two-galaxy collision, long-run energy drift and trajectory agreement), scaled to
what one GPU and the CPU oracle can check: (a) 2 x 32,768 bodies, 2,000 steps: energy drift of the fp32
path, and position divergence fp32 vs fp64 mode; (b) 2 x 2,048 bodies, planar, 20 steps: refcompat GPU
trajectory vs the all-reference CPU oracle pipeline (bit-exact).  One JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from nbodysim_b200 import Simulation, capi, ic
from nbodysim_b200.bodies import pos3

out = {}
n, steps, dt, eps = 65536, 2000, 1e-3, 0.01
b = ic.two_galaxy(n, seed=5, dims=3)
with Simulation(b, dt=dt, eps=eps, dims=3) as s32, Simulation(b, dt=dt, eps=eps, dims=3, precision=capi.PRECISION_F64) as s64:
    k0, w0, _ = s32.energy()
    drift, div = [], []
    t = time.perf_counter()
    for chunk in range(4):
        s32.step(steps // 4); s64.step(steps // 4)
        k, w, p = s32.energy()
        drift.append((k + w - k0 - w0) / abs(k0 + w0))
        p64, _, _ = s64.download_f64()
        div.append(float(np.abs(pos3(s32.bodies) - p64).max()))
    out.update({"a_n": n, "a_steps": steps, "a_energy_rel_drift_per_500_steps": drift, "a_fp32_vs_fp64_max_pos_diff": div,
                "a_momentum_abs": float(np.abs(p).max()), "a_wall_s": time.perf_counter() - t})

n2, steps2 = 4096, 20
c = ic.two_galaxy(n2, seed=6, dims=2)
want = O.orc_step_clean(c, eps, dt, steps2, dims=2)
with Simulation(c, dt=dt, eps=eps, dims=2, rsqrt_mode=capi.RSQRT_REFCOMPAT) as s:
    s.step(steps2)
    got = s.bodies.copy()
out.update({"b_n": n2, "b_steps": steps2,
            "b_refcompat_vs_oracle_bitexact": bool(all(np.array_equal(got[f].view(np.uint32), want[f].view(np.uint32)) for f in ("pos", "vel", "acc")))})
print(json.dumps(out))
