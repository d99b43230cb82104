#!/usr/bin/env python
"""fp64 tolerance mode: throughput of force_f64_kernel and its agreement with the fp32 kernel."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nbodysim_b200 import Simulation, capi, ic
from nbodysim_b200.bodies import acc3

for n in [int(a) for a in sys.argv[1:]] or [65536, 262144]:
    b = ic.plummer(n, seed=1, dims=3)
    with Simulation(b, dt=1e-3, eps=0.01, dims=3, precision=capi.PRECISION_F64) as s64, Simulation(b, dt=1e-3, eps=0.01, dims=3) as s32:
        s64.attract(); s64.sync()
        t = time.perf_counter(); s64.attract(); s64.sync(); dt = time.perf_counter() - t
        _, _, a64 = s64.download_f64()
        s32.attract()
        a32 = acc3(s32.download()).astype(np.float64)
        rel = np.linalg.norm(a32 - a64, axis=1) / np.linalg.norm(a64, axis=1)
        print(json.dumps({"n": n, "f64_ms": 1e3 * dt, "f64_G_inter_per_s": n * n / dt / 1e9,
                          "fp32_vs_fp64_rel_err_median": float(np.median(rel)), "p99": float(np.percentile(rel, 99)),
                          "max": float(rel.max())}), flush=True)
