#!/usr/bin/env python
"""Barnes-Hut path: GPU (tree build + walk + integrator, through the C ABI) next to the reference's
own Simulation::iterate on the host cores (oracle/_ref fast build = the reference's flags), at the
reference's shipped size N=25,000 and larger.  Not the headline metric (that is all-pairs, bench.py);
this measures the reference's REAL algorithm.  Prints one JSON line per N."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [25000, 100000, 1000000]
    R = O.reference("fast")
    for n in sizes:
        b = ic.spinning_disc(n, seed=3, scale=100.0 * np.sqrt(n / 1024.0), spin=0.3 / np.sqrt(n / 1024.0))
        b["mass"] = np.random.default_rng(3).uniform(0.1, 3.0, n).astype(np.float32)
        row = {"n": n, "theta": 1.0, "eps": 1.0}
        for mode, walk, name in ((capi.RSQRT_REFCOMPAT, 0, "refcompat"), (capi.RSQRT_REFCOMPAT, 2, "refcompat_warpwalk"),
                                 (capi.RSQRT_FAST, 0, "fast")):
            with Simulation(b, dt=0.01, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=1.0, eps=1.0, rsqrt_mode=mode,
                            bh_walk=walk, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
                s.step(3); s.sync()
                f, i = [], []
                for _ in range(10):
                    s.profile_next_step(True)
                    s.step(1)
                    inf = s.info()
                    f.append(inf["last_force_ms"]); i.append(inf["last_integ_ms"])
                s.step(20); s.sync()                      # warm the CUDA-graph path (capture + instantiate)
                t0 = time.perf_counter(); s.step(100); s.sync(); wall = (time.perf_counter() - t0) / 100
                row[f"gpu_{name}_force_ms"] = float(np.median(f))
                row[f"gpu_{name}_integ_ms"] = float(np.median(i))
                row[f"gpu_{name}_step_wall_ms"] = 1e3 * wall
                row["bh_nodes"] = inf["bh_nodes"]
        if R is not None and n <= 200000:
            c = b.copy()
            R.ref_iterate(c.ctypes.data, n, 1.0, 1.0, 0.01, 1)      # constructs the Simulation singleton + warm-up
            k = 5 if n <= 50000 else 2
            t0 = time.perf_counter()
            R.ref_iterate(c.ctypes.data, n, 1.0, 1.0, 0.01, k)
            row["cpu_reference_iterate_ms"] = 1e3 * (time.perf_counter() - t0) / k
            row["cpu_threads"] = os.cpu_count()
            row["speedup_vs_cpu_reference"] = row["cpu_reference_iterate_ms"] / row["gpu_refcompat_step_wall_ms"]
        print(json.dumps(row), flush=True)


def reference_scene():
    """The reference's own shipped configuration end to end: uniform_disc(25000), theta=1, eps=1, dt=0.01,
    Simulation::step() = BH iterate (clamp + boundary) + collide -- GPU vs the reference on the host cores."""
    n = 25000
    b = ic.reference_disc(n)
    row = {"scene": "reference uniform_disc", "n": n}
    with Simulation(b, dt=0.01, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=1.0, eps=1.0, collide=1,
                    rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
        s.step(5); s.sync()
        t0 = time.perf_counter(); s.step(200); s.sync()
        row["gpu_full_step_ms"] = 1e3 * (time.perf_counter() - t0) / 200
    R = O.reference("fast")
    if R is not None:
        c = b.copy()
        R.ref_step_full(c.ctypes.data, n, 1.0, 1.0, 0.01, 2)
        t0 = time.perf_counter()
        R.ref_step_full(c.ctypes.data, n, 1.0, 1.0, 0.01, 20)
        row["cpu_reference_step_ms"] = 1e3 * (time.perf_counter() - t0) / 20
        row["cpu_threads"] = os.cpu_count()
        row["speedup"] = row["cpu_reference_step_ms"] / row["gpu_full_step_ms"]
    print(json.dumps(row), flush=True)


def octree():
    """dims=3 generalisation: Plummer spheres, theta = 0.5, near leaves included, accurate rsqrt"""
    for n in (1048576, 4194304):
        b = ic.plummer(n, seed=3, dims=3)
        with Simulation(b, dt=1e-3, force_algo=capi.FORCE_BARNES_HUT, dims=3, theta=0.5, eps=0.01, bh_fix_near_leaves=1) as s:
            s.step(20); s.sync()
            t0 = time.perf_counter(); s.step(20); s.sync()
            ms = 1e3 * (time.perf_counter() - t0) / 20
            s.profile_next_step(True); s.step(1)
            inf = s.info()
            print(json.dumps({"octree_n": n, "theta": 0.5, "step_ms": ms, "force_ms": inf["last_force_ms"], "nodes": inf["bh_nodes"],
                              "allpairs_step_ms_same_gpu": n * n / 2.93e12 * 1e3}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "octree":
        octree()
        sys.exit(0)
    reference_scene()
    main()
