#!/usr/bin/env python
"""The reference's shipped scene (uniform_disc(25000), Simulation::step = BH theta=1 + clamp + boundary + collide,
refcompat) under the execution variants of the library (sort_impl 1 = one kernel per phase, 0 = default: collision pass
as one cluster kernel, 2 = build as one cluster kernel too), with / without collide and fused integrator.  One JSON line each:
CUDA-graph replay rate, cold single-step phase times.  Sizes other than 25,000 via argv."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbodysim_b200 import Simulation, capi, ic  # noqa: E402


def run(n, sort_impl, collide=1, fuse=-1):
    b = ic.reference_disc(n)
    st = torch.cuda.Stream()
    row = {"n": n, "sort_impl": sort_impl, "collide": collide, "fuse": fuse}
    with Simulation(b, dt=0.01, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=1.0, eps=1.0, collide=collide, sort_impl=sort_impl,
                    fuse_integrator=fuse, rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY,
                    stream=st.cuda_stream) as s:
        s.step(16); s.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0 = s.info()
        e0.record(st); s.step(400); e1.record(st); s.sync(); torch.cuda.synchronize()
        i1 = s.info()
        row["graph_ms_per_step"] = e0.elapsed_time(e1) / 400
        row["launches_per_step"] = (i1["kernel_launches"] - i0["kernel_launches"]) / 400
        ph = []
        for _ in range(7):
            s.profile_next_step(True); s.step(1)
            inf = s.info()
            ph.append((inf["last_bh_build_ms"], inf["last_force_ms"] - inf["last_bh_build_ms"], inf["last_integ_ms"], inf["last_collide_ms"]))
        med = np.median(np.array(ph), axis=0)
        row.update(build_ms=float(med[0]), walk_ms=float(med[1]), integ_ms=float(med[2]), collide_ms=float(med[3]), bh_nodes=inf["bh_nodes"])
        row["checksum"] = int(np.bitwise_xor.reduce(s.bodies["pos"].view(np.uint32).ravel()))
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [25000]
    for n in sizes:
        for sort_impl in (1, 0, 2):
            run(n, sort_impl)
        run(n, 0, collide=0)
        run(n, 0, collide=0, fuse=0)
