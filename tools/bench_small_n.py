#!/usr/bin/env python
"""All-pairs step time across small and medium N (one GPU): us/step, G pair-interactions/s, chosen geometry."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodysim_b200 import Simulation, ic

for n in [int(a) for a in sys.argv[1:]] or [1024, 4096, 16384, 25000, 32768, 65536, 131072, 262144]:
    b = ic.plummer(n, seed=1, dims=3)
    with Simulation(b, dt=1e-3, eps=0.01, dims=3) as s:
        s.step(20); s.sync()
        k = max(10, min(2000, int(2e11 / (n * n))))
        t = time.perf_counter(); s.step(k); s.sync(); dt = (time.perf_counter() - t) / k
        inf = s.info()
        print(json.dumps({"n": n, "us_per_step": round(1e6 * dt, 2), "G_inter_per_s": round(n * n / dt / 1e9, 1),
                          "pct_fp32_peak_20flop": round(100 * n * n / dt * 20 / 74.45e12, 1), "splits": inf["j_splits"],
                          "ctas": inf["force_ctas"], "streamk": inf["streamk_ctas"], "graph": inf["graph"], "fused": inf["fused"]}), flush=True)
