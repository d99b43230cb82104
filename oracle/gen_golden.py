#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generate the golden vectors under tests/golden/ by EXECUTING the
unmodified reference (oracle/_ref/libnbody_ref_strict.so = /root/reference headers behind
oracle/ref_harness.cpp, strict IEEE flags).  The reference ships no tests or vectors of its own
(SURVEY.md F5), so these are the pins.  Run in the build container (needs /root/reference):

    make -C oracle all && python oracle/gen_golden.py

Inputs are stored next to outputs so the tests do not depend on the IC generators staying stable.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from nbodysim_b200 import ic  # noqa: E402
from nbodysim_b200.bodies import empty_bodies  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def clean(b):
    """zero the Vec2 padding (the reference leaves it indeterminate)"""
    c = empty_bodies(b.shape[0])
    for f in ("pos", "vel", "acc", "mass", "radius"):
        c[f] = b[f]
    return c


def main():
    R = O.reference("strict")
    if R is None:
        sys.exit("oracle/_ref/libnbody_ref_strict.so missing: run `make -C oracle all` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)

    # 1. fast_inv_sqrt known answers (Quadtree.hpp:106-111)
    xs = [1.0, 2.0, 3.0, 5.0, 1.0e4, 1.0e10, 1.0e-3]
    xs += [float(np.float32(v)) for v in np.logspace(-20, 20, 193)]
    xs += [float(np.float32(v)) for v in np.linspace(1.0, 4.0, 64, endpoint=False)]
    kat = [{"x": float(np.float32(x)).hex(), "y": float(R.ref_fast_inv_sqrt(float(np.float32(x)))).hex()} for x in xs]
    json.dump({"source": "Quadtree::fast_inv_sqrt, Quadtree.hpp:106-111 (strict build)", "vectors": kat},
              open(os.path.join(OUT, "kat_fast_inv_sqrt.json"), "w"), indent=0)

    # 2. three-body direct sum
    b3 = empty_bodies(3)
    b3["pos"] = [(-1, 0), (1, 0), (0, 3)]
    b3["mass"] = [1, 1, 2]
    a3 = O.ref_acc(b3, 1.0)
    json.dump({"source": "Quadtree::acc leaf loop, Quadtree.hpp:133-144, eps=1",
               "pos": b3["pos"].tolist(), "mass": b3["mass"].tolist(),
               "acc_hex": [[float(v).hex() for v in row] for row in a3]},
              open(os.path.join(OUT, "kat_three_body.json"), "w"), indent=0)

    # 3. all-pairs accelerations + 100-step clean trajectory, N=1024 disc, reference units
    disc = ic.spinning_disc(1024, seed=12345, scale=100.0, spin=0.3, mass=1.0)
    acc = O.ref_acc(disc, 1.0)
    end = clean(O.ref_step_clean(disc, 1.0, 0.01, 100))
    np.savez_compressed(os.path.join(OUT, "disc1024.npz"), bodies=disc, eps=np.float32(1.0), dt=np.float32(0.01),
                        nsteps=100, acc=acc, end_pos=end["pos"], end_vel=end["vel"], end_acc=end["acc"])

    # 4. ragged size (not a multiple of any tile), unequal masses, planar Plummer in N-body units
    rag = ic.plummer(777, seed=7, dims=2)
    rng = np.random.default_rng(5)
    rag["mass"] = (rag["mass"] * rng.uniform(0.2, 3.0, 777)).astype(np.float32)
    acc_r = O.ref_acc(rag, 0.05)
    end_r = clean(O.ref_step_clean(rag, 0.05, 0.005, 20))
    np.savez_compressed(os.path.join(OUT, "plummer777.npz"), bodies=rag, eps=np.float32(0.05), dt=np.float32(0.005),
                        nsteps=20, acc=acc_r, end_pos=end_r["pos"], end_vel=end_r["vel"], end_acc=end_r["acc"])

    # 5. coincident bodies + a body at the origin, eps = 0 (exercises the r_sq > 0 guard)
    co = empty_bodies(6)
    co["pos"] = [(0, 0), (1, 1), (1, 1), (-2, 0.5), (0, 0), (3, -4)]
    co["mass"] = [1, 2, 3, 4, 5, 6]
    np.savez_compressed(os.path.join(OUT, "coincident6.npz"), bodies=co, acc_eps0=O.ref_acc(co, 0.0),
                        acc_eps1=O.ref_acc(co, 1.0))

    # 6. integrator extras of Simulation::iterate (clamp + soft boundary), isolated from the force:
    #    all masses zero -> attract() yields acc == 0, so iterate() applies only :129-163.
    ex = empty_bodies(256)
    r = np.random.default_rng(11)
    rad = r.uniform(1e3, 1.5e5, 256)
    ang = r.uniform(0, 2 * np.pi, 256)
    ex["pos"] = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1).astype(np.float32)
    spd = r.uniform(0, 3e3, 256)
    ang2 = r.uniform(0, 2 * np.pi, 256)
    ex["vel"] = np.stack([spd * np.cos(ang2), spd * np.sin(ang2)], 1).astype(np.float32)
    ex["mass"] = 0.0
    it = ex.copy()
    R.ref_iterate(it.ctypes.data, 256, 1.0, 1.0, 0.05, 3)
    it = clean(it)
    np.savez_compressed(os.path.join(OUT, "iterate_extras256.npz"), bodies=ex, dt=np.float32(0.05), nsteps=3,
                        end_pos=it["pos"], end_vel=it["vel"])
    # 7. Barnes-Hut: the reference's shipped algorithm (Quadtree::build + acc, theta=1, eps=1) and its
    #    real step Simulation::iterate (BH + clamp + boundary + drift; radius=0 so collide() is inert).
    bh = ic.spinning_disc(2000, seed=31, scale=140.0, spin=0.3, mass=1.0)
    bh["mass"] = np.random.default_rng(31).uniform(0.05, 3.0, 2000).astype(np.float32)
    acc_bh, nnodes = O.ref_bh_acc(bh, 1.0, 1.0)
    acc_bh05, _ = O.ref_bh_acc(bh, 0.5, 1.0)
    it = bh.copy()
    R.ref_iterate(it.ctypes.data, 2000, 1.0, 1.0, 0.01, 10)
    it = clean(it)
    np.savez_compressed(os.path.join(OUT, "bh2000.npz"), bodies=bh, theta=np.float32(1.0), eps=np.float32(1.0),
                        dt=np.float32(0.01), nsteps=10, acc=acc_bh, acc_theta05=acc_bh05, nnodes=nnodes,
                        end_pos=it["pos"], end_vel=it["vel"], end_acc=it["acc"])
    # 8. collision pass (Simulation::collide + resolve) on scenes whose colliding pairs are disjoint, so the
    #    reference's unspecified pair order cannot matter; and the reference's WHOLE step (Simulation::step =
    #    BH iterate + clamp + boundary + collide) over 8 steps on such a scene.
    def scene(n, L, rmax, seed, vsig=30.0):
        r = np.random.default_rng(seed)
        c = empty_bodies(n)
        c["pos"] = r.uniform(-L, L, (n, 2))
        c["vel"] = r.normal(0, vsig, (n, 2))
        c["mass"] = r.uniform(0.1, 5, n)
        c["radius"] = r.uniform(0.2 * rmax, rmax, n)
        return c

    col = scene(5000, 20000, 40, 2)
    after = clean(O.ref_collide(col))
    ora, npairs, nres = O.orc_collide(col)
    assert nres > 20 and all(np.array_equal(ora[f].view(np.uint32), after[f].view(np.uint32)) for f in ("pos", "vel")), \
        "collision golden scene is order-dependent; pick another seed"
    st = scene(4000, 9000, 30, 7, vsig=60.0)
    cur, ocur, tot = st.copy(), st.copy(), 0
    for _ in range(8):
        R.ref_step_full(cur.ctypes.data, 4000, 1.0, 1.0, 0.01, 1)
        ocur["acc"] = O.orc_bh_acc(ocur, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(ocur.ctypes.data, 4000, 0.01, 3, 2)
        ocur, _, k = O.orc_collide(ocur)
        tot += k
        assert all(np.array_equal(ocur[f].view(np.uint32), cur[f].view(np.uint32)) for f in ("pos", "vel")), \
            "step golden scene is order-dependent; pick another seed"
    assert tot > 10
    cur = clean(cur)
    np.savez_compressed(os.path.join(OUT, "collide.npz"), bodies=col, after_pos=after["pos"], after_vel=after["vel"],
                        resolved=nres, step_bodies=st, step_nsteps=8, step_dt=np.float32(0.01), step_resolved=tot,
                        step_end_pos=cur["pos"], step_end_vel=cur["vel"], step_end_acc=cur["acc"])
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
