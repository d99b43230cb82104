"""Collision pass (Simulation::collide + resolve, Simulation.hpp:216-346): oracle restatement vs golden /
compiled reference on CPU; GPU pass through the C ABI vs golden (bit-exact where the reference's pair
order cannot matter) and vs the oracle's canonical order on chain-heavy scenes."""
import os

import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import Simulation, capi
from nbodysim_b200.bodies import empty_bodies

G = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def scene(n, L, rmax, seed, vsig=30.0):
    r = np.random.default_rng(seed)
    c = empty_bodies(n)
    c["pos"] = r.uniform(-L, L, (n, 2))
    c["vel"] = r.normal(0, vsig, (n, 2))
    c["mass"] = r.uniform(0.1, 5, n)
    c["radius"] = r.uniform(0.2 * rmax, rmax, n)
    return c


def momentum(b):
    return (b["mass"].astype(np.float64)[:, None] * b["vel"].astype(np.float64)).sum(0)


# ------------------------------------------------------------------ CPU: oracle
def test_oracle_collide_golden():
    g = np.load(os.path.join(G, "collide.npz"))
    out, npairs, nres = O.orc_collide(g["bodies"])
    assert nres == int(g["resolved"]) and npairs >= nres
    assert np.array_equal(bits(out["pos"]), bits(g["after_pos"])) and np.array_equal(bits(out["vel"]), bits(g["after_vel"]))


def test_oracle_full_step_golden():
    """Simulation::step() = BH iterate (clamp + boundary) + collide, 8 steps"""
    g = np.load(os.path.join(G, "collide.npz"))
    b = g["step_bodies"].copy()
    tot = 0
    for _ in range(int(g["step_nsteps"])):
        b["acc"] = O.orc_bh_acc(b, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(b.ctypes.data, b.shape[0], float(g["step_dt"]), 3, 2)
        b, _, k = O.orc_collide(b)
        tot += k
    assert tot == int(g["step_resolved"])
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(b[f]), bits(g["step_end_" + f])), f


def test_oracle_resolve_conserves_momentum_and_separates():
    b = scene(3000, 1500, 25, 4)            # dense: chains of collisions
    out, npairs, nres = O.orc_collide(b)
    assert nres > 500
    assert np.abs(momentum(out) - momentum(b)).max() <= 1e-3 * np.abs(b["mass"][:, None] * b["vel"]).sum()
    assert np.isfinite(out["pos"]).all() and np.isfinite(out["vel"]).all()


@pytest.mark.skipif(O.reference("strict") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("n,L,rmax,seed", [(2000, 3000, 15, 1), (5000, 20000, 40, 2), (20000, 80000, 60, 3)])
def test_oracle_collide_vs_reference_when_order_cannot_matter(n, L, rmax, seed):
    b = scene(n, L, rmax, seed)
    out, _, nres = O.orc_collide(b)
    ref = O.ref_collide(b)
    assert nres > 10
    assert np.array_equal(bits(out["pos"]), bits(ref["pos"])) and np.array_equal(bits(out["vel"]), bits(ref["vel"]))


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_collide_bitexact_vs_golden():
    g = np.load(os.path.join(G, "collide.npz"))
    with Simulation(g["bodies"], dims=2, eps=1.0, collide=1) as s:
        s.collide()
        out = s.bodies.copy()
        cand, res = s.collide_stats()
    assert res == int(g["resolved"])
    assert np.array_equal(bits(out["pos"]), bits(g["after_pos"])) and np.array_equal(bits(out["vel"]), bits(g["after_vel"]))


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,rmax,seed", [(3000, 1500, 25, 4), (20000, 80000, 60, 3), (1000, 300, 30, 9)])
def test_gpu_collide_equals_oracle_canonical_order(n, L, rmax, seed):
    """chains included: the GPU pass and the oracle resolve in the same canonical order"""
    b = scene(n, L, rmax, seed)
    want, _, nres = O.orc_collide(b)
    with Simulation(b, dims=2, eps=1.0, collide=1) as s:
        s.collide()
        out = s.bodies.copy()
        cand, res = s.collide_stats()
    assert res == nres
    assert np.array_equal(bits(out["pos"]), bits(want["pos"])) and np.array_equal(bits(out["vel"]), bits(want["vel"]))


def chain_scene():
    """A-B-C-D: only A and B overlap at first; resolving (A,B) pushes B into C, resolving (B,C) pushes C into D.
    All x intervals overlap (sweep pairs exist), offsets in y keep the later pairs apart until their turn."""
    c = empty_bodies(6)
    c["pos"] = [[0, 0], [15, 0], [30, 14], [45, 28], [300, 300], [310, 300.5]]      # + one ordinary overlapping pair
    c["radius"] = 10.0
    c["mass"] = [1.0, 1.0, 1.0, 1.0, 2.0, 3.0]
    c["vel"] = [[0, 0], [0, 0], [0, 0], [0, 0], [1, 0], [-1, 0]]
    return c


def test_oracle_chain_propagates_through_four_bodies():
    b = chain_scene()
    out, npairs, nres = O.orc_collide(b)
    assert nres == 4                                    # (A,B), (B,C), (C,D) and the ordinary pair
    assert not np.array_equal(out["pos"][3], b["pos"][3])     # D moved although neither C nor D touched anything at first


@pytest.mark.gpu
def test_gpu_chain_of_four_equals_oracle():
    """the kept pairs must be closed over chains (connected components of the sweep-pair graph), not
    just one hop away from an overlapping pair"""
    b = chain_scene()
    want, _, nres = O.orc_collide(b)
    with Simulation(b, dims=2, eps=1.0, collide=1) as s:
        s.collide()
        out = s.bodies.copy()
        _, res = s.collide_stats()
    assert res == nres == 4
    assert np.array_equal(bits(out["pos"]), bits(want["pos"])) and np.array_equal(bits(out["vel"]), bits(want["vel"]))


@pytest.mark.gpu
def test_gpu_collision_buffer_overflow_is_reported():
    """a body spanning more than 4096 grid cells abandons the pass: sync / download must say so"""
    from nbodysim_b200.simulation import NbodyError

    b = scene(500, 1500, 25, 4)
    b["radius"][7] = 1.0e5
    with Simulation(b, dims=2, eps=1.0, collide=1) as s:
        s.collide()
        with pytest.raises(NbodyError) as ei:
            s.sync()
        assert ei.value.code == capi.ESTATE and "overflow" in str(ei.value)


@pytest.mark.gpu
@pytest.mark.parametrize("sort_impl", [0, 2])
def test_gpu_full_reference_step_bitexact_vs_golden(sort_impl):
    """nbody_gpu_step with BH + clamp + boundary + collide == Simulation::step(), 8 steps, bit for bit
    (with the library radix sort and with the hand-written one)"""
    g = np.load(os.path.join(G, "collide.npz"))
    with Simulation(g["step_bodies"], dt=float(g["step_dt"]), dims=2, theta=1.0, eps=1.0, collide=1, sort_impl=sort_impl,
                    force_algo=capi.FORCE_BARNES_HUT, rsqrt_mode=capi.RSQRT_REFCOMPAT,
                    integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
        s.step(int(g["step_nsteps"]))
        out = s.bodies
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(out[f]), bits(g["step_end_" + f])), f


@pytest.mark.gpu
def test_gpu_full_step_with_collisions_graph_replay_is_bit_identical():
    """the collision pass keeps its counters on the device, so whole steps (BH + integrate + collide) replay from a
    CUDA graph: 24 steps of a collision-dense scene, graph replay against plain launches"""
    b = scene(3000, 1500, 25, 4)
    b["vel"] = np.random.default_rng(5).normal(0.0, 30.0, (3000, 2)).astype(np.float32)
    outs = []
    for use_graph in (0, 1):
        with Simulation(b, dt=0.01, dims=2, theta=1.0, eps=1.0, collide=1, use_graph=use_graph,
                        force_algo=capi.FORCE_BARNES_HUT, rsqrt_mode=capi.RSQRT_REFCOMPAT,
                        integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
            s.step(24)
            outs.append(s.bodies.copy())
            assert s.info()["graph"] == use_graph
            assert s.collide_stats()[0] > 0
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(outs[0][f]), bits(outs[1][f])), f


@pytest.mark.gpu
def test_gpu_collide_inert_for_zero_radii_and_rejected_in_3d():
    import ctypes as C
    from nbodysim_b200 import ic
    from nbodysim_b200.simulation import default_params

    b = ic.spinning_disc(2000, seed=3)
    with Simulation(b, dims=2, eps=1.0, collide=1) as s:
        s.collide()
        out = s.bodies.copy()
        assert s.collide_stats()[1] == 0
    assert np.array_equal(bits(out["pos"]), bits(b["pos"]))
    p = default_params(dims=3, collide=1)
    ctx = C.c_void_p()
    assert capi.gpu_lib().nbody_gpu_init(C.byref(ctx), C.byref(p), b.ctypes.data, 2000) == capi.EINVAL
