"""Barnes-Hut restatement in the oracle (Quadtree build / propagate / acc) against the golden vectors
and, where oracle/_ref is present, against the compiled reference on fresh inputs -- bit-exact."""
import os

import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import ic

G = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_bh_acc_golden():
    g = np.load(os.path.join(G, "bh2000.npz"))
    b = g["bodies"]
    nodes = O.orc_bh_build(b)
    assert nodes.shape[0] == int(g["nnodes"])
    assert np.array_equal(bits(O.orc_bh_acc(b, 1.0, 1.0, nodes)), bits(g["acc"]))
    assert np.array_equal(bits(O.orc_bh_acc(b, 0.5, 1.0, nodes)), bits(g["acc_theta05"]))


def test_bh_iterate_golden():
    """Simulation::iterate = BH attract + clamp + boundary + drift, 10 steps (Simulation.hpp:116-164)."""
    g = np.load(os.path.join(G, "bh2000.npz"))
    b = g["bodies"].copy()
    for _ in range(int(g["nsteps"])):
        b["acc"] = O.orc_bh_acc(b, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(b.ctypes.data, b.shape[0], float(g["dt"]), 3, 2)
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(b[f]), bits(g["end_" + f])), f


def test_bh_theta0_equals_nothing_dropped_property():
    """theta -> 0 opens every branch: with the reference's quirk (near leaves dropped) every
    contribution vanishes; with fix_near_leaves the walk becomes the direct sum over leaves."""
    b = ic.spinning_disc(300, seed=2)
    nodes = O.orc_bh_build(b)
    assert not O.orc_bh_acc(b, 0.0, 1.0, nodes).any()
    fixed = O.orc_bh_acc(b, 0.0, 1.0, nodes, fix_near_leaves=True).astype(np.float64)
    direct = O.orc_acc_f64sum(b, 1.0, dims=2)
    assert np.abs(fixed - direct).max() <= 2e-5 * np.abs(direct).max()


@pytest.mark.skipif(O.reference("strict") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("n,seed,theta", [(1, 1, 1.0), (2, 2, 1.0), (7, 3, 1.0), (500, 4, 0.7), (25000, 5, 1.0)])
def test_bh_vs_reference_bitexact(n, seed, theta):
    b = ic.spinning_disc(n, seed=seed, scale=100.0 * np.sqrt(max(n, 1024) / 1024.0))
    if n > 2:
        b["mass"] = np.random.default_rng(seed).uniform(0.1, 3.0, n).astype(np.float32)
    nodes = O.orc_bh_build(b)
    a_ref, m = O.ref_bh_acc(b, theta, 1.0)
    assert m == nodes.shape[0]
    f, u = O.ref_bh_nodes(b)
    for k, col in (("px", 0), ("py", 1), ("mass", 2), ("cx", 3), ("cy", 4), ("size", 5)):
        assert np.array_equal(bits(f[:, col]), bits(nodes[k])), k
    assert np.array_equal(u[:, 0], nodes["children"]) and np.array_equal(u[:, 1], nodes["next"])
    assert np.array_equal(bits(O.orc_bh_acc(b, theta, 1.0, nodes)), bits(a_ref))


@pytest.mark.skipif(O.reference("strict") is None, reason="oracle/_ref not built")
def test_bh_coincident_bodies_merge_like_reference():
    b = ic.spinning_disc(64, seed=8)
    b[10]["pos"] = b[3]["pos"]
    b[40]["pos"] = b[3]["pos"]
    a_ref, m = O.ref_bh_acc(b, 1.0, 1.0)
    nodes = O.orc_bh_build(b)
    assert m == nodes.shape[0]
    assert np.array_equal(bits(O.orc_bh_acc(b, 1.0, 1.0, nodes)), bits(a_ref))


def test_octree_oracle_properties():
    """the 3-D generalisation in the oracle (checker of the library's dims=3 path): with theta -> 0 and near
    leaves included the walk IS the direct sum over bodies; with theta = 0.5 it approximates exact math"""
    b = ic.plummer(1500, seed=3, dims=3)
    nodes = O.orc_bh3_build(b)
    leaves = nodes[(nodes["children"] == 0) & (nodes["mass"] > 0)]
    assert leaves.shape[0] == 1500 and abs(leaves["mass"].astype(np.float64).sum() - 1.0) < 1e-5
    assert abs(float(nodes[0]["mass"]) - 1.0) < 1e-5                      # root holds the total mass
    direct = O.orc_acc_f64sum(b, 0.01, dims=3)
    fixed = O.orc_bh3_acc(b, 0.0, 0.01, nodes, fix_near_leaves=True).astype(np.float64)
    assert np.abs(fixed - direct).max() <= 3e-5 * np.abs(direct).max()
    assert not O.orc_bh3_acc(b, 0.0, 0.01, nodes).any()                   # the reference's quirk: near leaves dropped
    ex = O.orc_exact_acc(b, 0.01, dims=3)
    a = O.orc_bh3_acc(b, 0.5, 0.01, nodes, fix_near_leaves=True).astype(np.float64)
    rel = np.linalg.norm(a - ex, axis=1) / np.linalg.norm(ex, axis=1)
    assert np.median(rel) < 8e-3
