"""GPU Barnes-Hut path (-m gpu) through the C ABI against the oracle / golden vectors: the tree as a
set of cells (centres of mass, quads, walk order) and the accelerations are BIT-EXACT with the
reference's Quadtree::build + acc, quirks included; the full reference step Simulation::iterate
(BH + clamp + soft boundary + drift) is reproduced bit for bit over 10 steps."""
import os

import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import Simulation, capi, ic

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def oracle_preorder(nodes):
    """non-empty nodes of the oracle/reference tree in depth-first quadrant order (the walk order)"""
    out, stack = [], [0]
    while stack:
        i = stack.pop()
        nd = nodes[i]
        if nd["children"] == 0 and nd["mass"] == 0.0 and i != 0:
            continue                                   # empty leaf: the GPU build does not materialise it
        out.append(i)
        if nd["children"] != 0:
            c = int(nd["children"])
            stack.extend([c + 3, c + 2, c + 1, c])
    return nodes[out]


def bh_sim(b, **kw):
    return Simulation(b, force_algo=capi.FORCE_BARNES_HUT, dims=2, rsqrt_mode=capi.RSQRT_REFCOMPAT, **kw)


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (3, 3), (100, 4), (5000, 5), (25000, 6)])
def test_tree_matches_reference_cells(n, seed):
    b = ic.spinning_disc(n, seed=seed, scale=100.0 * np.sqrt(max(n, 1024) / 1024.0))
    if n > 2:
        b["mass"] = np.random.default_rng(seed).uniform(0.1, 3.0, n).astype(np.float32)
    want = oracle_preorder(O.orc_bh_build(b))
    with bh_sim(b, theta=1.0, eps=1.0) as s:
        s.attract()
        f8, nxt, depth, leaf = s.bh_nodes()
        assert s.info()["bh_nodes"] == f8.shape[0]
    assert f8.shape[0] == want.shape[0]
    for col, k in ((0, "px"), (1, "py"), (3, "mass"), (4, "cx"), (5, "cy"), (7, "size")):
        assert np.array_equal(bits(f8[:, col]), bits(want[k])), k
    assert not f8[:, 2].any() and not f8[:, 6].any()
    assert np.array_equal(depth, want["depth"].astype(np.uint32))
    assert np.array_equal(leaf, want["children"] == 0)


@pytest.mark.parametrize("theta,key", [(1.0, "acc"), (0.5, "acc_theta05")])
def test_bh_acc_bitexact_vs_golden(theta, key):
    g = np.load(os.path.join(G, "bh2000.npz"))
    with bh_sim(g["bodies"], theta=theta, eps=1.0) as s:
        s.attract()
        out = s.download()["acc"].copy()
    assert np.array_equal(bits(out), bits(g[key]))


def test_bh_reference_step_bitexact_vs_golden():
    """Simulation::iterate on the GPU: 10 steps of BH(theta=1) + clamp + soft boundary + drift."""
    g = np.load(os.path.join(G, "bh2000.npz"))
    with bh_sim(g["bodies"], dt=float(g["dt"]), theta=1.0, eps=1.0,
                integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
        s.step(int(g["nsteps"]))
        out = s.bodies
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(out[f]), bits(g["end_" + f])), f


@pytest.mark.parametrize("n,seed,theta", [(7, 3, 1.0), (3000, 9, 0.7), (25000, 5, 1.0), (100000, 7, 1.0)])
def test_bh_acc_bitexact_vs_oracle_seeded(n, seed, theta):
    b = ic.spinning_disc(n, seed=seed, scale=100.0 * np.sqrt(max(n, 1024) / 1024.0))
    b["mass"] = np.random.default_rng(seed).uniform(0.1, 3.0, n).astype(np.float32)
    with bh_sim(b, theta=theta, eps=1.0) as s:
        s.attract()
        out = s.download()["acc"].copy()
    assert np.array_equal(bits(out), bits(O.orc_bh_acc(b, theta, 1.0)))


@pytest.mark.parametrize("sort_impl", [1, 2])
def test_bh_sort_implementations_bitexact(sort_impl):
    """the hand-written single-pass sort + scan (default) and cub's give the same tree"""
    b = ic.spinning_disc(50000, seed=23, scale=700.0)
    b["mass"] = np.random.default_rng(23).uniform(0.1, 3.0, 50000).astype(np.float32)
    b[100]["pos"] = b[7]["pos"]            # coincident pair: merged in index order -> needs a STABLE sort
    with bh_sim(b, theta=1.0, eps=1.0, sort_impl=sort_impl) as s:
        s.attract()
        out = s.download()["acc"].copy()
    assert np.array_equal(bits(out), bits(O.orc_bh_acc(b, 1.0, 1.0)))


def test_radix_sort_checker_tool():
    """tools/sort_check.cu: the hand-written sort against std::stable_sort on 46 cases"""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "sort_check")
    if not os.path.exists(exe):
        pytest.skip("build/sort_check not built (make tools)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("walk", [1, 2])
def test_bh_both_walks_bitexact(walk):
    """per-thread (1) and warp-cooperative (2) walks visit each target's nodes in the same order"""
    b = ic.spinning_disc(30000, seed=21, scale=600.0)
    b["mass"] = np.random.default_rng(21).uniform(0.1, 3.0, 30000).astype(np.float32)
    with bh_sim(b, theta=0.8, eps=1.0, bh_walk=walk) as s:
        s.attract()
        out = s.download()["acc"].copy()
    assert np.array_equal(bits(out), bits(O.orc_bh_acc(b, 0.8, 1.0)))


def test_bh_graph_replay_bitexact_vs_golden():
    """the whole Barnes-Hut step (build incl. sort, scan, COM pass; walk; integrator) replayed from
    a CUDA graph: 10 reference steps, still bit-exact"""
    g = np.load(os.path.join(G, "bh2000.npz"))
    with bh_sim(g["bodies"], dt=float(g["dt"]), theta=1.0, eps=1.0, use_graph=1,
                integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
        s.step(int(g["nsteps"]))
        out = s.bodies
        assert s.info()["graph"] == 1
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(out[f]), bits(g["end_" + f])), f


def test_bh_coincident_bodies_merge():
    b = ic.spinning_disc(64, seed=8)
    b[10]["pos"] = b[3]["pos"]
    b[40]["pos"] = b[3]["pos"]
    with bh_sim(b, theta=1.0, eps=1.0) as s:
        s.attract()
        out = s.download()["acc"].copy()
    assert np.array_equal(bits(out), bits(O.orc_bh_acc(b, 1.0, 1.0)))


def test_bh_fixed_near_leaves_and_fast_rsqrt():
    b = ic.spinning_disc(4000, seed=12, scale=200.0)
    want_fix = O.orc_bh_acc(b, 1.0, 1.0, fix_near_leaves=True)
    with bh_sim(b, theta=1.0, eps=1.0, bh_fix_near_leaves=1) as s:
        s.attract()
        got_fix = s.download()["acc"].copy()
    assert np.array_equal(bits(got_fix), bits(want_fix))
    # accurate rsqrt in the walk: same tree, forces higher by the Quake deficit (0 .. 0.53 %)
    with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=1.0, eps=1.0) as s:
        s.attract()
        fast = s.download()["acc"].astype(np.float64)
    ref = O.orc_bh_acc(b, 1.0, 1.0).astype(np.float64)
    rel = np.linalg.norm(fast - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert 5e-4 < np.median(rel) < 5e-3


def test_bh_rejects_f64():
    b = ic.plummer(128, dims=3)
    lib = capi.gpu_lib()
    import ctypes as C
    from nbodysim_b200.simulation import default_params

    for kw in ({"dims": 3, "precision": capi.PRECISION_F64}, {"dims": 2, "precision": capi.PRECISION_F64}):
        p = default_params(force_algo=capi.FORCE_BARNES_HUT, **kw)
        ctx = C.c_void_p()
        assert lib.nbody_gpu_init(C.byref(ctx), C.byref(p), b.ctypes.data, 128) == capi.EINVAL


# ------------------------------------------------------------------ octree (dims = 3)
def oracle_preorder3(nodes):
    out, stack = [], [0]
    while stack:
        i = stack.pop()
        nd = nodes[i]
        if nd["children"] == 0 and nd["mass"] == 0.0 and i != 0:
            continue
        out.append(i)
        if nd["children"] != 0:
            c = int(nd["children"])
            stack.extend(range(c + 7, c - 1, -1))
    return nodes[out]


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (9, 3), (500, 4), (20000, 5)])
def test_octree_matches_oracle_cells(n, seed):
    """dims=3: the same construction one dimension up (8 children, 3 bits x 21 levels); cells, centres of mass
    and walk order equal the oracle's octree bit for bit"""
    b = ic.plummer(n, seed=seed, dims=3)
    if n > 2:
        b["mass"] = (b["mass"] * np.random.default_rng(seed).uniform(0.2, 3.0, n)).astype(np.float32)
    want = oracle_preorder3(O.orc_bh3_build(b))
    with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, dims=3, theta=0.7, eps=0.01, rsqrt_mode=capi.RSQRT_REFCOMPAT) as s:
        s.attract()
        f8, nxt, depth, leaf = s.bh_nodes()
    assert f8.shape[0] == want.shape[0]
    for col, k in enumerate(("px", "py", "pz", "mass", "cx", "cy", "cz", "size")):
        assert np.array_equal(bits(f8[:, col]), bits(want[k])), k
    assert np.array_equal(depth, want["depth"].astype(np.uint32)) and np.array_equal(leaf, want["children"] == 0)


@pytest.mark.parametrize("theta,fix,walk", [(1.0, 0, 0), (0.5, 1, 1), (0.5, 1, 2), (0.3, 0, 2)])
def test_octree_acc_bitexact_vs_oracle(theta, fix, walk):
    b = ic.plummer(30000, seed=11, dims=3)
    with Simulation(b, force_algo=capi.FORCE_BARNES_HUT, dims=3, theta=theta, eps=0.01, rsqrt_mode=capi.RSQRT_REFCOMPAT,
                    bh_fix_near_leaves=fix, bh_walk=walk) as s:
        s.attract()
        out = s.download().copy()
    want = O.orc_bh3_acc(b, theta, 0.01, fix_near_leaves=bool(fix))
    got = np.concatenate([out["acc"], out["acc_z"][:, None]], axis=1)
    assert np.array_equal(bits(got), bits(want))


def test_octree_fast_accuracy_and_energy():
    """accurate rsqrt + near leaves included + theta = 0.4: a usable O(N log N) 3-D force (median error ~1e-3),
    and a 200-step run conserves energy to 1e-3"""
    b = ic.plummer(50000, seed=8, dims=3)
    with Simulation(b, dt=1e-3, force_algo=capi.FORCE_BARNES_HUT, dims=3, theta=0.4, eps=0.01, bh_fix_near_leaves=1) as s:
        s.attract()
        out = s.download().copy()
        a = np.concatenate([out["acc"], out["acc_z"][:, None]], axis=1).astype(np.float64)
        ex = O.orc_exact_acc(b, float(np.float32(0.01)), dims=3, i0=0, i1=2000)
        rel = np.linalg.norm(a[:2000] - ex, axis=1) / np.linalg.norm(ex, axis=1)
        assert np.median(rel) < 3e-3 and np.percentile(rel, 99) < 3e-2, (np.median(rel), np.percentile(rel, 99))
        k0, w0, _ = s.energy()
        s.step(200)
        k1, w1, _ = s.energy()
    assert abs((k1 + w1 - k0 - w0) / (k0 + w0)) < 1e-3


def close_pairs_scene(n_pairs=600, seed=31):
    """every body has a twin two ulps away: each pair hangs from a chain of ~20 single-child cells, so the tree needs
    far more than the default 4 cells per body"""
    r = np.random.default_rng(seed)
    from nbodysim_b200.bodies import empty_bodies

    b = empty_bodies(2 * n_pairs)
    p = r.uniform(500.0, 1000.0, (n_pairs, 2)).astype(np.float32)
    b["pos"][0::2] = p
    b["pos"][1::2] = np.nextafter(np.nextafter(p, np.float32(2000)), np.float32(2000))
    b["mass"] = r.uniform(0.5, 2.0, 2 * n_pairs).astype(np.float32)
    return b


def test_bh_cell_overflow_is_reported_and_capacity_can_be_raised(monkeypatch):
    """a tree that needs more cells than reserved must not truncate forces silently"""
    from nbodysim_b200.simulation import NbodyError

    b = close_pairs_scene()
    ncells = oracle_preorder(O.orc_bh_build(b)).shape[0]
    assert ncells > 4 * b.shape[0] + 1024
    monkeypatch.delenv("NBODY_BH_NODE_FACTOR", raising=False)
    with bh_sim(b, theta=1.0, eps=1.0) as s:
        s.attract()
        with pytest.raises(NbodyError) as ei:
            s.sync()
        assert ei.value.code == capi.ENOMEM and "NBODY_BH_NODE_FACTOR" in str(ei.value)
    monkeypatch.setenv("NBODY_BH_NODE_FACTOR", "33")
    with bh_sim(b, theta=1.0, eps=1.0) as s:
        s.attract()
        out = s.download()["acc"].copy()
        assert s.info()["bh_nodes"] == ncells
    assert np.array_equal(bits(out), bits(O.orc_bh_acc(b, 1.0, 1.0)))


@pytest.mark.parametrize("sort_impl", [1, 2])
def test_build_paths_and_fused_walk_bitexact(sort_impl):
    """the two build paths (one kernel per phase, single cluster kernel) and the two integrators (walk threads integrate
    their own targets -- small scenes -- / stand-alone kernel) change the execution, never a bit of the result"""
    b = ic.spinning_disc(25000, seed=5, scale=100.0 * np.sqrt(25000 / 1024.0))
    b["mass"] = np.random.default_rng(5).uniform(0.1, 3.0, 25000).astype(np.float32)
    with bh_sim(b, theta=1.0, eps=1.0, sort_impl=sort_impl) as s:
        s.attract()
        out = s.download()["acc"].copy()
    assert np.array_equal(bits(out), bits(O.orc_bh_acc(b, 1.0, 1.0)))
    want = b.copy()
    for _ in range(3):
        want["acc"] = O.orc_bh_acc(want, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(want.ctypes.data, want.shape[0], 0.01, 3, 2)
    outs = []
    for fuse in (-1, 0):
        with bh_sim(b, dt=0.01, theta=1.0, eps=1.0, sort_impl=sort_impl, fuse_integrator=fuse,
                    integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
            s.step(3)
            outs.append(s.bodies.copy())
    for o in outs:
        for f in ("pos", "vel", "acc"):
            assert np.array_equal(bits(o[f]), bits(want[f])), f


def test_per_frame_calls_replay_one_step_graphs_bit_identically():
    """a caller that steps once per call (viewer loop, upload / step / download cycle) is switched to one-step CUDA graphs
    (one per buffer parity) after a few calls: same state as one long call, bit for bit; a changed dt re-captures"""
    b = ic.reference_disc(6000)
    kw = dict(theta=1.0, eps=1.0, collide=1, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY)
    with bh_sim(b, dt=0.01, **kw) as s:
        s.step(14)
        s.dt = 0.005
        s.step(9)
        want = s.bodies.copy()
    with bh_sim(b, dt=0.01, **kw) as s:
        for _ in range(14):
            s.step(1)
        assert s.info()["graph"] == 1
        s.dt = 0.005
        for _ in range(3):
            s.step(3)
        got = s.bodies.copy()
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(got[f]), bits(want[f])), f


def test_bh_reservation_grows_before_it_overflows():
    """a tree that fills more than 3/4 of the cell reservation makes the next nbody_gpu_step call double it (buffers
    reallocated, graphs dropped): the steps before and after stay bit-exact and nothing overflows"""
    from nbodysim_b200.bodies import empty_bodies

    r = np.random.default_rng(31)
    n_pairs, n_single = 350, 2000
    n = 2 * n_pairs + n_single
    b = empty_bodies(n)
    p = r.uniform(500.0, 1000.0, (n_pairs, 2)).astype(np.float32)
    b["pos"][0:2 * n_pairs:2] = p
    b["pos"][1:2 * n_pairs:2] = np.nextafter(np.nextafter(p, np.float32(2000)), np.float32(2000))
    b["pos"][2 * n_pairs:] = r.uniform(500.0, 1000.0, (n_single, 2)).astype(np.float32)
    b["mass"] = r.uniform(0.5, 2.0, n).astype(np.float32)
    ncells = oracle_preorder(O.orc_bh_build(b)).shape[0]
    assert 0.75 * (4 * n + 1024) < ncells < 4 * n + 1024
    want = b.copy()
    for _ in range(4):
        want["acc"] = O.orc_bh_acc(want, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(want.ctypes.data, n, 0.01, 0, 2)
    with bh_sim(b, dt=0.01, theta=1.0, eps=1.0) as s:
        s.step(1)
        s.sync()
        s.step(3)                      # the reservation is doubled here, before the steps are enqueued
        got = s.bodies.copy()
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(got[f]), bits(want[f])), f


@pytest.mark.parametrize("n", [6000, 30000])
def test_execution_variants_bit_identical(n):
    """the A/B switches of the build, the sort and the collision pass change how a step runs, never a bit of its result:
    tools/step_checksum.py (state after 6 reference steps with collisions + every cell of the next tree) under each switch.
    NBODY_BH_TOP_MAX_EDGES=100 forces the one-CTA top-of-tree kernel into its fall-back (the atomic climb)."""
    import subprocess
    import sys

    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "step_checksum.py")
    variants = [{}, {"NBODY_BH_LOCAL": "0"}, {"NBODY_SORT_LAZY": "0"}, {"NBODY_BH_CTA_CLIMB": "0"}, {"NBODY_COL_STRIP": "0"},
                {"NBODY_BH_FUSE_INSERT": "0"}, {"NBODY_SORT_COOP": "0"}, {"NBODY_BHL_GEOM": "2"}, {"NBODY_BHL_GEOM": "6"}, {"NBODY_BH_TOP_MAX_EDGES": "100"}, {"NBODY_PDL": "0"}]
    lines = []
    for v in variants:
        env = dict(os.environ, **v)
        r = subprocess.run([sys.executable, tool, str(n), "6"], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (v, r.stderr[-2000:])
        lines.append(r.stdout.strip().splitlines()[-1])
    if n == 6000:     # sort_impl = 1: one kernel per phase, in the build and in the collision pass (its large-scene pipeline)
        r = subprocess.run([sys.executable, tool, str(n), "6", "1"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        variants.append({"sort_impl": "1"})
        lines.append(r.stdout.strip().splitlines()[-1])
    assert all(l == lines[0] for l in lines), list(zip(variants, lines))
