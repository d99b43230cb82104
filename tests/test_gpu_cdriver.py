"""The plain-C host path end to end (-m gpu): host/_build/nbody_run (C driver over the C ABI) reads a
snapshot, steps on the GPU, writes a snapshot; the result must equal what the Python binding of the same
ABI produces, bit for bit -- and, in the reference's full configuration, what the oracle produces."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import Simulation, capi, ic
from nbodysim_b200.bodies import BODY_DTYPE, empty_bodies

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "_build", "nbody_run")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def write_snapshot(path, b, dims):
    hdr = np.zeros(64, dtype=np.uint8)
    hdr[12:16] = np.frombuffer(np.uint32(dims).tobytes(), dtype=np.uint8)
    hdr[16:24] = np.frombuffer(np.uint64(b.shape[0]).tobytes(), dtype=np.uint8)
    assert capi.host_lib().nbody_snapshot_write(path.encode(), hdr.ctypes.data, b.ctypes.data) == 0


def read_snapshot(path, n):
    hdr = np.zeros(64, dtype=np.uint8)
    b = empty_bodies(n)
    assert capi.host_lib().nbody_snapshot_read(path.encode(), hdr.ctypes.data, b.ctypes.data, n) == 0
    return b


def run(args):
    if not os.path.exists(EXE):
        pytest.skip("host/_build/nbody_run not built (make host)")
    r = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_c_driver_allpairs_equals_python_binding(tmp_path):
    b = ic.plummer(5000, seed=4, dims=3)
    src, dst = str(tmp_path / "in.nbody"), str(tmp_path / "out.nbody")
    write_snapshot(src, b, 3)
    run(["--in", src, "--out", dst, "--steps", "25", "--dt", "0.001", "--eps", "0.01"])
    got = read_snapshot(dst, 5000)
    with Simulation(b, dt=1e-3, eps=0.01, dims=3) as s:
        s.step(25)
        want = s.bodies
    for f in ("pos", "pos_z", "vel", "vel_z", "acc", "acc_z"):
        assert np.array_equal(bits(got[f]), bits(want[f])), f


def test_c_driver_reference_configuration_equals_oracle(tmp_path):
    """--algo bh --rsqrt refcompat --clamp on --boundary on --collide on == Simulation::step()"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "collide.npz"))
    src, dst = str(tmp_path / "in.nbody"), str(tmp_path / "out.nbody")
    write_snapshot(src, g["step_bodies"], 2)
    out = run(["--in", src, "--out", dst, "--steps", str(int(g["step_nsteps"])), "--dt", "0.01", "--eps", "1", "--theta", "1",
               "--algo", "bh", "--rsqrt", "refcompat", "--clamp", "on", "--boundary", "on", "--collide", "on"])
    assert "Barnes-Hut" in out
    got = read_snapshot(dst, g["step_bodies"].shape[0])
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(got[f]), bits(g["step_end_" + f])), f


def test_c_driver_no_argument_run_is_the_reference_simulation(tmp_path):
    """no flags = Simulation(): uniform_disc(25000), BH theta=1, eps=1, dt=0.01, clamp + boundary + collide,
    the reference's rsqrt; checked against the oracle pipeline on the same scene (first 3 steps)"""
    dst = str(tmp_path / "out.nbody")
    out = run(["--steps", "3", "--out", dst])
    assert "n=25000" in out and "3 steps in" in out
    got = read_snapshot(dst, 25000)
    want = ic.reference_disc(25000)
    for _ in range(3):
        want["acc"] = O.orc_bh_acc(want, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(want.ctypes.data, 25000, 0.01, 3, 2)
        want, _, _ = O.orc_collide(want)
    assert np.array_equal(bits(got["acc"]), bits(want["acc"]))
    inside = (want["pos"].astype(np.float64) ** 2).sum(1) < (0.79e5) ** 2     # beyond: expf differs by an ulp
    assert np.array_equal(bits(got["pos"][inside]), bits(want["pos"][inside]))
    np.testing.assert_allclose(got["pos"], want["pos"], rtol=2e-6, atol=1e-3)


def test_viewer_hand_off_threads():
    """host/nbody_viewer_feed.c: simulation thread (step, then download under the lock = `SHARED_BODIES =
    simulation->bodies`, main.cpp:612-635) next to a 60 Hz consumer thread; no torn or failed frames"""
    import json

    exe = os.path.join(ROOT, "host", "_build", "nbody_viewer_feed")
    if not os.path.exists(exe):
        pytest.skip("host/_build/nbody_viewer_feed not built (make host)")
    r = subprocess.run([exe, "--seconds", "1.5"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["failed"] == 0 and d["bad_frames"] == 0 and d["n"] == 25000
    assert d["steps"] >= 100 and d["render_frames"] >= 30          # >= ~70 steps/s and the 60 Hz consumer kept running
    assert d["publish_ms"] < 5.0
