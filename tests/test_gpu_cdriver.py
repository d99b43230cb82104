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


def test_c_driver_reference_scene_runs():
    out = run(["--ic", "reference", "--n", "25000", "--steps", "20", "--algo", "bh", "--rsqrt", "refcompat", "--clamp", "on",
               "--boundary", "on", "--collide", "on"])
    assert "20 steps in" in out
