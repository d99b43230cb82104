"""The drop-in EXECUTED: the reference's own `Simulation` class with oracle/integration.patch applied
-- Simulation::step()'s iterate(dt) + collide() (Simulation.hpp:67-75) replaced by nbody_gpu_step + nbody_gpu_download --
compiled against the reference's unmodified headers and linked with libnbody_gpu.so (oracle/_ref/dropin_patched, built
where /root/reference exists, run on the GPU box), against the UNPATCHED Simulation::step() of the same headers
(oracle/_ref/libnbody_ref_strict.so) on the reference's shipped scene uniform_disc(25000)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import BODY_DTYPE
from nbodysim_b200.bodies import empty_bodies

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_patched")
REF_HDR = "/root/reference/Nbodysim/headers/Simulation.hpp"
N = 25000                      # Simulation::Simulation(), Simulation.hpp:61


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.skipif(not os.path.exists(REF_HDR) or shutil.which("patch") is None, reason="/root/reference absent (GPU box)")
def test_integration_patch_applies_to_the_reference_header(tmp_path):
    out = tmp_path / "Simulation.hpp"
    subprocess.check_call(["patch", "-s", "-o", str(out), REF_HDR, os.path.join(ROOT, "oracle", "integration.patch")])
    text = out.read_text()
    assert "nbody_gpu_step(gpu, current_dt, 1)" in text and "nbody_gpu_init(&gpu" in text
    assert "\n        iterate(current_dt);" not in text            # the CPU hot path is no longer called from step()
    assert os.path.exists(EXE), "make -C oracle ref builds oracle/_ref/dropin_patched from the patched header"


def unpatched(nsteps):
    R = O.reference("strict")
    b = empty_bodies(N)
    R.ref_uniform_disc(b.ctypes.data, N)
    start = b.copy()
    R.ref_step_full(b.ctypes.data, N, 1.0, 1.0, 0.01, nsteps)
    return start, b


def run_patched(tmp_path, nsteps, *extra):
    out = tmp_path / f"bodies_{nsteps}_{len(extra)}.bin"
    r = subprocess.run([EXE, str(nsteps), str(out), *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.fromfile(out, dtype=BODY_DTYPE), json.loads(r.stdout.strip().splitlines()[-1])


needs = pytest.mark.skipif(not os.path.exists(EXE) or O.reference("strict") is None, reason="oracle/_ref not built")


@pytest.mark.gpu
@needs
def test_patched_reference_simulation_equals_unpatched_after_5_steps(tmp_path):
    got, info = run_patched(tmp_path, 5)
    start, want = unpatched(5)
    assert info["frame"] == 5 and info["n"] == N and got.shape[0] == N
    for f in ("mass", "radius"):
        assert np.array_equal(bits(got[f]), bits(want[f])), f
    assert np.array_equal(bits(got["acc"]), bits(want["acc"]))
    # bodies beyond the soft boundary go through expf (1-2 ulp between glibc and CUDA): bit-exact inside, tolerance outside
    inside = (start["pos"].astype(np.float64) ** 2).sum(1) < (0.79e5) ** 2
    assert inside.sum() > 0.8 * N
    for f in ("pos", "vel"):
        assert np.array_equal(bits(got[f][inside]), bits(want[f][inside])), f
    np.testing.assert_allclose(got["vel"], want["vel"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(got["pos"], want["pos"], rtol=2e-6, atol=1e-3)


@pytest.mark.gpu
@needs
def test_patched_simulation_under_the_viewer_threads(tmp_path):
    """simulation thread (step + publish under UPDATE_LOCK) and a 60 Hz consumer, as main.cpp:612-635 runs them:
    same state as the plain loop, bit for bit"""
    plain, _ = run_patched(tmp_path, 30)
    thr, info = run_patched(tmp_path, 30, "threaded")
    assert info["threaded"] is True and info["frame"] == 30 and info["render_frames"] >= 1
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(plain[f]), bits(thr[f])), f
