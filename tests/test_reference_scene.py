"""The reference's own scene: Simulation::uniform_disc (Simulation.hpp:347-603) restated in C
(host/nbody_ic.c: nbody_ic_reference_disc) -- bit-identical to the compiled reference, ties of the
unstable std::sort included -- and, on the GPU, the reference's real step on that scene."""
import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import Simulation, capi, ic
from nbodysim_b200.bodies import empty_bodies


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_scene_shape_and_invariants():
    b = ic.reference_disc(25000)          # the shipped size, Simulation.hpp:61
    assert b[0]["mass"] == np.float32(1e9) and b[0]["radius"] == 200.0 and not b[0]["pos"].any()
    x, y = b["pos"][:, 0], b["pos"][:, 1]
    r2 = (x * x + y * y).astype(np.float64)                 # Vec2::mag_sq in fp32, the sort key
    assert (np.diff(r2) >= 0).all()                         # sorted by |pos|
    assert np.allclose(b["radius"][1:], np.cbrt(b["mass"][1:]), rtol=1e-6)
    m = b["mass"][1:]
    assert m.min() >= 0.00005 and m.max() <= 50.0
    assert 0.75 < (m <= 0.8).mean() < 0.92                  # ~ 84.6 % light bucket
    assert np.sqrt(r2.max()) > 0.8e5                        # the soft boundary is active from step 0


@pytest.mark.skipif(O.reference("strict") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [2, 17, 1000, 25000])
def test_bit_identical_to_compiled_reference(n):
    r = empty_bodies(n)
    O.reference("strict").ref_uniform_disc(r.ctypes.data, n)
    m = ic.reference_disc(n)
    for f in ("pos", "vel", "mass", "radius"):
        assert np.array_equal(bits(r[f]), bits(m[f])), f


@pytest.mark.gpu
def test_reference_scene_step_on_gpu_matches_oracle():
    """Simulation::iterate on the reference's own scene (n = 5000 keeps the CPU oracle quick):
    BH theta=1, eps=1, dt=0.01, clamp + soft boundary.  collide() is excluded (radii zeroed)."""
    b = ic.reference_disc(5000)
    b["radius"] = 0
    want = b.copy()
    for _ in range(5):
        want["acc"] = O.orc_bh_acc(want, 1.0, 1.0)
        O.oracle().orc_iterate_after_attract(want.ctypes.data, want.shape[0], 0.01, 3, 2)
    with Simulation(b, dt=0.01, force_algo=capi.FORCE_BARNES_HUT, dims=2, theta=1.0, eps=1.0,
                    rsqrt_mode=capi.RSQRT_REFCOMPAT, integ_flags=capi.INTEG_CLAMP | capi.INTEG_BOUNDARY) as s:
        s.step(5)
        out = s.bodies
    assert np.array_equal(bits(out["acc"]), bits(want["acc"]))
    # bodies beyond the soft boundary go through expf (1-2 ulp between libm and CUDA): tolerance there
    np.testing.assert_allclose(out["vel"], want["vel"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(out["pos"], want["pos"], rtol=2e-6, atol=1e-3)
    inside = (b["pos"].astype(np.float64) ** 2).sum(1) < (0.79e5) ** 2
    assert np.array_equal(bits(out["pos"][inside]), bits(want["pos"][inside]))
