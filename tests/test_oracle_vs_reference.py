"""The plain-C oracle against the UNMODIFIED reference compiled from /root/reference
(oracle/_ref/libnbody_ref_strict.so) on fresh seeded inputs -- bit-exact.  Skipped only when the
harness has not been built (it is built wherever /root/reference exists and travels prebuilt)."""
import numpy as np
import pytest

import oracle_lib as O
from nbodysim_b200 import ic

pytestmark = pytest.mark.skipif(O.reference("strict") is None, reason="oracle/_ref not built")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_layout_matches_reference():
    R = O.reference("strict")
    assert R.ref_sizeof_body() == 64 and R.ref_sizeof_node() == 128


def test_fast_inv_sqrt_random():
    R = O.reference("strict")
    rng = np.random.default_rng(1)
    xs = np.exp(rng.uniform(-40, 40, 5000)).astype(np.float32)
    for x in xs:
        assert O.oracle().orc_fast_inv_sqrt(float(x)) == R.ref_fast_inv_sqrt(float(x))


@pytest.mark.parametrize("n,eps,seed", [(1, 1.0, 1), (2, 0.0, 2), (97, 0.3, 3), (1024, 1.0, 4), (3000, 0.01, 5)])
def test_allpairs_bitexact(n, eps, seed):
    b = ic.plummer(n, seed=seed, dims=2)
    if n > 2:
        b["mass"] *= np.random.default_rng(seed).uniform(0.1, 4.0, n).astype(np.float32)
    assert np.array_equal(bits(O.orc_acc(b, eps)), bits(O.ref_acc(b, eps)))


def test_allpairs_subrange_and_threads():
    b = ic.spinning_disc(2048, seed=9)
    full = O.ref_acc(b, 1.0, nthreads=1)
    assert np.array_equal(bits(full), bits(O.ref_acc(b, 1.0, nthreads=8)))
    assert np.array_equal(bits(full[100:700]), bits(O.orc_acc(b, 1.0, i0=100, i1=700)))


def test_trajectory_bitexact():
    b = ic.spinning_disc(512, seed=77)
    o = O.orc_step_clean(b, 1.0, 0.01, 50)
    r = O.ref_step_clean(b, 1.0, 0.01, 50)
    for f in ("pos", "vel", "acc"):
        assert np.array_equal(bits(o[f]), bits(r[f])), f


def test_fast_build_is_close_but_not_the_pin():
    """the reference's own flags (-O3 -ffast-math) change results slightly; documented, not a pin"""
    if O.reference("fast") is None:
        pytest.skip("fast build absent")
    b = ic.spinning_disc(1024, seed=3)
    s, f = O.ref_acc(b, 1.0, kind="strict"), O.ref_acc(b, 1.0, kind="fast")
    rel = np.linalg.norm(s - f, axis=1) / np.linalg.norm(s, axis=1)
    assert np.percentile(rel, 99) < 1e-4
