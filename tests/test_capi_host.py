"""CPU-only checks of the boundary: both shared libraries load, export every symbol the headers
declare, structs have the declared layout, and the plain-C host helpers behave.  No GPU compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nbodysim_b200 import BODY_DTYPE, capi, ic
from nbodysim_b200.bodies import empty_bodies

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(nbody_[a-z0-9_]+)\s*\(", src))


def test_gpu_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.GPU_LIB_PATH)  # loads without a GPU: cudart is linked statically, NCCL is dlopen'ed
    declared = _declared("nbody_gpu.h")
    assert declared == set(capi.GPU_SYMBOLS), declared ^ set(capi.GPU_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_host_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.HOST_LIB_PATH)
    declared = _declared("nbody_host.h")
    assert declared == set(capi.HOST_SYMBOLS), declared ^ set(capi.HOST_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_body_record_layout():
    assert BODY_DTYPE.itemsize == 64
    off = {k: BODY_DTYPE.fields[k][1] for k in BODY_DTYPE.names}
    assert (off["pos"], off["vel"], off["acc"], off["mass"], off["radius"]) == (0, 16, 32, 48, 52)
    assert (off["pos_z"], off["vel_z"], off["acc_z"]) == (8, 24, 40)


def test_params_default_are_the_reference_literals():
    p = capi.NbodyParams()
    capi.gpu_lib().nbody_params_default(C.byref(p))
    assert p.struct_size == C.sizeof(capi.NbodyParams)
    assert (p.dims, p.eps, p.G, p.theta) == (2, 1.0, 1.0, 1.0)          # Simulation.hpp:59, Vec2
    assert (p.max_velocity, p.boundary_radius) == (1000.0, 100000.0)     # Simulation.hpp:120,124
    assert abs(p.soft_boundary - 0.8) < 1e-7 and abs(p.boundary_force - 0.9) < 1e-7
    assert abs(p.damping - 0.9995) < 1e-7
    assert p.precision == capi.PRECISION_F32 and p.rsqrt_mode == capi.RSQRT_FAST


def test_error_paths_without_gpu():
    lib = capi.gpu_lib()
    assert lib.nbody_gpu_strerror(capi.EINVAL) == b"invalid argument"
    assert b"sm_100a" in lib.nbody_gpu_version()
    ctx = C.c_void_p()
    p = capi.NbodyParams()
    lib.nbody_params_default(C.byref(p))
    b = empty_bodies(4)
    assert lib.nbody_gpu_init(C.byref(ctx), C.byref(p), b.ctypes.data, 0) == capi.EINVAL
    p.dims = 5
    assert lib.nbody_gpu_init(C.byref(ctx), C.byref(p), b.ctypes.data, 4) == capi.EINVAL
    for field, bad in (("exchange", 3), ("sort_impl", -1), ("bh_walk", 7), ("theta", -1.0), ("force_algo", 9), ("precision", 4)):
        lib.nbody_params_default(C.byref(p))
        setattr(p, field, bad)
        assert lib.nbody_gpu_init(C.byref(ctx), C.byref(p), b.ctypes.data, 4) == capi.EINVAL, field
        assert lib.nbody_gpu_last_error(None) != b""
    # the cross-process peer exchange keeps tables of NBODY_MAX_GPUS entries: larger worlds must ask for NCCL
    lib.nbody_params_default(C.byref(p))
    p.world, p.rank, p.exchange = 17, 3, 0
    assert lib.nbody_gpu_init(C.byref(ctx), C.byref(p), b.ctypes.data, 4) == capi.EINVAL
    assert b"NBODY_MAX_GPUS" in lib.nbody_gpu_last_error(None)
    assert lib.nbody_gpu_step(None, 0.01, 1) == capi.EINVAL
    assert lib.nbody_gpu_download(None, b.ctypes.data, 4, 7) == capi.EINVAL
    lib.nbody_gpu_shutdown(None)  # must be a no-op


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(capi, "_gpu", None)
    monkeypatch.setattr(capi, "GPU_LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(capi.NativeLibraryMissing):
        capi.gpu_lib()


# ---- host helpers -------------------------------------------------------------------------------
def test_ic_deterministic_and_normalised():
    for gen in (lambda s: ic.plummer(4096, seed=s), lambda s: ic.uniform_sphere(4096, seed=s, virial=0.5),
                lambda s: ic.two_galaxy(4096, seed=s)):
        a, b, c = gen(1), gen(1), gen(2)
        assert a.tobytes() == b.tobytes() and a.tobytes() != c.tobytes()
        assert abs(a["mass"].astype(np.float64).sum() - 1.0) < 1e-5
        assert not a["radius"].any() and np.isfinite(a["pos"]).all() and np.isfinite(a["vel"]).all()


def test_plummer_is_virialised():
    import oracle_lib as O

    b = ic.plummer(2048, seed=11, dims=3)
    K, W, P = O.orc_energy(b, 0.0, dims=3)
    assert 0.35 < 2 * K / abs(W) / 2 < 0.65        # 2K/|W| ~ 1 for the untruncated model
    assert np.abs(P).max() < 1e-6
    r = np.sqrt((b["pos"] ** 2).sum(1) + b["pos_z"] ** 2)
    assert r.max() <= 10 * 3 * np.pi / 16 + 0.2


def test_planar_variants_have_zero_z():
    for b in (ic.plummer(512, dims=2), ic.uniform_sphere(512, dims=2, virial=0.5), ic.two_galaxy(512, dims=2)):
        assert not b["pos_z"].any() and not b["vel_z"].any()


def test_uniform_sphere_virial_ratio():
    import oracle_lib as O

    b = ic.uniform_sphere(2048, seed=4, dims=3, virial=0.5)
    K, W, _ = O.orc_energy(b, 0.0, dims=3)
    assert abs(2 * K / abs(W) - 0.5) < 0.05


def test_snapshot_roundtrip(tmp_path):
    lib = capi.host_lib()
    b = ic.plummer(300, seed=3)
    hdr = np.zeros(64, dtype=np.uint8)
    hdr[12:16] = np.frombuffer(np.uint32(3).tobytes(), dtype=np.uint8)       # dims
    hdr[16:24] = np.frombuffer(np.uint64(300).tobytes(), dtype=np.uint8)     # n
    path = str(tmp_path / "s.nbody").encode()
    assert lib.nbody_snapshot_write(path, hdr.ctypes.data, b.ctypes.data) == 0
    assert os.path.getsize(path) == 64 + 300 * 64
    h2 = np.zeros(64, dtype=np.uint8)
    b2 = empty_bodies(300)
    assert lib.nbody_snapshot_read(path, h2.ctypes.data, b2.ctypes.data, 300) == 0
    assert b2.tobytes() == b.tobytes() and bytes(h2[:7]) == b"NBODYB2"
    assert lib.nbody_snapshot_read(path, h2.ctypes.data, b2.ctypes.data, 299) != 0   # capacity too small


@pytest.mark.parametrize("n,world", [(1, 1), (1024, 1), (1025, 2), (65536, 8), (4194304, 8), (1000, 3)])
def test_shard_plan_covers_everything_once(n, world):
    spans = [ic.shard_plan(n, world, r) for r in range(world)]
    npad = spans[0][0]
    assert npad >= n and npad % (2048 * world) == 0 and npad - n < 2048 * world
    pos = 0
    for (np_, start, count) in spans:
        assert np_ == npad and start == pos and count == npad // world and count % 2048 == 0
        pos += count
    assert pos == npad


def test_streamk_partition_arithmetic():
    """the stream-K force kernel and the integrator agree on who owns which (tile, stage) unit through one piece of integer
    arithmetic (sk_owner): every unit has exactly one owner, owners are non-decreasing in unit order and every CTA owns a
    non-empty contiguous run of nearly equal length, so the CTAs that share a tile are consecutive -- its partial slots
    0 .. ns-1 are all written -- and no tile needs more slots than the library reserves"""
    lib = capi.gpu_lib()
    rng = np.random.default_rng(5)
    cases = [(512, 2048, 148), (8, 16, 128), (1, 1, 1), (13, 7, 91), (49, 98, 592), (3, 1000, 148), (2048, 3, 148)]
    cases += [(int(t), int(s), int(g)) for t, s, g in zip(rng.integers(1, 300, 20), rng.integers(1, 300, 20), rng.integers(1, 700, 20))]
    for tiles, stages, ctas in cases:
        U = tiles * stages
        ctas = min(ctas, U)
        owner = np.array([lib.nbody_gpu_streamk_owner(u, U, ctas) for u in range(U)]) if U <= 20000 else None
        worst = lib.nbody_gpu_streamk_slots(tiles, stages, ctas)
        assert worst >= 1
        if owner is None:
            continue
        assert owner[0] == 0 and owner[-1] == ctas - 1 and (np.diff(owner) >= 0).all() and (np.diff(owner) <= 1).all()
        runs = np.bincount(owner, minlength=ctas)
        assert runs.min() >= 1 and runs.max() - runs.min() <= 1                  # equal runs to within one unit
        starts = np.array([(k * U) // ctas for k in range(ctas)])              # the kernel's sk_start
        assert np.array_equal(np.searchsorted(owner, np.arange(ctas)), starts)
        ns = owner[stages - 1::stages] - owner[0::stages] + 1                     # slots per tile
        assert ns.max() == worst
    assert lib.nbody_gpu_streamk_slots(4, 4, 17) == capi.EINVAL and lib.nbody_gpu_streamk_owner(5, 4, 2) == capi.EINVAL
