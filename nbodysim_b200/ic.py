"""Seeded initial conditions -- thin wrappers over the C generators in host/nbody_ic.c
(include/nbody_host.h), so the oracle, the C driver and the GPU path all consume the same bytes."""
import ctypes as C

import numpy as np

from ._capi import host_lib
from .bodies import BODY_DTYPE, empty_bodies


def _check(rc, what):
    if rc != 0:
        raise ValueError(f"{what} failed with code {rc}")


def uniform_sphere(n, seed=1, dims=3, virial=0.0):
    b = empty_bodies(n)
    _check(host_lib().nbody_ic_uniform_sphere(b.ctypes.data, n, seed, dims, float(virial)), "uniform_sphere")
    return b


def plummer(n, seed=1, dims=3):
    b = empty_bodies(n)
    _check(host_lib().nbody_ic_plummer(b.ctypes.data, n, seed, dims), "plummer")
    return b


def two_galaxy(n, seed=1, dims=3):
    b = empty_bodies(n)
    _check(host_lib().nbody_ic_two_galaxy(b.ctypes.data, n, seed, dims), "two_galaxy")
    return b


def spinning_disc(n, seed=12345, scale=100.0, spin=0.3, mass=1.0):
    """SURVEY.md section 4 KAT disc in reference units (use with eps=1)."""
    b = empty_bodies(n)
    _check(host_lib().nbody_ic_spinning_disc(b.ctypes.data, n, seed, scale, spin, mass), "spinning_disc")
    return b


def reference_disc(n=25000):
    """The reference's own scene, Simulation::uniform_disc (Simulation.hpp:347-603), bit-identical."""
    b = empty_bodies(n)
    _check(host_lib().nbody_ic_reference_disc(b.ctypes.data, n), "reference_disc")
    return b


def rescale(b, lscale=1.0, vscale=1.0, mscale=1.0):
    assert b.dtype == BODY_DTYPE and b.flags.c_contiguous
    host_lib().nbody_ic_rescale(b.ctypes.data, b.shape[0], lscale, vscale, mscale)
    return b


TARGET_GRANULE = 2048  # == nb::TARGET_GRANULE (csrc/kernels.h): 8 blocks of 256 targets per force CTA


def shard_plan(n, world, rank, granule=TARGET_GRANULE):
    """(n_padded, start, count) of the multi-GPU driver's target shard for `rank`."""
    npad, start, count = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _check(host_lib().nbody_shard_plan(n, world, rank, granule, C.byref(npad), C.byref(start), C.byref(count)),
           "shard_plan")
    return npad.value, start.value, count.value
