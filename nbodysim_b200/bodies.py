"""Body-state record shared with the C ABI.

`BODY_DTYPE` is byte-for-byte the reference's ``struct alignas(16) Body``
(Nbodysim/headers/Body.hpp:6-14; ``Vec2`` is 16 bytes with 8 bytes of tail padding,
Nbodysim/headers/Vec2.hpp:17-20) == ``nbody_body_t`` in include/nbody_body.h: 64 bytes,
pos@0 vel@16 acc@32 mass@48 radius@52.  3-D runs carry z in the first padding float of each Vec2.
"""
import numpy as np

BODY_DTYPE = np.dtype(
    {
        "names": ["pos", "pos_z", "vel", "vel_z", "acc", "acc_z", "mass", "radius"],
        "formats": [("<f4", 2), "<f4", ("<f4", 2), "<f4", ("<f4", 2), "<f4", "<f4", "<f4"],
        "offsets": [0, 8, 16, 24, 32, 40, 48, 52],
        "itemsize": 64,
    }
)
assert BODY_DTYPE.itemsize == 64


def empty_bodies(n: int) -> np.ndarray:
    """n zero-initialised Body records (padding included, so byte compares are meaningful)."""
    return np.zeros(int(n), dtype=BODY_DTYPE)


def make_bodies(pos, vel=None, mass=None, radius=None) -> np.ndarray:
    """Build Body records from (n,2) or (n,3) position/velocity arrays."""
    pos = np.asarray(pos, dtype=np.float32)
    n = pos.shape[0]
    b = empty_bodies(n)
    b["pos"] = pos[:, :2]
    if pos.shape[1] == 3:
        b["pos_z"] = pos[:, 2]
    if vel is not None:
        vel = np.asarray(vel, dtype=np.float32)
        b["vel"] = vel[:, :2]
        if vel.shape[1] == 3:
            b["vel_z"] = vel[:, 2]
    b["mass"] = 1.0 if mass is None else np.asarray(mass, dtype=np.float32)
    b["radius"] = 0.0 if radius is None else np.asarray(radius, dtype=np.float32)
    return b


def pos3(b: np.ndarray) -> np.ndarray:
    return np.concatenate([b["pos"], b["pos_z"][:, None]], axis=1)


def vel3(b: np.ndarray) -> np.ndarray:
    return np.concatenate([b["vel"], b["vel_z"][:, None]], axis=1)


def acc3(b: np.ndarray) -> np.ndarray:
    return np.concatenate([b["acc"], b["acc_z"][:, None]], axis=1)
