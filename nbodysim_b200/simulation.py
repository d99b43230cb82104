"""Host-side mirror of the reference's `Simulation` for the hot path, over the C ABI.

Reference interface being mirrored (Nbodysim/headers/Simulation.hpp:49-75):
    class Simulation { float dt; size_t frame; std::vector<Body> bodies; ...; void step(); }
with `step()` = `iterate(SIMULATION_DT)` (+ collide, out of scope) + `++frame`, and the private
`attract()` (:176-214) that fills `bodies[i].acc`.  Same names, same meaning; the state lives on the
GPU between calls and `bodies` downloads it (the analogue of `SHARED_BODIES = simulation->bodies`,
main.cpp:625).  Everything here goes through `libnbody_gpu.so`; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from .bodies import BODY_DTYPE


class NbodyError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = capi.gpu_lib().nbody_gpu_strerror(code).decode()
        super().__init__(f"nbody_gpu error {code} ({msg}){': ' + detail if detail else ''}")


def default_params(**overrides) -> capi.NbodyParams:
    """nbody_params_default() (the reference's shipped literals) with keyword overrides."""
    p = capi.NbodyParams()
    capi.gpu_lib().nbody_params_default(C.byref(p))
    for k, v in overrides.items():
        if k == "device_ids":
            for i, d in enumerate(v):
                p.device_ids[i] = d
        elif k == "nccl_id":
            C.memmove(p.nccl_id, bytes(v), capi.NBODY_NCCL_ID_BYTES)
        elif not hasattr(p, k):
            raise TypeError(f"unknown nbody_params field {k!r}")
        else:
            setattr(p, k, v)
    return p


class Simulation:
    """GPU-resident simulation.  `Simulation(bodies, eps=..., dims=...)`; `step()`; `.bodies`."""

    def __init__(self, bodies: np.ndarray, dt: float = 0.01, **params):
        if bodies.dtype != BODY_DTYPE:
            raise TypeError("bodies must have nbodysim_b200.BODY_DTYPE (the reference's 64-byte Body)")
        bodies = np.ascontiguousarray(bodies)
        self._lib = capi.gpu_lib()
        self._params = default_params(**params)
        self._n = int(bodies.shape[0])
        self._host = bodies.copy()          # host mirror; refreshed by download()
        self._ctx = C.c_void_p()
        self.dt = float(dt)                 # SIMULATION_DT analogue (main.cpp:39), re-read every step()
        self.frame = 0                      # Simulation::frame
        rc = self._lib.nbody_gpu_init(C.byref(self._ctx), C.byref(self._params), self._host.ctypes.data, self._n)
        if rc != 0:
            raise NbodyError(rc, self._lib.nbody_gpu_last_error(None).decode())

    # -- error plumbing ----------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise NbodyError(rc, self._lib.nbody_gpu_last_error(self._ctx).decode())

    # -- the reference's surface -------------------------------------------------------------------
    def step(self, nsteps: int = 1):
        """Simulation::step(): iterate(dt) then ++frame (collide() is outside the hot path)."""
        self._check(self._lib.nbody_gpu_step(self._ctx, self.dt, int(nsteps)))
        self.frame += int(nsteps)

    def attract(self):
        """Simulation::attract(): accelerations at the current positions, no integration."""
        self._check(self._lib.nbody_gpu_accel_only(self._ctx))

    @property
    def bodies(self) -> np.ndarray:
        """Current state as Body records (downloads pos, vel, acc; synchronises)."""
        return self.download()

    # -- explicit transfers ------------------------------------------------------------------------
    def download(self, fields: int = capi.FIELD_ALL, out: np.ndarray = None) -> np.ndarray:
        dst = self._host if out is None else out
        if dst.dtype != BODY_DTYPE or dst.shape[0] != self._n or not dst.flags.c_contiguous:
            raise TypeError("download target must be a contiguous BODY_DTYPE array of n records")
        self._check(self._lib.nbody_gpu_download(self._ctx, dst.ctypes.data, self._n, fields))
        return dst

    def upload(self, bodies: np.ndarray):
        if bodies.dtype != BODY_DTYPE or bodies.shape[0] != self._n or not bodies.flags.c_contiguous:
            raise TypeError("upload source must be a contiguous BODY_DTYPE array of n records")
        self._check(self._lib.nbody_gpu_upload(self._ctx, bodies.ctypes.data, self._n))

    def download_f64(self, pos=True, vel=True, acc=True):
        """(pos, vel, acc) as (n,3) float64 arrays (None where not requested)."""
        outs = [np.zeros((self._n, 3)) if want else None for want in (pos, vel, acc)]
        ptrs = [o.ctypes.data if o is not None else None for o in outs]
        self._check(self._lib.nbody_gpu_download_f64(self._ctx, ptrs[0], ptrs[1], ptrs[2], self._n))
        return tuple(outs)

    def sync(self):
        self._check(self._lib.nbody_gpu_sync(self._ctx))

    def energy(self):
        """(K, W, P[3]) evaluated in fp64 on the device."""
        K, W, P = C.c_double(), C.c_double(), (C.c_double * 3)()
        self._check(self._lib.nbody_gpu_energy(self._ctx, C.byref(K), C.byref(W), P))
        return K.value, W.value, np.array(list(P))

    def collide(self):
        """Simulation::collide(): one collision pass on the current state."""
        self._check(self._lib.nbody_gpu_collide(self._ctx))

    def collide_stats(self):
        """(candidate pairs, resolved pairs) of the last collision pass."""
        a, b = C.c_uint32(), C.c_uint32()
        self._check(self._lib.nbody_gpu_collide_stats(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bh_nodes(self):
        """Barnes-Hut node array of the last tree built, in walk order:
        (f8[n,8] = x, y, z, mass, cx, cy, cz, size ; next[n] ; depth[n] ; is_leaf[n])."""
        cnt = C.c_size_t()
        self._check(self._lib.nbody_gpu_bh_nodes(self._ctx, None, None, 0, C.byref(cnt)))
        f8 = np.zeros((cnt.value, 8), dtype=np.float32)
        u2 = np.zeros((cnt.value, 2), dtype=np.uint32)
        self._check(self._lib.nbody_gpu_bh_nodes(self._ctx, f8.ctypes.data, u2.ctypes.data, cnt.value, C.byref(cnt)))
        return f8, u2[:, 0].copy(), (u2[:, 1] & 255).astype(np.uint32), (u2[:, 1] >> 8).astype(bool)

    def profile_next_step(self, enable=True):
        self._check(self._lib.nbody_gpu_profile_next_step(self._ctx, 1 if enable else 0))

    def info(self) -> dict:
        inf = capi.NbodyInfo()
        self._check(self._lib.nbody_gpu_get_info(self._ctx, C.byref(inf)))
        return inf.as_dict()

    @property
    def n(self):
        return self._n

    def close(self):
        if self._ctx:
            self._lib.nbody_gpu_shutdown(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
