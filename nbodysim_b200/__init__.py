"""nbodysim_b200 -- B200-native hot path of 7IBBE77S/nbodysim (force accumulation + kick-drift).

Only what the path needs: `csrc/` (hand-written sm_100a kernels + the C ABI of include/nbody_gpu.h),
a ctypes binding of that ABI, the host-side mirror of the reference's `Simulation` and the seeded
initial-condition generators.  Importing this package does not load CUDA; constructing a
`Simulation` does, and fails loudly if `libnbody_gpu.so` has not been built.
"""
from . import _capi as capi
from .bodies import BODY_DTYPE, acc3, empty_bodies, make_bodies, pos3, vel3
from .simulation import NbodyError, Simulation, default_params

__all__ = ["capi", "BODY_DTYPE", "empty_bodies", "make_bodies", "pos3", "vel3", "acc3",
           "Simulation", "NbodyError", "default_params"]
__version__ = "0.1.0"
