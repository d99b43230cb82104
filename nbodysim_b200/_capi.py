"""ctypes binding of the C ABI declared in include/nbody_gpu.h and include/nbody_host.h.

The product path is the CUDA shared library ``libnbody_gpu.so`` built in-tree by ``make lib``
(``__graft_entry__.build()``).  There is NO CPU fallback: if the library is missing, importing
the GPU entry points raises, loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB_PATH = os.path.join(_HERE, "libnbody_gpu.so")
HOST_LIB_PATH = os.path.join(_HERE, "libnbody_host.so")

NBODY_MAX_GPUS = 16
NBODY_NCCL_ID_BYTES = 128

# error codes
OK, EINVAL, ECUDA, ENOMEM, ENCCL, ENODEV, ESTATE = 0, -1, -2, -3, -4, -5, -6
PRECISION_F32, PRECISION_F64 = 0, 1
RSQRT_FAST, RSQRT_REFCOMPAT = 0, 1
FORCE_ALLPAIRS, FORCE_BARNES_HUT = 0, 1
INTEG_CLAMP, INTEG_BOUNDARY = 1, 2
FIELD_POS, FIELD_VEL, FIELD_ACC, FIELD_ALL = 1, 2, 4, 7


class NbodyParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("dims", C.c_int32),
        ("eps", C.c_float),
        ("G", C.c_float),
        ("precision", C.c_int32),
        ("rsqrt_mode", C.c_int32),
        ("force_algo", C.c_int32),
        ("theta", C.c_float),
        ("integ_flags", C.c_uint32),
        ("max_velocity", C.c_float),
        ("boundary_radius", C.c_float),
        ("soft_boundary", C.c_float),
        ("boundary_force", C.c_float),
        ("damping", C.c_float),
        ("j_splits", C.c_int32),
        ("fuse_integrator", C.c_int32),
        ("use_graph", C.c_int32),
        ("force_variant", C.c_int32),
        ("collide", C.c_int32),
        ("bh_fix_near_leaves", C.c_int32),
        ("sort_impl", C.c_int32),
        ("bh_walk", C.c_int32),
        ("exchange", C.c_int32),
        ("ngpus", C.c_int32),
        ("device_ids", C.c_int32 * NBODY_MAX_GPUS),
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("nccl_id", C.c_uint8 * NBODY_NCCL_ID_BYTES),
        ("stream", C.c_void_p),
    ]


class NbodyInfo(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("n_padded", C.c_uint64),
        ("shard_start", C.c_uint64),
        ("shard_count", C.c_uint64),
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("ngpus_local", C.c_int32),
        ("p2p_exchange", C.c_int32),
        ("sm_count", C.c_int32),
        ("sm_clock_khz", C.c_int32),
        ("j_splits", C.c_int32),
        ("force_ctas", C.c_int32),
        ("ctas_per_sm", C.c_int32),
        ("fused", C.c_int32),
        ("uniform_mass", C.c_int32),
        ("graph", C.c_int32),
        ("bh_nodes", C.c_uint32),
        ("streamk_ctas", C.c_uint32),
        ("kernel_launches", C.c_uint64),
        ("interactions", C.c_uint64),
        ("last_force_ms", C.c_float),
        ("last_integ_ms", C.c_float),
        ("last_bh_build_ms", C.c_float),
        ("last_collide_ms", C.c_float),
        ("last_bh_visits", C.c_uint64),
        ("last_bh_visits_max", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/nbody_gpu.h declares: name -> (restype, argtypes)
GPU_SYMBOLS = {
    "nbody_params_default": (None, [C.POINTER(NbodyParams)]),
    "nbody_gpu_init": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(NbodyParams), C.c_void_p, C.c_size_t]),
    "nbody_gpu_step": (C.c_int, [C.c_void_p, C.c_float, C.c_int]),
    "nbody_gpu_accel_only": (C.c_int, [C.c_void_p]),
    "nbody_gpu_sync": (C.c_int, [C.c_void_p]),
    "nbody_gpu_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint]),
    "nbody_gpu_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nbody_gpu_download_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "nbody_gpu_energy": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nbody_gpu_profile_next_step": (C.c_int, [C.c_void_p, C.c_int]),
    "nbody_gpu_get_info": (C.c_int, [C.c_void_p, C.POINTER(NbodyInfo)]),
    "nbody_gpu_collide": (C.c_int, [C.c_void_p]),
    "nbody_gpu_collide_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "nbody_gpu_bh_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "nbody_gpu_streamk_owner": (C.c_int, [C.c_longlong, C.c_longlong, C.c_int]),
    "nbody_gpu_streamk_slots": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "nbody_gpu_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "nbody_gpu_shutdown": (None, [C.c_void_p]),
    "nbody_gpu_strerror": (C.c_char_p, [C.c_int]),
    "nbody_gpu_last_error": (C.c_char_p, [C.c_void_p]),
    "nbody_gpu_version": (C.c_char_p, []),
}

# every symbol include/nbody_host.h declares
HOST_SYMBOLS = {
    "nbody_rng_seed": (None, [C.c_void_p, C.c_uint64]),
    "nbody_rng_uniform": (C.c_double, [C.c_void_p]),
    "nbody_rng_normal": (C.c_double, [C.c_void_p]),
    "nbody_ic_uniform_sphere": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_double]),
    "nbody_ic_plummer": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int]),
    "nbody_ic_two_galaxy": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int]),
    "nbody_ic_spinning_disc": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_float, C.c_float, C.c_float]),
    "nbody_ic_reference_disc": (C.c_int, [C.c_void_p, C.c_size_t]),
    "nbody_ic_rescale": (None, [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float]),
    "nbody_snapshot_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p]),
    "nbody_snapshot_read_header": (C.c_int, [C.c_char_p, C.c_void_p]),
    "nbody_snapshot_read": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "nbody_shard_plan": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
}


class NativeLibraryMissing(ImportError):
    pass


def _load(path, symbols, what):
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{what} not built: {path} is missing. Run `make` (or __graft_entry__.build()); "
            "there is no CPU fallback for the CUDA path."
        )
    lib = C.CDLL(path)
    for name, (res, args) in symbols.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_gpu = None
_host = None


def gpu_lib():
    """The CUDA product library (loaded lazily so that pure-host tools do not need it)."""
    global _gpu
    if _gpu is None:
        _gpu = _load(GPU_LIB_PATH, GPU_SYMBOLS, "CUDA library libnbody_gpu.so")
    return _gpu


def host_lib():
    global _host
    if _host is None:
        _host = _load(HOST_LIB_PATH, HOST_SYMBOLS, "host library libnbody_host.so")
    return _host
