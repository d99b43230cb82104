"""TEST INFRASTRUCTURE: ctypes access to the checker libraries under oracle/.

  * `oracle()`      -> oracle/_build/libnbody_oracle.so, the plain-C restatement (always available;
                       built on demand with gcc).
  * `reference()`   -> oracle/_ref/libnbody_ref_strict.so, the UNMODIFIED reference headers behind
                       a C harness (built in the container where /root/reference exists; travels
                       prebuilt to the GPU box).  None if absent.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libnbody_oracle.so")
REF_STRICT_SO = os.path.join(ORACLE_DIR, "_ref", "libnbody_ref_strict.so")
REF_FAST_SO = os.path.join(ORACLE_DIR, "_ref", "libnbody_ref_fast.so")

_f, _d, _sz, _i, _vp = C.c_float, C.c_double, C.c_size_t, C.c_int, C.c_void_p


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "nbody_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return ORACLE_SO


_oracle = None
_ref = {}


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(build_oracle())
        L.orc_fast_inv_sqrt.restype = _f
        L.orc_fast_inv_sqrt.argtypes = [_f]
        L.orc_allpairs_acc.argtypes = [_vp, _sz, _f, _i, _sz, _sz, _vp]
        L.orc_allpairs_acc_f64sum.argtypes = [_vp, _sz, _f, _i, _sz, _sz, _vp]
        L.orc_body_update.argtypes = [_vp, _sz, _f, _i]
        L.orc_step_clean.argtypes = [_vp, _sz, _f, _f, _i, _i]
        L.orc_iterate_after_attract.argtypes = [_vp, _sz, _f, C.c_uint, _i]
        L.orc_exact_acc_f64.argtypes = [_vp, _sz, _d, _d, _i, _sz, _sz, _vp]
        L.orc_energy_f64.argtypes = [_vp, _sz, _d, _d, _i, _vp, _vp, _vp]
        L.orc_bh_build.restype = _sz
        L.orc_bh_build.argtypes = [_vp, _sz, C.POINTER(_vp)]
        L.orc_bh_acc.argtypes = [_vp, _sz, _f, _f, _vp, _sz, _sz, _i, _vp]
        L.orc_free.argtypes = [_vp]
        L.orc_bh3_build.restype = _sz
        L.orc_bh3_build.argtypes = [_vp, _sz, C.POINTER(_vp)]
        L.orc_bh3_acc.argtypes = [_vp, _f, _f, _vp, _sz, _sz, _i, _vp]
        L.orc_collide.restype = _sz
        L.orc_collide.argtypes = [_vp, _sz, C.POINTER(_sz)]
        L.orc_resolve.argtypes = [_vp, _sz, _sz]
        _oracle = L
    return _oracle


def reference(kind="strict"):
    """The compiled reference harness, or None when oracle/_ref is absent."""
    path = REF_STRICT_SO if kind == "strict" else REF_FAST_SO
    if kind not in _ref:
        if not os.path.exists(path):
            _ref[kind] = None
        else:
            L = C.CDLL(path)
            L.ref_fast_inv_sqrt.restype = _f
            L.ref_fast_inv_sqrt.argtypes = [_f]
            L.ref_hardware_threads.restype = C.c_uint
            L.ref_allpairs_acc.argtypes = [_vp, _sz, _f, _sz, _sz, _vp, _i]
            L.ref_step_clean.argtypes = [_vp, _sz, _f, _f, _i, _i]
            L.ref_bh_acc.restype = _sz
            L.ref_bh_acc.argtypes = [_vp, _sz, _f, _f, _vp]
            L.ref_bh_nodes.restype = _sz
            L.ref_bh_nodes.argtypes = [_vp, _sz, _f, _f, _vp, _vp, _sz]
            L.ref_iterate.argtypes = [_vp, _sz, _f, _f, _f, _i]
            L.ref_collide.argtypes = [_vp, _sz]
            L.ref_step_full.argtypes = [_vp, _sz, _f, _f, _f, _i]
            L.ref_uniform_disc.argtypes = [_vp, _sz]
            _ref[kind] = L
    return _ref[kind]


# ---- numpy-level helpers ---------------------------------------------------------------------------
def orc_acc(b, eps, dims=2, i0=0, i1=None):
    n = b.shape[0]
    i1 = n if i1 is None else i1
    out = np.zeros((i1 - i0, dims), dtype=np.float32)
    oracle().orc_allpairs_acc(b.ctypes.data, n, eps, dims, i0, i1, out.ctypes.data)
    return out


def orc_acc_f64sum(b, eps, dims=2, i0=0, i1=None):
    n = b.shape[0]
    i1 = n if i1 is None else i1
    out = np.zeros((i1 - i0, dims), dtype=np.float64)
    oracle().orc_allpairs_acc_f64sum(b.ctypes.data, n, eps, dims, i0, i1, out.ctypes.data)
    return out


def orc_exact_acc(b, eps, G=1.0, dims=3, i0=0, i1=None):
    n = b.shape[0]
    i1 = n if i1 is None else i1
    out = np.zeros((i1 - i0, 3), dtype=np.float64)
    oracle().orc_exact_acc_f64(b.ctypes.data, n, eps, G, dims, i0, i1, out.ctypes.data)
    return out


def orc_step_clean(b, eps, dt, nsteps, dims=2):
    b = b.copy()
    oracle().orc_step_clean(b.ctypes.data, b.shape[0], eps, dt, nsteps, dims)
    return b


def orc_energy(b, eps, G=1.0, dims=3):
    K, W, P = C.c_double(), C.c_double(), (C.c_double * 3)()
    oracle().orc_energy_f64(b.ctypes.data, b.shape[0], eps, G, dims, C.byref(K), C.byref(W), P)
    return K.value, W.value, np.array(list(P))


def ref_acc(b, eps, i0=0, i1=None, nthreads=0, kind="strict"):
    n = b.shape[0]
    i1 = n if i1 is None else i1
    out = np.zeros((i1 - i0, 2), dtype=np.float32)
    reference(kind).ref_allpairs_acc(b.ctypes.data, n, eps, i0, i1, out.ctypes.data, nthreads)
    return out


def ref_step_clean(b, eps, dt, nsteps, nthreads=0, kind="strict"):
    b = b.copy()
    reference(kind).ref_step_clean(b.ctypes.data, b.shape[0], eps, dt, nsteps, nthreads)
    return b


ORC_NODE_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("mass", "<f4"), ("cx", "<f4"), ("cy", "<f4"), ("size", "<f4"),
                           ("children", "<u8"), ("next", "<u8"), ("depth", "<u8")])
assert ORC_NODE_DTYPE.itemsize == 48


def orc_bh_build(b):
    """Oracle restatement of Quadtree::build -> structured array of nodes."""
    ptr = _vp()
    m = oracle().orc_bh_build(b.ctypes.data, b.shape[0], C.byref(ptr))
    buf = (C.c_char * (m * ORC_NODE_DTYPE.itemsize)).from_address(ptr.value)
    nodes = np.frombuffer(buf, dtype=ORC_NODE_DTYPE, count=m).copy()
    oracle().orc_free(ptr)
    return nodes


def orc_bh_acc(b, theta, eps, nodes=None, i0=0, i1=None, fix_near_leaves=False):
    n = b.shape[0]
    i1 = n if i1 is None else i1
    nodes = orc_bh_build(b) if nodes is None else nodes
    out = np.zeros((i1 - i0, 2), dtype=np.float32)
    oracle().orc_bh_acc(nodes.ctypes.data, nodes.shape[0], theta, eps, b.ctypes.data, i0, i1,
                        1 if fix_near_leaves else 0, out.ctypes.data)
    return out


def ref_bh_acc(b, theta, eps, kind="strict"):
    out = np.zeros((b.shape[0], 2), dtype=np.float32)
    m = reference(kind).ref_bh_acc(b.ctypes.data, b.shape[0], theta, eps, out.ctypes.data)
    return out, m


def ref_bh_nodes(b, theta=1.0, eps=1.0, kind="strict"):
    cap = 8 * b.shape[0] + 64
    f = np.zeros((cap, 6), dtype=np.float32)
    u = np.zeros((cap, 3), dtype=np.uint64)
    m = reference(kind).ref_bh_nodes(b.ctypes.data, b.shape[0], theta, eps, f.ctypes.data, u.ctypes.data, cap)
    assert m <= cap
    return f[:m], u[:m]


def orc_collide(b):
    """returns (bodies after the collision pass, broad-phase pairs, resolved pairs)"""
    b = b.copy()
    res = _sz()
    npairs = oracle().orc_collide(b.ctypes.data, b.shape[0], C.byref(res))
    return b, npairs, res.value


def ref_collide(b, kind="strict"):
    b = b.copy()
    reference(kind).ref_collide(b.ctypes.data, b.shape[0])
    return b


ORC_NODE3_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("mass", "<f4"), ("cx", "<f4"), ("cy", "<f4"),
                            ("cz", "<f4"), ("size", "<f4"), ("children", "<u8"), ("next", "<u8"), ("depth", "<u8")])
assert ORC_NODE3_DTYPE.itemsize == 56


def orc_bh3_build(b):
    ptr = _vp()
    m = oracle().orc_bh3_build(b.ctypes.data, b.shape[0], C.byref(ptr))
    buf = (C.c_char * (m * ORC_NODE3_DTYPE.itemsize)).from_address(ptr.value)
    nodes = np.frombuffer(buf, dtype=ORC_NODE3_DTYPE, count=m).copy()
    oracle().orc_free(ptr)
    return nodes


def orc_bh3_acc(b, theta, eps, nodes=None, i0=0, i1=None, fix_near_leaves=False):
    n = b.shape[0]
    i1 = n if i1 is None else i1
    nodes = orc_bh3_build(b) if nodes is None else nodes
    out = np.zeros((i1 - i0, 3), dtype=np.float32)
    oracle().orc_bh3_acc(nodes.ctypes.data, theta, eps, b.ctypes.data, i0, i1, 1 if fix_near_leaves else 0, out.ctypes.data)
    return out
