"""ORACLE (test infrastructure): ctypes access to oracle/_build/libport.so, the
plain-C restatement of the reference's hit path and sample stream (oracle/port.c).
Only tests/, smoke() and bench.py's CPU-baseline legs may import this module."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libport.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_build/libport.so is missing: run `make -C oracle port`")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        vp, u32 = C.c_void_p, C.c_uint32
        L.port_trace_closest.argtypes = [vp, vp, C.c_size_t, vp]
        L.port_trace_any.argtypes = [vp, vp, C.c_size_t, vp]
        L.port_pixel_permutations.argtypes = [u32, u32, u32, u32, u32, vp]
        L.port_cmj_1d.restype = C.c_float
        L.port_cmj_1d.argtypes = [u32, u32, u32]
        L.port_cmj_2d.argtypes = [u32, u32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.port_camera_ray.argtypes = [vp, u32, u32, u32, u32, u32, u32, u32, vp]
        L.port_work_reset.argtypes = []
        L.port_work_get.argtypes = [vp]
        _lib = L
    return _lib


def trace_closest(desc, rays, hitex_dtype):
    rays = np.ascontiguousarray(rays)
    out = np.zeros(len(rays), hitex_dtype)
    lib().port_trace_closest(C.cast(desc, C.c_void_p), rays.ctypes.data, len(rays), out.ctypes.data)
    return out


def trace_any(desc, rays):
    rays = np.ascontiguousarray(rays)
    out = np.zeros(len(rays), np.uint8)
    lib().port_trace_any(C.cast(desc, C.c_void_p), rays.ctypes.data, len(rays), out.ctypes.data)
    return out


def pixel_permutations(width, height, depth, x, y):
    out = np.zeros(5 * depth + 3, np.uint32)
    rc = lib().port_pixel_permutations(width, height, depth, x, y, out.ctypes.data)
    return out if rc == 0 else None


def camera_ray(camera, width, height, ps, depth, x, y, psi, ray_dtype):
    out = np.zeros(1, ray_dtype)
    lib().port_camera_ray(C.byref(camera), width, height, ps, depth, x, y, psi, out.ctypes.data)
    return out[0]


WORK_FIELDS = ("node_pops", "tri_tests", "shape_tests", "xform_evals", "xform_keyed", "xform_pairs",
               "max_stack_mesh", "max_stack_top")


def work_reset():
    lib().port_work_reset()


def work_counters():
    """Counters accumulated by port_trace_* since the last work_reset() (oracle/port.c, 'work counters')."""
    out = np.zeros(8, np.uint64)
    lib().port_work_get(out.ctypes.data)
    return dict(zip(WORK_FIELDS, (int(v) for v in out)))
