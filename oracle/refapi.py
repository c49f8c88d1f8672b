"""ORACLE (test infrastructure): ctypes access to oracle/_ref/libref_s7.so, the
UNMODIFIED reference Stage 7 render core compiled from /root/reference by
oracle/Makefile.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module; the product never does.

The library is built in the CPU container (where /root/reference exists) and
travels to the GPU box inside the repository snapshot; nothing here reads
/root/reference at run time.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_s7.so")

REFHIT_DTYPE = np.dtype([("t", "<f4"), ("shape", "<i4"), ("face", "<i4"), ("tri", "<i4"),
                         ("normal", "<f4", 3), ("color_modifier", "<f4", 3)])


class RefRenderStats(C.Structure):
    _fields_ = [("render_seconds", C.c_double), ("prepare_seconds", C.c_double),
                ("closest_calls", C.c_uint64), ("any_calls", C.c_uint64), ("threads", C.c_uint)]


LIB6_PATH = os.path.join(HERE, "_ref", "libref_s6.so")

_lib = None
_lib6 = None


def available(stage=7):
    return os.path.exists(LIB6_PATH if stage == 6 else LIB_PATH)


class _Stage6Lib:
    """libref_s6.so (the unmodified Stage 6 core) under the Stage 7 driver's names: ref_x -> ref6_x"""

    def __init__(self):
        if not available(6):
            raise RuntimeError("oracle/_ref/libref_s6.so is missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB6_PATH, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        L.ref6_scene_create.restype = vp
        L.ref6_scene_create.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint]
        L.ref6_scene_destroy.argtypes = [vp]
        L.ref6_scene_num_finite.argtypes = [vp]
        L.ref6_scene_num_infinite.argtypes = [vp]
        L.ref6_scene_prepare_seconds.restype = C.c_double
        L.ref6_scene_prepare_seconds.argtypes = [vp]
        L.ref6_trace_closest.argtypes = [vp, vp, C.c_size_t, vp]
        L.ref6_trace_any.argtypes = [vp, vp, C.c_size_t, vp]
        L.ref6_render.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, vp,
                                  C.POINTER(RefRenderStats), C.c_int]
        L.ref6_recorded_rays.restype = C.c_size_t
        L.ref6_recorded_rays.argtypes = [C.c_int, vp, C.c_size_t]
        L.ref6_bvh_nodes.restype = C.c_uint
        L.ref6_bvh_nodes.argtypes = [vp, C.c_int, vp, C.c_uint]
        L.ref6_mesh_counts.argtypes = [vp, C.c_int] + [C.POINTER(C.c_uint)] * 4
        L.ref6_mesh_data.argtypes = [vp, C.c_int] + [vp] * 7
        L.ref6_build_info.restype = C.c_char_p
        self._L = L

    def __getattr__(self, name):
        if name.startswith("ref_"):
            return getattr(self._L, "ref6_" + name[4:])
        raise AttributeError(name)


def lib(stage=7):
    global _lib, _lib6
    if stage == 6:
        if _lib6 is None:
            _lib6 = _Stage6Lib()
        return _lib6
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_s7.so is missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        L.ref_scene_create.restype = vp
        L.ref_scene_create.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint]
        L.ref_scene_destroy.argtypes = [vp]
        L.ref_scene_num_finite.argtypes = [vp]
        L.ref_scene_num_infinite.argtypes = [vp]
        L.ref_scene_prepare_seconds.restype = C.c_double
        L.ref_scene_prepare_seconds.argtypes = [vp]
        L.ref_trace_closest.argtypes = [vp, vp, C.c_size_t, vp]
        L.ref_trace_any.argtypes = [vp, vp, C.c_size_t, vp]
        L.ref_render.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, vp,
                                 C.POINTER(RefRenderStats), C.c_int]
        L.ref_recorded_rays.restype = C.c_size_t
        L.ref_recorded_rays.argtypes = [C.c_int, vp, C.c_size_t]
        L.ref_bvh_nodes.restype = C.c_uint
        L.ref_bvh_nodes.argtypes = [vp, C.c_int, vp, C.c_uint]
        L.ref_mesh_counts.argtypes = [vp, C.c_int] + [C.POINTER(C.c_uint)] * 4
        L.ref_mesh_data.argtypes = [vp, C.c_int] + [vp] * 7
        L.ref_shape_keys.restype = C.c_uint
        L.ref_shape_keys.argtypes = [vp, C.c_int, vp, C.c_uint]
        L.ref_rng_sequence.argtypes = [C.c_uint, C.c_uint, C.c_size_t, vp]
        L.ref_cmj_1d.argtypes = [C.c_uint, C.c_uint, C.c_uint, vp]
        L.ref_cmj_2d.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint, vp]
        L.ref_camera_rays.argtypes = [vp, vp, C.c_size_t, vp]
        L.ref_build_info.restype = C.c_char_p
        _lib = L
    return _lib


class RefScene:
    """A recipe scene built against the reference's own classes and prepared once."""

    def __init__(self, recipe, obj_path=None, grid=(0, 0), stage=7):
        self.stage = stage
        self.handle = lib(stage).ref_scene_create(recipe, obj_path.encode() if obj_path else None, grid[0], grid[1])
        if not self.handle:
            raise RuntimeError("ref_scene_create failed")

    @property
    def num_finite(self):
        return lib(self.stage).ref_scene_num_finite(self.handle)

    @property
    def num_infinite(self):
        return lib(self.stage).ref_scene_num_infinite(self.handle)

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), REFHIT_DTYPE)
        lib(self.stage).ref_trace_closest(self.handle, rays.ctypes.data, len(rays), out.ctypes.data)
        return out

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), np.uint8)
        lib(self.stage).ref_trace_any(self.handle, rays.ctypes.data, len(rays), out.ctypes.data)
        return out

    def render(self, camera_spec14, width, height, ps, ls=1, depth=3, record_rays=False):
        spec = np.ascontiguousarray(camera_spec14, np.float32)
        img = np.zeros((height, width, 3), np.float32)
        stats = RefRenderStats()
        rc = lib(self.stage).ref_render(self.handle, spec.ctypes.data, width, height, ps, ls, depth, img.ctypes.data,
                              C.byref(stats), 1 if record_rays else 0)
        if rc != 0:
            raise RuntimeError("ref_render failed")
        return img, stats

    def recorded_rays(self, kind, dtype):
        n = lib(self.stage).ref_recorded_rays(kind, None, 0)
        out = np.zeros(n, dtype)
        lib(self.stage).ref_recorded_rays(kind, out.ctypes.data, n)
        return out

    def bvh_nodes(self, shape=-1):
        n = lib(self.stage).ref_bvh_nodes(self.handle, shape, None, 0)
        out = np.zeros((n, 8), np.uint32)
        if n:
            lib(self.stage).ref_bvh_nodes(self.handle, shape, out.ctypes.data, n)
        return out

    def mesh(self, shape):
        counts = [C.c_uint() for _ in range(4)]
        if lib(self.stage).ref_mesh_counts(self.handle, shape, *[C.byref(c) for c in counts]) != 0:
            return None
        nv, nn, nf, ni = [c.value for c in counts]
        m = {
            "vertices": np.zeros((nv, 3), np.float32), "normals": np.zeros((nn, 3), np.float32),
            "face_sizes": np.zeros(nf, np.uint32), "vertex_index": np.zeros(ni, np.uint32),
            "normal_index": np.zeros(ni, np.uint32), "area_cdf": np.zeros(nf + 1, np.float32),
            "bbox": np.zeros(6, np.float32),
        }
        lib(self.stage).ref_mesh_data(self.handle, shape, m["vertices"].ctypes.data, m["normals"].ctypes.data,
                            m["face_sizes"].ctypes.data, m["vertex_index"].ctypes.data,
                            m["normal_index"].ctypes.data, m["area_cdf"].ctypes.data, m["bbox"].ctypes.data)
        return m

    def shape_keys(self, shape):
        n = lib(self.stage).ref_shape_keys(self.handle, shape, None, 0)
        out = np.zeros((n, 11), np.float32)
        if n:
            lib(self.stage).ref_shape_keys(self.handle, shape, out.ctypes.data, n)
        return out

    def close(self):
        if self.handle:
            lib(self.stage).ref_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_sequence(z, w, count):
    out = np.zeros(count, np.uint32)
    lib().ref_rng_sequence(z, w, count, out.ctypes.data)
    return out


def cmj_1d(samples, permutation, count):
    out = np.zeros(count, np.float32)
    lib().ref_cmj_1d(samples, permutation, count, out.ctypes.data)
    return out


def cmj_2d(xs, ys, permutation, count):
    out = np.zeros((count, 2), np.float32)
    lib().ref_cmj_2d(xs, ys, permutation, count, out.ctypes.data)
    return out


def camera_rays(camera_spec14, args5, ray_dtype):
    spec = np.ascontiguousarray(camera_spec14, np.float32)
    args = np.ascontiguousarray(args5, np.float32)
    out = np.zeros(len(args), ray_dtype)
    lib().ref_camera_rays(spec.ctypes.data, args.ctypes.data, len(args), out.ctypes.data)
    return out


# ---- Stage 2 / Stage 3 (serial-Rng programs; oracle/ref_s2_driver.cpp, ref_s3_driver.cpp) ----
_stage_libs = {}


def stage_lib(stage):
    """libref_s2.so / libref_s3.so: the unmodified Stage 2 / Stage 3 code with a variable sample count"""
    if stage not in _stage_libs:
        path = os.path.join(HERE, "_ref", "libref_s%d.so" % stage)
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run `make -C oracle ref` where /root/reference exists" % path)
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        if stage == 2:
            L.ref2_render.argtypes = [C.c_uint, C.c_uint, C.c_uint, vp, vp, vp, vp]
            L.ref2_constants.argtypes = [vp]
            L.ref2_build_info.restype = C.c_char_p
        else:
            L.ref3_render.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp, vp]
            L.ref3_constants.argtypes = [vp]
            L.ref3_build_info.restype = C.c_char_p
        _stage_libs[stage] = L
    return _stage_libs[stage]


def stage_available(stage):
    return os.path.exists(os.path.join(HERE, "_ref", "libref_s%d.so" % stage))


def stage_binary(stage):
    """The unmodified CLI (writes out.ppm into its cwd)"""
    return os.path.join(HERE, "_ref", "stage%d" % stage)


def stage_render(stage, width, height, samples_u, samples_v=1, want_flags=False):
    """Returns (float rgb before clamp, uint8 rgb as written to out.ppm, hit flags or None, rays).
    Stage 2 takes samples_u purely random samples per pixel; Stage 3 samples_u x samples_v strata."""
    L = stage_lib(stage)
    rgb = np.zeros((height, width, 3), np.float32)
    rgb8 = np.zeros((height, width, 3), np.uint8)
    n = width * height * samples_u * (samples_v if stage == 3 else 1)
    flags = np.zeros(n, np.uint8) if want_flags else None
    rays = C.c_uint64(0)
    fp = flags.ctypes.data if want_flags else None
    if stage == 2:
        L.ref2_render(width, height, samples_u, rgb.ctypes.data, rgb8.ctypes.data, fp, C.addressof(rays))
    else:
        L.ref3_render(width, height, samples_u, samples_v, rgb.ctypes.data, rgb8.ctypes.data, fp, C.addressof(rays))
    return rgb, rgb8, flags, rays.value
