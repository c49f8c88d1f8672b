"""ctypes bindings of the two native libraries (include/rayito_b200.h,
include/rayito_b200_host.h).  Thin by design: the product is the CUDA core; Python
only moves pointers.  Loading fails loudly if the libraries are not built, and
every compute call raises when the CUDA core reports an error -- there is no
Python or CPU fallback for any of them."""
import ctypes as C
import os

import numpy as np

from . import build as _build


class RtError(RuntimeError):
    pass


class RtRay(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3), ("tmax", C.c_float), ("time", C.c_float)]


RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("tmax", "<f4"), ("time", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("shape", "<i4"), ("face", "<i4"), ("tri", "<i4")])
HITEX_DTYPE = np.dtype([("t", "<f4"), ("shape", "<i4"), ("face", "<i4"), ("tri", "<i4"),
                        ("normal", "<f4", 3), ("color_modifier", "<f4")])


class RtCamera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("forward", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3),
                ("tan_fov", C.c_float), ("focal_distance", C.c_float), ("lens_radius", C.c_float),
                ("shutter_open", C.c_float), ("shutter_close", C.c_float)]


class RtRenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("pixel_samples_hint", C.c_uint32), ("light_samples_hint", C.c_uint32),
                ("max_ray_depth", C.c_uint32), ("tile_size", C.c_uint32),
                ("rank", C.c_uint32), ("world", C.c_uint32),
                ("max_batch_samples", C.c_uint32), ("flags", C.c_uint32)]


RT_RENDER_COUNT_WORK = 1
RT_RENDER_TIME_TRACE = 2
RT_RENDER_UNIFIED_TRAVERSAL = 4
RT_RENDER_DYNAMIC_TOP = 8


class RtRenderStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("any_rays", C.c_uint64),
                ("node_pops", C.c_uint64), ("tri_tests", C.c_uint64), ("shape_tests", C.c_uint64),
                ("xform_evals", C.c_uint64), ("xform_keyed", C.c_uint64), ("xform_pairs", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("trace_launches", C.c_uint64),
                ("render_ms", C.c_float), ("trace_ms", C.c_float), ("upload_ms", C.c_float), ("download_ms", C.c_float)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class RtSceneDesc(C.Structure):
    """Opaque to Python except for a few counters used by tests (layout of
    include/rayito_b200.h RtSceneDesc)."""
    _p = C.c_void_p
    _u = C.c_uint32
    _fields_ = [("abi_version", _u), ("set_xform", _u),
                ("num_finite", _u), ("num_infinite", _u), ("shapes", _p),
                ("num_top_nodes", _u), ("top_nodes", _p),
                ("num_xforms", _u), ("xforms", _p),
                ("num_keys", _u), ("key_time", _p), ("key_scale", _p), ("key_rotation", _p), ("key_translation", _p),
                ("num_planes", _u), ("planes", _p), ("num_spheres", _u), ("spheres", _p),
                ("num_rects", _u), ("rects", _p), ("num_meshes", _u), ("meshes", _p),
                ("num_vertices", _u), ("vertices", _p), ("num_normals", _u), ("normals", _p),
                ("num_faces", _u), ("face_start", _p), ("face_has_normals", _p),
                ("num_indices", _u), ("vertex_index", _p), ("normal_index", _p),
                ("num_mesh_nodes", _u), ("mesh_nodes", _p),
                ("num_cdf", _u), ("face_area_cdf", _p),
                ("num_materials", _u), ("materials", _p),
                ("num_lights", _u), ("lights", _p), ("semantics", _u)]


class RtMesh(C.Structure):
    _u = C.c_uint32
    _fields_ = [("first_vertex", _u), ("num_vertices", _u), ("first_normal", _u), ("num_normals", _u),
                ("first_face", _u), ("num_faces", _u), ("first_node", _u), ("num_nodes", _u),
                ("first_cdf", _u), ("total_area", C.c_float)]


class RtShape(C.Structure):
    _u = C.c_uint32
    _fields_ = [("type", _u), ("geom", _u), ("xform", _u), ("material", _u), ("light", C.c_int32)]


class RtXform(C.Structure):
    _fields_ = [("first_key", C.c_uint32), ("num_keys", C.c_uint32)]


RECIPE_STAGE7_SCENE1 = 1
RECIPE_STAGE7_SCENE2 = 2
RECIPE_STAGE7_SCENE1_MESHLIGHT = 3
RECIPE_SYNTHETIC_MESH = 5
RECIPE_STAGE6_SCENE = 6
RECIPE_EDGE_LINEAR_LIST = 7
RECIPE_EDGE_NO_LIGHTS = 8
RECIPE_EDGE_EMPTY = 9
RECIPE_EDGE_DEEP_MESH = 10
RECIPE_EDGE_DEEP_BOTH = 11

# Every symbol include/rayito_b200.h declares (checked by the CPU test-suite)
CORE_SYMBOLS = [
    "rt_last_error_string", "rt_abi_version", "rt_device_count",
    "rt_scene_create", "rt_scene_destroy",
    "rt_trace_closest", "rt_trace_closest_ex", "rt_trace_any",
    "rt_trace_closest_device", "rt_trace_any_device", "rt_trace_closest_counted", "rt_trace_any_counted",
    "rt_scene_create_ex", "rt_scene_mesh_nodes",
    "rt_render", "rt_render_device", "rt_generate_camera_rays", "rt_tonemap_bgra8", "rt_tonemap_bgra8_device",
    "rt_stage1_render_float",
    "rt_tile_owners", "rt_sample_permutations", "rt_cmj_sample1d", "rt_cmj_sample2d", "rt_stage1_render",
    "rt_libm_eval", "rt_stage23_render", "rt_release_cached_memory",
    "rt_comm_unique_id", "rt_comm_create", "rt_comm_from_nccl", "rt_comm_destroy", "rt_render_multi",
    "rt_packed_floats", "rt_render_tiles_packed", "rt_unpack_tiles", "rt_render_multi_host", "rt_comm_rank",
]
HOST_SYMBOLS = [
    "rth_last_error_string", "rth_camera", "rth_stage1_render", "rth_stage1_render_float", "rth_stage23_render",
    "rth_set_tree_mode",
]
TREE_REFERENCE, TREE_SAH, TREE_DEVICE, TREE_AUTO = 0, 1, 2, 3     # rth_set_tree_mode (include/rayito_b200_host.h); AUTO is the library default
RT_SCENE_BUILD_MESH_BVH = 1                         # rt_scene_create_ex flags (include/rayito_b200.h)
# fixtures/rayito_fixtures.h (test infrastructure: the recipe scenes)
FIXTURE_SYMBOLS = [
    "rthf_last_error_string", "rth_scene_create", "rth_scene_destroy", "rth_scene_desc",
    "rth_scene_prepare_seconds", "rth_scene_depth", "rth_scene_default_camera", "rth_raytrace",
    "rth_app_create", "rth_app_destroy", "rth_app_raytrace", "rth_app_raytrace_image", "rth_app_raytrace_multi",
]

_core = None
_host = None


def core():
    """librayito_b200.so (built on demand; raises if it cannot be built or loaded)."""
    global _core
    if _core is None:
        path = _build.build_core()
        lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
        lib.rt_last_error_string.restype = C.c_char_p
        lib.rt_scene_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
        lib.rt_scene_create_ex.argtypes = [vp, C.c_int, u32, C.POINTER(vp)]
        lib.rt_scene_mesh_nodes.argtypes = [vp, u32, vp, u32, C.POINTER(u32), C.POINTER(C.c_float)]
        lib.rt_scene_destroy.argtypes = [vp]
        lib.rt_trace_closest.argtypes = [vp, vp, sz, vp]
        lib.rt_trace_closest_ex.argtypes = [vp, vp, sz, vp]
        lib.rt_trace_any.argtypes = [vp, vp, sz, vp]
        lib.rt_trace_closest_counted.argtypes = [vp, vp, sz, vp, vp]
        lib.rt_trace_any_counted.argtypes = [vp, vp, sz, vp, vp]
        lib.rt_trace_closest_device.argtypes = [vp, vp, sz, vp, vp, vp]
        lib.rt_trace_any_device.argtypes = [vp, vp, sz, vp, vp, vp]
        lib.rt_render.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp, C.POINTER(RtRenderStats)]
        lib.rt_render_device.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp,
                                         C.POINTER(RtRenderStats), vp]
        lib.rt_comm_unique_id.argtypes = [vp]
        lib.rt_comm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        lib.rt_comm_from_nccl.argtypes = [vp, C.c_int, C.POINTER(vp)]
        lib.rt_comm_destroy.argtypes = [vp]
        lib.rt_render_multi.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp, C.c_int, vp,
                                        C.POINTER(RtRenderStats), C.POINTER(C.c_float), vp]
        lib.rt_packed_floats.restype = sz
        lib.rt_packed_floats.argtypes = [u32, u32, u32, u32, u32]
        lib.rt_render_tiles_packed.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp, sz,
                                               C.POINTER(RtRenderStats), vp]
        lib.rt_unpack_tiles.argtypes = [C.c_int, vp, u32, u32, u32, u32, u32, vp, vp]
        lib.rt_generate_camera_rays.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), u32, vp]
        lib.rt_tonemap_bgra8.argtypes = [C.c_int, vp, sz, C.c_float, C.c_float, vp]
        lib.rt_tonemap_bgra8_device.argtypes = [C.c_int, vp, sz, C.c_float, C.c_float, vp, vp]
        lib.rt_libm_eval.argtypes = [C.c_int, vp, vp, sz, vp]
        lib.rt_tile_owners.argtypes = [u32, u32, u32, u32, vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        lib.rt_sample_permutations.argtypes = [u32, u32, u32, u32, u32, vp]
        lib.rt_cmj_sample1d.restype = C.c_float
        lib.rt_cmj_sample1d.argtypes = [u32, u32, u32]
        lib.rt_cmj_sample2d.argtypes = [u32, u32, u32, u32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _core = lib
    return _core


class _HostLibs:
    """librayito_host.so (product) and fixtures/librayito_fixtures.so (recipe scenes) behind one
    attribute lookup, so that callers write lib.rth_xxx whichever library holds it."""

    def __init__(self, host_lib, fixtures_lib):
        self.product = host_lib
        self.fixtures = fixtures_lib

    def __getattr__(self, name):
        # the product library first: dlsym on the fixtures handle also finds the product's symbols (a
        # dependency), and would hand back a function object without the argument types declared below
        for lib in (self.__dict__["product"], self.__dict__["fixtures"]):
            try:
                return getattr(lib, name)
            except AttributeError:
                continue
        raise AttributeError(name)

    def rth_last_error_string(self):
        """Last error texts of both libraries on this thread (each keeps its own)."""
        a, b = self.product.rth_last_error_string(), self.fixtures.rthf_last_error_string()
        return a + (b" | " if a and b else b"") + b


def host():
    """librayito_host.so (the C++ mirror of the Rayito API) plus the fixtures library."""
    global _host
    if _host is None:
        core()
        vp = C.c_void_p
        hl = C.CDLL(_build.build_host(), mode=C.RTLD_GLOBAL)
        hl.rth_last_error_string.restype = C.c_char_p
        hl.rth_camera.argtypes = [vp, C.POINTER(RtCamera)]
        hl.rth_stage1_render.argtypes = [C.c_int, C.c_uint, C.c_uint, vp]
        hl.rth_stage1_render_float.argtypes = [C.c_int, C.c_uint, C.c_uint, vp, vp]
        hl.rth_stage23_render.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp]
        hl.rth_set_tree_mode.argtypes = [C.c_uint]
        lib = C.CDLL(_build.build_fixtures(), mode=C.RTLD_LOCAL)
        lib.rthf_last_error_string.restype = C.c_char_p
        lib.rth_scene_create.restype = vp
        lib.rth_scene_create.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint]
        lib.rth_scene_destroy.argtypes = [vp]
        lib.rth_scene_desc.restype = C.POINTER(RtSceneDesc)
        lib.rth_scene_desc.argtypes = [vp]
        lib.rth_scene_prepare_seconds.restype = C.c_double
        lib.rth_scene_prepare_seconds.argtypes = [vp]
        lib.rth_scene_depth.restype = C.c_uint
        lib.rth_scene_depth.argtypes = [vp, C.c_int]
        lib.rth_scene_default_camera.argtypes = [vp, vp]
        lib.rth_raytrace.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint, vp, C.c_uint, C.c_uint,
                                     C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_int,
                                     vp, C.POINTER(RtRenderStats)]
        lib.rth_app_create.restype = vp
        lib.rth_app_create.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint]
        lib.rth_app_destroy.restype = None
        lib.rth_app_destroy.argtypes = [vp]
        lib.rth_app_raytrace.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                         C.c_int, C.c_uint, C.c_uint, C.c_int, vp, C.c_int, C.POINTER(RtRenderStats)]
        lib.rth_app_raytrace_image.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                               C.c_int, C.c_uint, C.c_uint, C.c_int, C.POINTER(C.c_void_p),
                                               C.POINTER(RtRenderStats)]
        lib.rth_app_raytrace_multi.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, vp, C.c_int,
                                               C.POINTER(C.c_void_p), C.POINTER(RtRenderStats)]
        _host = _HostLibs(hl, lib)
    return _host


def check(rc, what="rayito_b200"):
    if rc != 0:
        raise RtError("%s failed (%d): %s" % (what, rc, core().rt_last_error_string().decode()))


class HostScene:
    """A recipe scene built with the C++ host API, prepared and flattened."""

    def __init__(self, recipe, obj_path=None, grid=(0, 0), tree=TREE_REFERENCE):
        lib = host()
        path = obj_path.encode() if obj_path else None
        # tree: which face BVH prepare() builds.  The recipe scenes default to trees built on the HOST so that their
        # flattened description is complete (tests compare it, rt_scene_create takes it); TREE_AUTO is what an
        # application gets (large meshes left to the GPU: create the DeviceScene with build_bvh_on_device=True),
        # TREE_SAH the perf-mode tree (parity measured, not bit-exact)
        if lib.rth_set_tree_mode(tree) != 0:
            raise RtError("rth_set_tree_mode: " + lib.rth_last_error_string().decode())
        try:
            self.handle = lib.rth_scene_create(recipe, path, grid[0], grid[1])
        finally:
            lib.rth_set_tree_mode(TREE_AUTO)
        if not self.handle:
            raise RtError("rth_scene_create: " + lib.rth_last_error_string().decode())
        self.recipe = recipe
        self.obj_path = obj_path
        self.grid = grid

    @property
    def desc(self):
        return host().rth_scene_desc(self.handle)

    @property
    def prepare_seconds(self):
        return host().rth_scene_prepare_seconds(self.handle)

    def depth(self, mesh=-1):
        return host().rth_scene_depth(self.handle, mesh)

    def default_camera_spec(self):
        spec = np.zeros(14, np.float32)
        host().rth_scene_default_camera(self.handle, spec.ctypes.data)
        return spec

    def close(self):
        if self.handle:
            host().rth_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera_from_spec(spec14):
    spec = np.ascontiguousarray(spec14, np.float32)
    cam = RtCamera()
    if host().rth_camera(spec.ctypes.data, C.byref(cam)) != 0:
        raise RtError("rth_camera failed")
    return cam


class DeviceScene:
    """RtScene handle: a flattened scene uploaded to one GPU."""

    def __init__(self, desc, device=0, build_bvh_on_device=False):
        self.handle = C.c_void_p()
        if build_bvh_on_device:
            check(core().rt_scene_create_ex(C.cast(desc, C.c_void_p), device, RT_SCENE_BUILD_MESH_BVH, C.byref(self.handle)),
                  "rt_scene_create_ex")
        else:
            check(core().rt_scene_create(C.cast(desc, C.c_void_p), device, C.byref(self.handle)), "rt_scene_create")
        self.device = device

    def mesh_nodes(self, mesh, num_faces):
        """(nodes as uint32 [2F-1, 8] in the reference's RtBvhNode layout, deepest leaf, device build ms)"""
        nodes = np.zeros((max(2 * num_faces - 1, 0), 8), np.uint32)
        depth, ms = C.c_uint32(0), C.c_float(0.0)
        check(core().rt_scene_mesh_nodes(self.handle, mesh, nodes.ctypes.data, nodes.shape[0], C.byref(depth), C.byref(ms)),
              "rt_scene_mesh_nodes")
        return nodes, int(depth.value), float(ms.value)

    def trace_closest(self, rays, extended=False):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = np.empty(len(rays), HITEX_DTYPE if extended else HIT_DTYPE)
        fn = core().rt_trace_closest_ex if extended else core().rt_trace_closest
        check(fn(self.handle, rays.ctypes.data, len(rays), hits.ctypes.data), "rt_trace_closest")
        return hits

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = np.empty(len(rays), np.uint8)
        check(core().rt_trace_any(self.handle, rays.ctypes.data, len(rays), hits.ctypes.data), "rt_trace_any")
        return hits

    WORK_FIELDS = ("node_pops", "tri_tests", "shape_tests", "xform_evals", "xform_keyed", "xform_pairs")

    def trace_counted(self, rays, any_hit=False):
        """(hits, work counters) of one batch: rt_trace_closest_counted / rt_trace_any_counted"""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        work = np.zeros(6, np.uint64)
        if any_hit:
            hits = np.empty(len(rays), np.uint8)
            check(core().rt_trace_any_counted(self.handle, rays.ctypes.data, len(rays), hits.ctypes.data, work.ctypes.data))
        else:
            hits = np.empty(len(rays), HIT_DTYPE)
            check(core().rt_trace_closest_counted(self.handle, rays.ctypes.data, len(rays), hits.ctypes.data,
                                                  work.ctypes.data))
        return hits, dict(zip(self.WORK_FIELDS, (int(v) for v in work)))

    def render(self, camera, width, height, ps, ls=1, depth=3, rank=0, world=1, tile_size=0,
               max_batch_samples=0, count_work=False, time_trace=False, unified=False, out=None, dynamic_top=False):
        params = RtRenderParams(width, height, ps, ls, depth, tile_size, rank, world, max_batch_samples,
                                (RT_RENDER_COUNT_WORK if count_work else 0) | (RT_RENDER_TIME_TRACE if time_trace else 0)
                                | (RT_RENDER_UNIFIED_TRAVERSAL if unified else 0)
                                | (RT_RENDER_DYNAMIC_TOP if dynamic_top else 0))
        if out is None:
            out = np.zeros((height, width, 3), np.float32)
        stats = RtRenderStats()
        check(core().rt_render(self.handle, C.byref(camera), C.byref(params), out.ctypes.data, C.byref(stats)),
              "rt_render")
        return out, stats

    def render_device(self, camera, params, d_rgb_ptr, stream=None):
        stats = RtRenderStats()
        check(core().rt_render_device(self.handle, C.byref(camera), C.byref(params), d_rgb_ptr, C.byref(stats),
                                      stream), "rt_render_device")
        return stats

    def camera_rays(self, camera, width, height, ps, psi, ls=1, depth=3):
        params = RtRenderParams(width, height, ps, ls, depth, 0, 0, 1, 0, 0)
        rays = np.zeros(width * height, RAY_DTYPE)
        check(core().rt_generate_camera_rays(self.handle, C.byref(camera), C.byref(params), psi, rays.ctypes.data),
              "rt_generate_camera_rays")
        return rays

    def close(self):
        if self.handle:
            core().rt_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """RtComm: the tile-assembly communicator of one rank (rt_comm_create)."""

    ID_BYTES = 128

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * Comm.ID_BYTES)()
        check(core().rt_comm_unique_id(buf), "rt_comm_unique_id")
        return bytes(buf)

    def __init__(self, uid, rank, world, device):
        self.handle = C.c_void_p()
        buf = (C.c_uint8 * Comm.ID_BYTES).from_buffer_copy(uid)
        check(core().rt_comm_create(buf, rank, world, device, C.byref(self.handle)), "rt_comm_create")
        self.rank, self.world, self.device = rank, world, device

    def render_multi(self, scene, camera, params, d_rgb_ptr, root=0, stream=None):
        """rt_render_multi: (this rank's RtRenderStats, assemble ms)"""
        stats = RtRenderStats()
        ms = C.c_float(0.0)
        check(core().rt_render_multi(scene.handle, C.byref(camera), C.byref(params), self.handle, root, d_rgb_ptr,
                                     C.byref(stats), C.byref(ms), stream), "rt_render_multi")
        return stats, ms.value

    def close(self):
        if self.handle:
            core().rt_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def packed_floats(width, height, world, rank, tile_size=0):
    return int(core().rt_packed_floats(width, height, tile_size, world, rank))


def tonemap_bgra8(rgb, exposure_stops=0.0, gamma=2.2, device=0):
    rgb = np.ascontiguousarray(rgb, np.float32)
    n = rgb.size // 3
    out = np.empty((n, 4), np.uint8)
    check(core().rt_tonemap_bgra8(device, rgb.ctypes.data, n, exposure_stops, gamma, out.ctypes.data), "rt_tonemap_bgra8")
    return out.reshape(rgb.shape[:-1] + (4,))


def stage1_render(width=512, height=512, device=0):
    """The Stage 1 program on the GPU: P6 payload as an (H, W, 3) uint8 array."""
    out = np.zeros((height, width, 3), np.uint8)
    if host().rth_stage1_render(device, width, height, out.ctypes.data) != 0:
        raise RtError("rth_stage1_render: " + host().rth_last_error_string().decode())
    return out


def stage23_render(stage, width=512, height=512, samples_u=4, samples_v=4, device=0):
    """The Stage 2 / Stage 3 program on the GPU.  Returns (float rgb before clamp,
    uint8 P6 payload, RtRenderStats).  Stage 2: samples_u random samples per pixel."""
    rgb = np.zeros((height, width, 3), np.float32)
    rgb8 = np.zeros((height, width, 3), np.uint8)
    stats = RtRenderStats()
    if host().rth_stage23_render(device, stage, width, height, samples_u, samples_v, rgb.ctypes.data,
                                 rgb8.ctypes.data, C.addressof(stats)) != 0:
        raise RtError("rth_stage23_render: " + host().rth_last_error_string().decode())
    return rgb, rgb8, stats


def libm_eval(kind, x, y=None):
    """Host build of the device's sinf (0) / cosf (1) / powf (2)."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    yp = None
    if y is not None:
        y = np.ascontiguousarray(y, np.float32)
        yp = y.ctypes.data
    check(core().rt_libm_eval(kind, x.ctypes.data, yp, x.size, out.ctypes.data), "rt_libm_eval")
    return out


def tile_owners(width, height, world, tile_size=0):
    """(owners[tiles_y, tiles_x], tile size) of the screen-tile partition."""
    tx, ty, ts = C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(core().rt_tile_owners(width, height, tile_size, world, None, C.byref(tx), C.byref(ty), C.byref(ts)))
    owners = np.zeros((ty.value, tx.value), np.uint32)
    check(core().rt_tile_owners(width, height, tile_size, world, owners.ctypes.data, None, None, None))
    return owners, ts.value


def sample_permutations(width, height, depth, x, y):
    out = np.zeros(5 * depth + 3, np.uint32)
    check(core().rt_sample_permutations(width, height, depth, x, y, out.ctypes.data))
    return out


def make_rays(origins, directions, tmax=1.0e30, time=0.0):
    n = len(origins)
    rays = np.zeros(n, RAY_DTYPE)
    rays["origin"] = origins
    rays["direction"] = directions
    rays["tmax"] = tmax
    rays["time"] = time
    return rays
