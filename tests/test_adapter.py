"""integration/rayito_ref_adapter.h: the reference-side binding (INTEGRATION.md option B).

The adapter is compiled against the reference's own headers and sources (oracle/Makefile ->
oracle/_ref/libref_adapter.so).  CPU: a scene built with the REFERENCE's classes, prepared by
the reference's own prepare() and flattened by the adapter must be the very scene this repo's
host library flattens for the same recipe -- every array of the RtSceneDesc byte for byte.
GPU: the adapter's raytrace() (reference scene -> C ABI -> B200) returns the reference's own
image bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.raybatches import bits

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_adapter.so")


@pytest.fixture(scope="module")
def adapter(capi):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_adapter.so not built (needs /root/reference at build time)")
    capi.core()
    L = C.CDLL(LIB, mode=C.RTLD_LOCAL)
    L.adapter_last_error.restype = C.c_char_p
    L.adapter_scene_create.restype = C.c_void_p
    L.adapter_scene_create.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint]
    L.adapter_scene_destroy.argtypes = [C.c_void_p]
    L.adapter_scene_desc.restype = C.POINTER(capi.RtSceneDesc)
    L.adapter_scene_desc.argtypes = [C.c_void_p]
    L.adapter_camera.argtypes = [C.c_void_p, C.POINTER(capi.RtCamera)]
    L.adapter_raytrace.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_uint,
                                   C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.POINTER(capi.RtRenderStats)]
    return L


# (count field or expression, pointer field, bytes per element)
ARRAYS = [
    (lambda d: d.num_finite + d.num_infinite, "shapes", 20), ("num_top_nodes", "top_nodes", 32), ("num_xforms", "xforms", 8),
    ("num_keys", "key_time", 4), ("num_keys", "key_scale", 12), ("num_keys", "key_rotation", 16),
    ("num_keys", "key_translation", 12), ("num_planes", "planes", 28), ("num_spheres", "spheres", 16),
    ("num_rects", "rects", 36), ("num_meshes", "meshes", 40), ("num_vertices", "vertices", 12),
    ("num_normals", "normals", 12), (lambda d: d.num_faces + 1 if d.num_faces else 0, "face_start", 4),
    ("num_faces", "face_has_normals", 4), ("num_indices", "vertex_index", 4), ("num_mesh_nodes", "mesh_nodes", 32),
    ("num_cdf", "face_area_cdf", 4), ("num_materials", "materials", 32), ("num_lights", "lights", 4),
]
SCALARS = ["abi_version", "set_xform", "num_finite", "num_infinite", "num_top_nodes", "num_xforms", "num_keys", "num_planes",
           "num_spheres", "num_rects", "num_meshes", "num_vertices", "num_normals", "num_faces", "num_indices",
           "num_mesh_nodes", "num_cdf", "num_materials", "num_lights", "semantics"]


def _bytes(desc, count, field, width):
    n = count(desc) if callable(count) else getattr(desc, count)
    ptr = getattr(desc, field)
    if n == 0:
        return b""
    return C.string_at(ptr, n * width)


@pytest.mark.parametrize("recipe,grid", [(1, (0, 0)), (3, (0, 0)), (2, (0, 0)), (5, (64, 48)), (7, (0, 0)), (8, (0, 0)),
                                         (9, (0, 0)), (10, (40, 8)), (11, (44, 8))])
def test_adapter_flattens_the_reference_scene_like_the_host_library(adapter, capi, obj_path, recipe, grid):
    obj = obj_path.encode() if recipe in (1, 3) else None
    h = adapter.adapter_scene_create(recipe, obj, grid[0], grid[1])
    assert h, adapter.adapter_last_error()
    try:
        theirs = adapter.adapter_scene_desc(h).contents
        host = capi.HostScene(recipe, obj_path if recipe in (1, 3) else None, grid)
        mine = host.desc.contents
        for f in SCALARS:
            assert getattr(theirs, f) == getattr(mine, f), f
        for count, field, width in ARRAYS:
            assert _bytes(theirs, count, field, width) == _bytes(mine, count, field, width), field
        # normal indices only where a face has normals (the rest is padding in both)
        if mine.num_indices:
            fs = np.frombuffer(_bytes(mine, lambda d: d.num_faces + 1, "face_start", 4), np.uint32)
            has = np.frombuffer(_bytes(mine, "num_faces", "face_has_normals", 4), np.uint32)
            a = np.frombuffer(_bytes(theirs, "num_indices", "normal_index", 4), np.uint32)
            b = np.frombuffer(_bytes(mine, "num_indices", "normal_index", 4), np.uint32)
            keep = np.repeat(has != 0, np.diff(fs))
            assert np.array_equal(a[keep], b[keep])
        # the camera description
        spec = host.default_camera_spec()
        cam_a, cam_b = capi.RtCamera(), capi.camera_from_spec(spec)
        adapter.adapter_camera(spec.ctypes.data, C.byref(cam_a))
        assert bytes(cam_a) == bytes(cam_b)
    finally:
        adapter.adapter_scene_destroy(h)


@pytest.mark.gpu
@pytest.mark.parametrize("recipe,W,H,ps,ls,depth", [(1, 96, 54, 3, 1, 3), (2, 64, 36, 2, 2, 2), (7, 48, 27, 2, 1, 3)])
def test_adapter_raytrace_matches_reference(adapter, capi, ref, obj_path, recipe, W, H, ps, ls, depth):
    """The reference application's scene, built and prepared by the reference's own classes, rendered
    through rayito_b200_adapter::raytrace(): the reference's image, bit for bit."""
    obj = obj_path if recipe == 1 else None
    host = capi.HostScene(recipe, obj)
    spec = host.default_camera_spec()
    theirs, rstats = ref.RefScene(recipe, obj).render(spec, W, H, ps, ls=ls, depth=depth)
    rgb = np.zeros((H, W, 3), np.float32)
    stats = capi.RtRenderStats()
    rc = adapter.adapter_raytrace(recipe, obj.encode() if obj else None, 0, 0, spec.ctypes.data, W, H, ps, ls, depth, 0,
                                  rgb.ctypes.data, C.byref(stats))
    assert rc == 0, adapter.adapter_last_error()
    assert np.array_equal(bits(rgb), bits(theirs))
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls
