"""CPU tests of the boundary: the libraries load, export every symbol the headers
declare, and fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, text)))


def test_core_exports_every_declared_symbol(capi):
    declared = _declared("rayito_b200.h", "rt_")
    assert sorted(capi.CORE_SYMBOLS) == declared
    lib = capi.core()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rt_abi_version() == 1


def test_host_exports_every_declared_symbol(capi):
    declared = _declared("rayito_b200_host.h", "rth_")
    assert sorted(capi.HOST_SYMBOLS) == declared
    lib = capi.host()
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback(capi, scene2_host):
    """Without a device every compute entry point must fail, not fall back."""
    if capi.core().rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.RtError, match="no CUDA device"):
        capi.DeviceScene(scene2_host.desc)
    with pytest.raises(capi.RtError, match="no CUDA device"):
        capi.tonemap_bgra8(np.zeros((4, 3), np.float32))
    spec = scene2_host.default_camera_spec()
    img = np.zeros((8, 8, 3), np.float32)
    stats = capi.RtRenderStats()
    rc = capi.host().rth_raytrace(2, None, 0, 0, spec.ctypes.data, 8, 8, 1, 1, 1, 0, 0, 1, 0, img.ctypes.data, C.byref(stats))
    assert rc != 0 and b"no CUDA device" in capi.host().rth_last_error_string()


def test_scene_validation_rejects_bad_descriptions(capi, scene2_host):
    """Argument checks run before any CUDA call, so they are testable on CPU."""
    import copy
    good = scene2_host.desc.contents
    bad = capi.RtSceneDesc.from_buffer_copy(bytes(good))
    bad.abi_version = 99
    h = C.c_void_p()
    rc = capi.core().rt_scene_create(C.byref(bad), 0, C.byref(h))
    assert rc == -1 and b"abi_version" in capi.core().rt_last_error_string()
    bad = capi.RtSceneDesc.from_buffer_copy(bytes(good))
    bad.num_top_nodes = 0
    rc = capi.core().rt_scene_create(C.byref(bad), 0, C.byref(h))
    assert rc == -1
    assert capi.core().rt_scene_create(None, 0, C.byref(h)) == -1


def test_deep_bvh_is_refused(capi):
    """A BVH deeper than 49 would overflow the reference's 50-entry traversal stack
    (RAccel.h:379,414,502); the boundary refuses it instead of truncating silently."""
    # a degenerate chain: 60 faces on a line, each split peels one face off
    n = 60
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    f32 = lambda v: np.float32(v).view(np.uint32)
    nxt = 1
    cur = 0
    for k in range(n - 1):
        nodes[cur, :3] = f32(0.0); nodes[cur, 3:6] = f32(1.0)
        nodes[cur, 6] = nxt; nodes[cur, 7] = 0
        nodes[nxt, 6] = k; nodes[nxt, 7] = 4            # left child: leaf k
        cur = nxt + 1
        nxt += 2
    nodes[cur, 6] = n - 1; nodes[cur, 7] = 4
    verts = np.zeros((3, 3), np.float32); verts[1, 0] = 1; verts[2, 1] = 1
    face_start = (np.arange(n + 1) * 3).astype(np.uint32)
    vidx = np.tile(np.array([0, 1, 2], np.uint32), n)
    nidx = np.full(3 * n, 0xffffffff, np.uint32)
    has_n = np.zeros(n, np.uint32)
    cdf = np.zeros(n + 1, np.float32)
    mesh = capi.RtMesh(0, 3, 0, 0, 0, n, 0, 2 * n - 1, 0, 1.0)
    shape = capi.RtShape(3, 0, 0, 0, -1)
    xf = capi.RtXform(0, 0)
    mat = (C.c_float * 8)()
    d = capi.RtSceneDesc()
    d.abi_version = 1; d.set_xform = 0; d.num_finite = 1; d.num_infinite = 0
    d.shapes = C.addressof(shape); d.num_top_nodes = 0
    d.num_xforms = 1; d.xforms = C.addressof(xf)
    d.num_meshes = 1; d.meshes = C.addressof(mesh)
    d.num_vertices = 3; d.vertices = verts.ctypes.data
    d.num_faces = n; d.face_start = face_start.ctypes.data; d.face_has_normals = has_n.ctypes.data
    d.num_indices = 3 * n; d.vertex_index = vidx.ctypes.data; d.normal_index = nidx.ctypes.data
    d.num_mesh_nodes = 2 * n - 1; d.mesh_nodes = nodes.ctypes.data
    d.num_cdf = n + 1; d.face_area_cdf = cdf.ctypes.data
    d.num_materials = 1; d.materials = C.addressof(mat)
    h = C.c_void_p()
    rc = capi.core().rt_scene_create(C.byref(d), 0, C.byref(h))
    assert rc == -3, capi.core().rt_last_error_string()
