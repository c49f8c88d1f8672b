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
    assert lib.rt_abi_version() == 2


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
    d.abi_version = 2; d.set_xform = 0; d.num_finite = 1; d.num_infinite = 0
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


def _single_mesh_desc(capi, nodes, n):
    """RtSceneDesc of one mesh shape with n triangle faces over the given node table;
    returns (desc, keepalive) -- the arrays must outlive the call."""
    verts = np.zeros((3, 3), np.float32); verts[1, 0] = 1; verts[2, 1] = 1
    face_start = (np.arange(n + 1) * 3).astype(np.uint32)
    vidx = np.tile(np.array([0, 1, 2], np.uint32), n)
    nidx = np.full(3 * n, 0xffffffff, np.uint32)
    has_n = np.zeros(n, np.uint32)
    cdf = np.zeros(n + 1, np.float32)
    mesh = capi.RtMesh(0, 3, 0, 0, 0, n, 0, len(nodes), 0, 1.0)
    shape = capi.RtShape(3, 0, 0, 0, -1)
    xf = capi.RtXform(0, 0)
    mat = (C.c_float * 8)()
    d = capi.RtSceneDesc()
    d.abi_version = 2; d.set_xform = 0; d.num_finite = 1; d.num_infinite = 0
    d.shapes = C.addressof(shape); d.num_top_nodes = 0
    d.num_xforms = 1; d.xforms = C.addressof(xf)
    d.num_meshes = 1; d.meshes = C.addressof(mesh)
    d.num_vertices = 3; d.vertices = verts.ctypes.data
    d.num_faces = n; d.face_start = face_start.ctypes.data; d.face_has_normals = has_n.ctypes.data
    d.num_indices = 3 * n; d.vertex_index = vidx.ctypes.data; d.normal_index = nidx.ctypes.data
    d.num_mesh_nodes = len(nodes); d.mesh_nodes = nodes.ctypes.data
    d.num_cdf = n + 1; d.face_area_cdf = cdf.ctypes.data
    d.num_materials = 1; d.materials = C.addressof(mat)
    return d, (verts, face_start, vidx, nidx, has_n, cdf, mesh, shape, xf, mat, nodes)


def _balanced_tree(n):
    """2n-1 nodes over n faces in the reference's numbering (children after the parent,
    left subtree's descendants before the right's, RAccel.h:366-371)."""
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    one = np.float32(1.0).view(np.uint32)
    nodes[:, 3:6] = one
    todo = [(0, 1, n, 0)]               # node, first free slot, faces, first face
    while todo:
        node, base, cnt, first = todo.pop()
        if cnt == 1:
            nodes[node, 6] = first; nodes[node, 7] = 4
            continue
        left = cnt // 2
        nodes[node, 6] = base; nodes[node, 7] = 0
        todo.append((base + 1, base + 2 * left, cnt - left, first + left))
        todo.append((base, base + 2, left, first))
    return nodes


def test_large_bvh_validation_on_worker_threads(capi):
    """Trees of 65 536 nodes and more are validated data-parallel (rt_scene.cuh bvh_depth);
    the verdicts must be those of the serial walk.  Argument checks run before the device is
    looked for, so a CPU-only box sees them; a well-formed tree gets as far as 'no CUDA device'."""
    n = 40000
    good = _balanced_tree(n)
    h = C.c_void_p()
    lib = capi.core()
    ok_codes = (0,) if lib.rt_device_count() > 0 else (-2,)

    def create(nodes):
        d, keep = _single_mesh_desc(capi, np.ascontiguousarray(nodes), n)
        rc = lib.rt_scene_create(C.byref(d), 0, C.byref(h))
        if rc == 0:
            lib.rt_scene_destroy(h)
        return rc, lib.rt_last_error_string()

    rc, msg = create(good)
    assert rc in ok_codes, msg
    interior = np.flatnonzero(good[:, 7] == 0)
    leaves = np.flatnonzero(good[:, 7] == 4)

    bad = good.copy(); bad[leaves[1234], 6] = n             # primitive that does not exist
    rc, msg = create(bad); assert rc == -1 and b"primitive" in msg, msg
    bad = good.copy(); bad[interior[77], 7] = 3             # split axis 3
    rc, msg = create(bad); assert rc == -1 and b"axis" in msg, msg
    bad = good.copy(); bad[interior[500], 6] = 2 * n - 2    # second child out of range
    rc, msg = create(bad); assert rc == -1 and b"out of range" in msg, msg
    bad = good.copy(); bad[interior[900], 6] = good[interior[901], 6]   # two parents for one pair
    rc, msg = create(bad); assert rc == -1 and b"shared" in msg, msg
    bad = good.copy(); bad[interior[-1], 6] = 0             # points back at the root: not forward-numbered,
    rc, msg = create(bad); assert rc == -1 and b"cycle" in msg, msg   # the serial walk finds the cycle

    # the same tree numbered backwards (children BEFORE the parent) is legal for the ABI: serial walk
    perm = np.arange(2 * n - 1)[::-1].copy()               # new index of old node i is perm[i]
    back = np.zeros_like(good)
    back[perm] = good
    is_int = back[:, 7] == 0
    back[is_int, 6] = perm[back[is_int, 6]] - 1             # children (c, c+1) become (perm[c]-1, perm[c]): first child is perm[c+1]
    rc, msg = create(back)
    # mirrored numbering swaps left and right but is a well-formed tree of the same depth
    assert rc in ok_codes, msg

    # a 60-deep chain hanging under a big well-formed tree is refused like a small one
    deep = good.copy()
    extra = 60
    chain = np.zeros((2 * extra, 8), np.uint32)
    chain[:, 3:6] = np.float32(1.0).view(np.uint32)
    base = 2 * n - 1
    victim = leaves[0]
    prim = deep[victim, 6]
    deep[victim, 6] = base; deep[victim, 7] = 0
    for k in range(extra):
        a, b = 2 * k, 2 * k + 1
        chain[a, 6] = n + k; chain[a, 7] = 4                # leaf (new face)
        if k + 1 < extra:
            chain[b, 6] = base + 2 * (k + 1); chain[b, 7] = 0
        else:
            chain[b, 6] = prim; chain[b, 7] = 4
    big = np.concatenate([deep, chain])
    d, keep = _single_mesh_desc(capi, np.ascontiguousarray(big), n + extra)
    rc = lib.rt_scene_create(C.byref(d), 0, C.byref(h))
    assert rc == -3, lib.rt_last_error_string()


def test_scene_create_ex_argument_checks(capi, obj_path):
    """rt_scene_create_ex (RT_SCENE_BUILD_MESH_BVH): unknown flags are refused; a mesh that comes without nodes needs the
    flag (it must not silently vanish from the scene); Stage 6 rules cannot be built on the device; with the flag a
    node-less Stage 7 scene passes every host-side check and only stops at 'no CUDA device' on a CPU box."""
    lib = capi.core()
    h = C.c_void_p()
    no_gpu = lib.rt_device_count() == 0
    host_built = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, (24, 16))
    assert lib.rt_scene_create_ex(C.cast(host_built.desc, C.c_void_p), 0, 0x80, C.byref(h)) == -1
    assert b"flags" in lib.rt_last_error_string()
    left_to_gpu = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, (24, 16), tree=capi.TREE_DEVICE)
    assert left_to_gpu.desc.contents.num_mesh_nodes == 0 and left_to_gpu.desc.contents.num_faces == 24 * 16
    assert lib.rt_scene_create(C.cast(left_to_gpu.desc, C.c_void_p), 0, C.byref(h)) == -1
    assert b"no BVH nodes" in lib.rt_last_error_string()
    rc = lib.rt_scene_create_ex(C.cast(left_to_gpu.desc, C.c_void_p), 0, capi.RT_SCENE_BUILD_MESH_BVH, C.byref(h))
    if no_gpu:
        assert rc == -2 and b"no CUDA device" in lib.rt_last_error_string()
    else:
        assert rc == 0
        lib.rt_scene_destroy(h)
    # Stage 6 rules: prepare() builds on the host whatever the tree mode, so the flag has nothing to do ...
    s6 = capi.HostScene(capi.RECIPE_STAGE6_SCENE, obj_path, tree=capi.TREE_DEVICE)
    assert s6.desc.contents.num_mesh_nodes > 0
    # ... and a Stage 6 description WITHOUT nodes is refused
    bad = capi.RtSceneDesc.from_buffer_copy(bytes(left_to_gpu.desc.contents))
    bad.semantics = 6
    rc = lib.rt_scene_create_ex(C.byref(bad), 0, capi.RT_SCENE_BUILD_MESH_BVH, C.byref(h))
    assert rc in (-1, -4)       # (keyed transforms are refused under Stage 6 rules before the build question comes up)


def test_tree_mode_switch(capi):
    lib = capi.host()
    assert lib.rth_set_tree_mode(7) != 0 and b"unknown mode" in lib.rth_last_error_string()
    for mode in (capi.TREE_REFERENCE, capi.TREE_SAH, capi.TREE_DEVICE, capi.TREE_AUTO):
        assert lib.rth_set_tree_mode(mode) == 0
    # the default: small meshes keep their host-built tree, meshes of 65 536 faces or more leave it to the GPU
    small = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, (64, 48), tree=capi.TREE_AUTO)
    assert small.desc.contents.num_mesh_nodes == 2 * 64 * 48 - 1
    big = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, (320, 256), tree=capi.TREE_AUTO)
    assert big.desc.contents.num_mesh_nodes == 0 and big.desc.contents.num_faces == 320 * 256
