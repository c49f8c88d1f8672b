"""CPU tests: the C++ host mirror of the Rayito API must hand the GPU exactly the
state the reference holds after scene.prepare() -- same OBJ data, same transform
keys (including the Quaternion::operator*= quirk), same BVH node for node.  The
checker is the compiled reference (oracle/_ref)."""
import ctypes as C

import numpy as np
import pytest


def _array(ptr, ctype, count):
    if count == 0:
        return np.zeros(0, np.dtype(ctype))
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).copy()


def _nodes(ptr, count):
    return _array(ptr, C.c_uint32, count * 8).reshape(count, 8)


def _host_mesh_views(capi, desc):
    meshes = np.ctypeslib.as_array(C.cast(desc.meshes, C.POINTER(capi.RtMesh)), shape=(desc.num_meshes,)) \
        if desc.num_meshes else []
    return [capi.RtMesh.from_buffer_copy(bytes(m)) for m in meshes]


def _shapes(capi, desc):
    n = desc.num_finite + desc.num_infinite
    arr = (capi.RtShape * n).from_address(desc.shapes)
    return [capi.RtShape.from_buffer_copy(bytes(s)) for s in arr]


@pytest.mark.parametrize("which", ["scene1", "scene2", "scene6", "scene8", "deep", "deepboth"])
def test_top_level_bvh_identical(which, request):
    host = request.getfixturevalue(which + "_host")
    ref = request.getfixturevalue(which + "_ref")
    d = host.desc.contents
    mine = _nodes(d.top_nodes, d.num_top_nodes)
    theirs = ref.bvh_nodes(-1)
    assert mine.shape == theirs.shape and mine.shape[0] > 0
    assert np.array_equal(mine, theirs)
    assert d.num_finite == ref.num_finite and d.num_infinite == ref.num_infinite


@pytest.mark.parametrize("which", ["scene1", "scene2", "scene6", "scene7", "deep", "deepboth"])
def test_mesh_data_and_bvh_identical(which, request, capi):
    host = request.getfixturevalue(which + "_host")
    ref = request.getfixturevalue(which + "_ref")
    d = host.desc.contents
    shapes = _shapes(capi, d)
    meshes = _host_mesh_views(capi, d)
    verts = _array(d.vertices, C.c_float, d.num_vertices * 3).reshape(-1, 3)
    norms = _array(d.normals, C.c_float, d.num_normals * 3).reshape(-1, 3)
    face_start = _array(d.face_start, C.c_uint32, d.num_faces + 1)
    vidx = _array(d.vertex_index, C.c_uint32, d.num_indices)
    nidx = _array(d.normal_index, C.c_uint32, d.num_indices)
    cdf = _array(d.face_area_cdf, C.c_float, d.num_cdf)
    all_nodes = _nodes(d.mesh_nodes, d.num_mesh_nodes)
    seen = 0
    for si, sh in enumerate(shapes[:d.num_finite]):
        if sh.type != 3:
            continue
        seen += 1
        m = meshes[sh.geom]
        r = ref.mesh(si)
        assert r is not None
        v = verts[m.first_vertex:m.first_vertex + m.num_vertices]
        assert np.array_equal(v.view(np.uint32), r["vertices"].view(np.uint32))
        n = norms[m.first_normal:m.first_normal + m.num_normals]
        assert np.array_equal(n.view(np.uint32), r["normals"].view(np.uint32))
        fs = face_start[m.first_face:m.first_face + m.num_faces + 1]
        assert np.array_equal(np.diff(fs), r["face_sizes"])
        assert np.array_equal(vidx[fs[0]:fs[-1]], r["vertex_index"])
        assert np.array_equal(nidx[fs[0]:fs[-1]], r["normal_index"])
        c = cdf[m.first_cdf:m.first_cdf + m.num_faces + 1]
        assert np.array_equal(c.view(np.uint32), r["area_cdf"].view(np.uint32))
        assert np.float32(m.total_area).view(np.uint32) == r["area_cdf"][-1].view(np.uint32)
        mine = all_nodes[m.first_node:m.first_node + m.num_nodes]
        theirs = ref.bvh_nodes(si)
        assert mine.shape == theirs.shape
        assert np.array_equal(mine, theirs), "mesh BVH differs for shape %d" % si
    assert seen >= (1 if which in ("scene7", "deep", "deepboth") else 2)


@pytest.mark.parametrize("which", ["scene1", "scene2", "scene7", "scene8"])
def test_transform_keys_identical(which, request, capi):
    host = request.getfixturevalue(which + "_host")
    ref = request.getfixturevalue(which + "_ref")
    d = host.desc.contents
    shapes = _shapes(capi, d)
    xforms = (capi.RtXform * d.num_xforms).from_address(d.xforms)
    kt = _array(d.key_time, C.c_float, d.num_keys)
    ks = _array(d.key_scale, C.c_float, d.num_keys * 3).reshape(-1, 3)
    kr = _array(d.key_rotation, C.c_float, d.num_keys * 4).reshape(-1, 4)
    ktr = _array(d.key_translation, C.c_float, d.num_keys * 3).reshape(-1, 3)
    multi = 0
    for si, sh in enumerate(shapes):
        x = xforms[sh.xform]
        theirs = ref.shape_keys(si)
        assert theirs.shape[0] == x.num_keys
        sl = slice(x.first_key, x.first_key + x.num_keys)
        mine = np.concatenate([kt[sl, None], ks[sl], kr[sl], ktr[sl]], axis=1)
        assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32)), "keys differ for shape %d" % si
        multi += x.num_keys > 1
    assert multi >= (1 if which in ("scene7", "scene8") else 3)


def test_bumpy_tree_shape(scene1_host):
    d = scene1_host.desc.contents
    assert d.num_top_nodes == 15           # 8 finite shapes -> 2*8-1 nodes
    assert d.num_faces == 24576 + 6        # bumpy quads + the box
    assert scene1_host.depth(1) == 20      # bumpy.obj tree depth 20
