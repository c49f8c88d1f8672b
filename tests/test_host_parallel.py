"""CPU tests: the host worker threads (rayito_b200/host/rayito_b200/parallel.hpp) must not
change a single bit of what prepare() + flatten hand to the GPU.  A mesh large enough to
take the threaded path (>= 65 536 faces: subtree jobs, chunked bounds / areas / face tables)
is prepared with 1, 3 and 8 workers and compared array by array; the single-thread result is
compared node for node with the compiled reference (oracle/_ref), Bvh<T>::build
(Rayito_Stage7_QT/RAccel.h:262-374) and Mesh::prepare (RMesh.h:89-129)."""
import ctypes as C
import os

import numpy as np
import pytest

GRID = (640, 512)       # 327 680 quads


def _bytes(ptr, nbytes):
    if nbytes == 0 or not ptr:
        return b""
    return C.string_at(ptr, nbytes)


def _snapshot(capi, threads, tree=0):
    saved = {k: os.environ.get(k) for k in ("RAYITO_B200_HOST_THREADS",)}
    os.environ["RAYITO_B200_HOST_THREADS"] = str(threads)
    try:
        scene = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, GRID, tree=tree)
    finally:
        for k, v in saved.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    d = scene.desc.contents
    snap = {
        "top_nodes": _bytes(d.top_nodes, d.num_top_nodes * 32),
        "mesh_nodes": _bytes(d.mesh_nodes, d.num_mesh_nodes * 32),
        "vertices": _bytes(d.vertices, d.num_vertices * 12),
        "normals": _bytes(d.normals, d.num_normals * 12),
        "face_start": _bytes(d.face_start, (d.num_faces + 1) * 4),
        "face_has_normals": _bytes(d.face_has_normals, d.num_faces * 4),
        "vertex_index": _bytes(d.vertex_index, d.num_indices * 4),
        "normal_index": _bytes(d.normal_index, d.num_indices * 4),
        "cdf": _bytes(d.face_area_cdf, d.num_cdf * 4),
        "key_rotation": _bytes(d.key_rotation, d.num_keys * 16),
        "depth": scene.depth(0),
        "counts": (d.num_faces, d.num_mesh_nodes, d.num_indices, d.num_cdf),
    }
    return scene, snap


@pytest.fixture(scope="module")
def serial(capi):
    return _snapshot(capi, 1)


@pytest.mark.parametrize("threads", [3, 8])
def test_threaded_prepare_is_bit_identical(capi, serial, threads):
    _scene, want = serial
    _scene2, got = _snapshot(capi, threads)
    assert got["counts"] == want["counts"] and want["counts"][0] == GRID[0] * GRID[1]
    for key in want:
        assert got[key] == want[key], "%s differs with %d host threads" % (key, threads)


def test_large_mesh_matches_reference_node_for_node(capi, serial, ref):
    scene, _snap = serial
    r = ref.RefScene(5, None, GRID)
    d = scene.desc.contents
    shapes = (capi.RtShape * (d.num_finite + d.num_infinite)).from_address(d.shapes)
    mesh_shape = [i for i in range(d.num_finite) if shapes[i].type == 3]
    assert len(mesh_shape) == 1
    theirs = r.bvh_nodes(mesh_shape[0])
    mine = np.frombuffer(_snap["mesh_nodes"], np.uint32).reshape(-1, 8)
    assert mine.shape == theirs.shape == (2 * GRID[0] * GRID[1] - 1, 8)
    assert np.array_equal(mine, theirs)
    rm = r.mesh(mesh_shape[0])
    assert np.array_equal(np.frombuffer(_snap["cdf"], np.uint32), rm["area_cdf"].view(np.uint32))
    top = np.frombuffer(_snap["top_nodes"], np.uint32).reshape(-1, 8)
    assert np.array_equal(top, r.bvh_nodes(-1))      # the mesh's world box over its key times


def test_app_handle_builds_without_preparing(capi):
    """rth_app_create is the application's scene-building code only (no GPU needed);
    rth_app_raytrace needs the device and must fail loudly without one."""
    lib = capi.host()
    app = lib.rth_app_create(capi.RECIPE_STAGE7_SCENE2, None, 0, 0)
    assert app
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        spec = np.zeros(14, np.float32)
        spec[:] = [30, -4, 5, 15, 0, 0, 0, 0, 1, 0, 16, 0, 0, 1]
        img = np.zeros((8, 8, 3), np.float32)
        rc = lib.rth_app_raytrace(app, spec.ctypes.data, 8, 8, 1, 1, 1, 0, 0, 1, 0, img.ctypes.data, 0, None)
        assert rc != 0
        assert b"CUDA" in lib.rth_last_error_string() or b"device" in lib.rth_last_error_string()
    lib.rth_app_destroy(app)
    assert lib.rth_app_create(12345, None, 0, 0) in (None, 0)


def _check_tree(nodes, num_faces):
    """Structural validity of a face BVH in the reference's node format: every face in exactly one
    leaf, children numbered b, b+1, every child's box inside its parent's, the first child on the
    HIGH side of the split axis (what the traversal's near / far choice relies on)."""
    boxes = nodes[:, :6].view(np.float32)
    word, flags = nodes[:, 6], nodes[:, 7]
    leaf = (flags & 4) != 0
    assert nodes.shape[0] == 2 * num_faces - 1
    assert np.array_equal(np.sort(word[leaf]), np.arange(num_faces, dtype=np.uint32))
    inner = np.nonzero(~leaf)[0]
    first = word[inner].astype(np.int64)
    kids = np.sort(np.concatenate([first, first + 1]))
    assert np.array_equal(kids, np.arange(1, nodes.shape[0]))          # every node but the root is somebody's child, once
    for c in (first, first + 1):
        assert (boxes[c, :3] >= boxes[inner, :3]).all() and (boxes[c, 3:] <= boxes[inner, 3:]).all()
    axis = (flags[inner] & 3).astype(np.int64)
    assert (axis <= 2).all()
    centre = 0.5 * (boxes[:, :3] + boxes[:, 3:])
    rows = np.arange(len(inner))
    # box centres of the two sides: the first child is the high side (ties allowed where the range was cut in half)
    assert (centre[first, axis][rows] >= centre[first + 1, axis][rows] - 1e-6).mean() > 0.99
    depth = np.zeros(nodes.shape[0], np.int32)
    order = inner[np.argsort(inner)]
    for i in order:                 # children are numbered above their parent
        depth[word[i]] = depth[word[i] + 1] = depth[i] + 1
    return int(depth.max())


def test_perf_mode_tree_is_a_valid_bvh_and_threads_do_not_change_it(capi, serial):
    """rth_set_tree_mode(1): the binned-SAH face BVH (PERF MODE, rayito_b200/host/rayito_b200/accel.hpp) has the
    reference's node format and conventions, differs from the reference's tree, and is the same whatever the
    number of host workers; everything else prepare() produces is untouched."""
    _ref_scene, want = serial
    scene1, one = _snapshot(capi, 1, tree=capi.TREE_SAH)
    _scene8, eight = _snapshot(capi, 8, tree=capi.TREE_SAH)
    assert one["mesh_nodes"] == eight["mesh_nodes"] and one["depth"] == eight["depth"]
    assert one["mesh_nodes"] != want["mesh_nodes"]
    for key in ("vertices", "normals", "face_start", "vertex_index", "normal_index", "cdf", "top_nodes", "counts"):
        assert one[key] == want[key], key
    nodes = np.frombuffer(one["mesh_nodes"], np.uint32).reshape(-1, 8)
    depth = _check_tree(nodes, GRID[0] * GRID[1])
    assert depth == one["depth"]
    ref_depth = _check_tree(np.frombuffer(want["mesh_nodes"], np.uint32).reshape(-1, 8), GRID[0] * GRID[1])
    assert ref_depth == want["depth"]
    assert depth <= ref_depth       # the midpoint rule degenerates at the poles of the displaced sphere
    # a later scene on the same thread gets the reference's tree again (the mode is per call of HostScene)
    _again, back = _snapshot(capi, 1)
    assert back["mesh_nodes"] == want["mesh_nodes"]
