"""CPU tests: the host worker threads (rayito_b200/host/rayito_b200/parallel.hpp) must not
change a single bit of what prepare() + flatten hand to the GPU.  A mesh large enough to
take the threaded path (>= 65 536 faces: subtree jobs, chunked bounds / areas / face tables)
(and, at 327 680 faces with RAYITO_B200_WIDE_SPLITS=1, the all-worker splits at the top of the tree
with their exact parallel std::partition) is prepared with 1, 3 and 8 workers and compared array by array; the single-thread result is
compared node for node with the compiled reference (oracle/_ref), Bvh<T>::build
(Rayito_Stage7_QT/RAccel.h:262-374) and Mesh::prepare (RMesh.h:89-129)."""
import ctypes as C
import os

import numpy as np
import pytest

GRID = (640, 512)       # 327 680 quads: above the 262 144-element threshold of the opt-in all-worker top splits (RAYITO_B200_WIDE_SPLITS=1)


def _bytes(ptr, nbytes):
    if nbytes == 0 or not ptr:
        return b""
    return C.string_at(ptr, nbytes)


def _snapshot(capi, threads, wide=False):
    saved = {k: os.environ.get(k) for k in ("RAYITO_B200_HOST_THREADS", "RAYITO_B200_WIDE_SPLITS")}
    os.environ["RAYITO_B200_HOST_THREADS"] = str(threads)
    os.environ["RAYITO_B200_WIDE_SPLITS"] = "1" if wide else "0"
    try:
        scene = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, GRID)
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


@pytest.mark.parametrize("threads,wide", [(3, False), (8, False), (3, True), (8, True)])
def test_threaded_prepare_is_bit_identical(capi, serial, threads, wide):
    _scene, want = serial
    _scene2, got = _snapshot(capi, threads, wide)
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


def test_parallel_partition_equals_std_partition(tmp_path):
    """rayito_b200::parallelPartition reproduces libstdc++'s std::partition element for element
    (300 size / split / thread-count cases, including all-true, all-false and tiny ranges)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "parallel_partition_test")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++11", "-pthread", "-I" + os.path.join(root, "rayito_b200", "host"),
                    os.path.join(root, "tests", "cpp", "parallel_partition_test.cpp"), "-o", exe], check=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    assert "300 cases identical" in out, out
