"""Device-side face-BVH build (rt_scene_create_ex with RT_SCENE_BUILD_MESH_BVH, rayito_b200/csrc/rt_build.cuh).

The GPU has to reproduce Bvh<T>::build / buildRange (Rayito_Stage7_QT/RAccel.h:262-374) node for node:
std::partition's element order, the recursion's slot numbering, boxes down to the sign of a zero.  The host
builder is pinned to the compiled reference node for node by the CPU tests (tests/test_host_parity.py,
tests/test_host_parallel.py), so here the tree read back from the device is compared BYTE FOR BYTE with the
host builder's on every mesh of the GUI scenes (bumpy.obj, the 6-quad cube: the one-thread phase alone),
the synthetic sphere at three sizes up to config C5's own 10 M triangles, and the deliberately skewed wedge
(46 deep: "cut in half" fallbacks all the way down).  Then hits and a whole raytrace() through the host API in
device-build mode against the reference."""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import BIG_GRID, DEEPER_GRID, SYNTH_GRID
from tests.raybatches import bits, random_rays

pytestmark = pytest.mark.gpu


def _host_nodes(capi, scene):
    d = scene.desc.contents
    meshes = (capi.RtMesh * d.num_meshes).from_address(C.cast(d.meshes, C.c_void_p).value) if d.num_meshes else []
    out = []
    for m in meshes:
        addr = C.cast(d.mesh_nodes, C.c_void_p).value + 32 * m.first_node
        nodes = np.frombuffer(C.string_at(addr, 32 * m.num_nodes), np.uint32).reshape(-1, 8) if m.num_nodes else np.zeros((0, 8), np.uint32)
        out.append((m.num_faces, nodes))
    return out


def _compare(capi, recipe, obj, grid, label):
    ref_scene = capi.HostScene(recipe, obj, grid)
    dev_scene = capi.HostScene(recipe, obj, grid, tree=capi.TREE_DEVICE)
    dd = dev_scene.desc.contents
    assert dd.num_mesh_nodes == 0          # prepare() left the trees to the GPU
    want = _host_nodes(capi, ref_scene)
    dev = capi.DeviceScene(dev_scene.desc, build_bvh_on_device=True)
    deepest, total_ms = 0, 0.0
    for m, (faces, nodes) in enumerate(want):
        got, depth, ms = dev.mesh_nodes(m, faces)
        assert got.shape == nodes.shape, label
        if not np.array_equal(got, nodes):
            bad = np.nonzero((got != nodes).any(axis=1))[0]
            raise AssertionError("%s mesh %d: %d of %d nodes differ, first %d: device %s host %s" % (
                label, m, len(bad), len(nodes), bad[0], got[bad[0]], nodes[bad[0]]))
        deepest, total_ms = depth, ms
    assert deepest == max(ref_scene.depth(m) for m in range(len(want))), label
    print("%s: %d meshes, %d faces, device build %.2f ms, host prepare %.1f ms" % (
        label, len(want), sum(f for f, _ in want), total_ms, 1e3 * ref_scene.prepare_seconds))
    return ref_scene, dev_scene, dev


def test_device_build_gui_scenes(capi, obj_path, scene1_ref):
    _ref, _devs, dev = _compare(capi, capi.RECIPE_STAGE7_SCENE1, obj_path, (0, 0), "scene 1")
    rays = random_rays(1 << 17, seed=301, center=(0.1, 0, 0), radius=6.0, target_radius=1.6, shadow_fraction=0.25)
    hits = dev.trace_closest(rays, extended=True)
    want = scene1_ref.trace_closest(rays)
    for f in ("shape", "face", "tri"):
        assert np.array_equal(hits[f], want[f]), f
    assert np.array_equal(bits(hits["t"]), bits(want["t"]))
    assert np.array_equal(dev.trace_any(rays), scene1_ref.trace_any(rays))
    dev.close()
    _compare(capi, capi.RECIPE_STAGE7_SCENE2, None, (0, 0), "scene 2")[2].close()


@pytest.mark.parametrize("grid", [(7, 5), (33, 1), SYNTH_GRID, (640, 512), BIG_GRID])
def test_device_build_synthetic_sphere(capi, grid):
    _compare(capi, capi.RECIPE_SYNTHETIC_MESH, None, grid, "sphere %dx%d" % grid)[2].close()


def test_device_build_skewed_wedge(capi):
    _compare(capi, capi.RECIPE_EDGE_DEEP_BOTH, None, DEEPER_GRID, "wedge")[2].close()


def test_device_build_c5_mesh(capi):
    """Config C5's own mesh: 4 999 696 quads, 9 999 391 nodes, 37 deep."""
    ref_scene, _d, dev = _compare(capi, capi.RECIPE_SYNTHETIC_MESH, None, (2236, 2236), "C5 mesh")
    assert ref_scene.depth(0) == 37
    dev.close()


@pytest.mark.parametrize("on_device", [0, 1])
def test_device_build_mixed_with_host_trees(capi, scene1_host, scene1_ref, on_device):
    """A scene in which one mesh brings its nodes and the other leaves them to the GPU (what the default tree mode
    produces when only some meshes are large): both trees are where they belong and hits are the reference's."""
    d = scene1_host.desc.contents
    assert d.num_meshes == 2
    meshes = (capi.RtMesh * 2).from_address(C.cast(d.meshes, C.c_void_p).value)
    mixed = (capi.RtMesh * 2)()
    for m in range(2):
        C.memmove(C.byref(mixed[m]), C.byref(meshes[m]), C.sizeof(capi.RtMesh))
    mixed[on_device].num_nodes = 0
    desc = capi.RtSceneDesc()
    C.memmove(C.byref(desc), C.byref(d), C.sizeof(capi.RtSceneDesc))
    desc.meshes = C.addressof(mixed)
    dev = capi.DeviceScene(C.pointer(desc), build_bvh_on_device=True)
    want = _host_nodes(capi, scene1_host)
    for m, (faces, nodes) in enumerate(want):
        got, _depth, _ms = dev.mesh_nodes(m, faces)
        assert np.array_equal(got, nodes), "mesh %d (built on the %s)" % (m, "device" if m == on_device else "host")
    rays = random_rays(1 << 16, seed=302, center=(0.1, 0, 0), radius=6.0, target_radius=2.5, shadow_fraction=0.25)
    hits = dev.trace_closest(rays, extended=True)
    ref_hits = scene1_ref.trace_closest(rays)
    for f in ("shape", "face", "tri"):
        assert np.array_equal(hits[f], ref_hits[f]), f
    assert np.array_equal(bits(hits["t"]), bits(ref_hits["t"]))
    dev.close()


def test_default_tree_mode_builds_large_meshes_on_the_device(capi):
    """rayito_b200::treeMode() defaults to kTreeAuto: a mesh of 65 536 faces or more comes out of prepare() without
    nodes and gets the reference's tree from the GPU; smaller ones are built on the host as before."""
    small = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, SYNTH_GRID, tree=capi.TREE_AUTO)
    assert small.desc.contents.num_mesh_nodes == 2 * SYNTH_GRID[0] * SYNTH_GRID[1] - 1
    grid = (640, 512)
    big = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, grid, tree=capi.TREE_AUTO)
    assert big.desc.contents.num_mesh_nodes == 0
    ref_scene = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, grid)
    dev = capi.DeviceScene(big.desc, build_bvh_on_device=True)
    got, depth, _ms = dev.mesh_nodes(0, grid[0] * grid[1])
    assert np.array_equal(got, _host_nodes(capi, ref_scene)[0][1]) and depth == ref_scene.depth(0)
    dev.close()


def test_build_flag_is_harmless_for_a_stage6_scene(capi, scene6_host, scene6_ref):
    """Stage 6 rules root the face BVH in the all-vertex box, which only the host builds: prepare() keeps building there
    whatever the tree mode, and RT_SCENE_BUILD_MESH_BVH on a scene whose meshes all bring their nodes builds nothing."""
    assert scene6_host.desc.contents.num_mesh_nodes > 0
    dev = capi.DeviceScene(scene6_host.desc, build_bvh_on_device=True)
    rays = random_rays(1 << 15, seed=303, center=(0, 0, 0), radius=8.0, target_radius=2.0)
    hits = dev.trace_closest(rays, extended=True)
    want = scene6_ref.trace_closest(rays)
    for f in ("shape", "face", "tri"):
        assert np.array_equal(hits[f], want[f]), f
    assert np.array_equal(bits(hits["t"]), bits(want["t"]))
    dev.close()


def test_raytrace_with_device_build_matches_reference(capi, obj_path, scene1_host, scene1_ref):
    """Rayito::raytrace() with rayito_b200::treeMode() = kTreeDevice: same image as the reference, bit for bit."""
    lib = capi.host()
    spec = scene1_host.default_camera_spec()
    W, H, ps = 96, 54, 2
    theirs, rstats = scene1_ref.render(spec, W, H, ps, ls=1, depth=3)
    app = lib.rth_app_create(capi.RECIPE_STAGE7_SCENE1, obj_path.encode(), 0, 0)
    assert app
    img = np.zeros((H, W, 3), np.float32)
    stats = capi.RtRenderStats()
    assert lib.rth_set_tree_mode(capi.TREE_DEVICE) == 0
    try:
        rc = lib.rth_app_raytrace(app, spec.ctypes.data, W, H, ps, 1, 3, 0, 0, 1, 0, img.ctypes.data, 0, C.byref(stats))
    finally:
        lib.rth_set_tree_mode(capi.TREE_AUTO)
        lib.rth_app_destroy(app)
    assert rc == 0, lib.rth_last_error_string()
    assert np.array_equal(bits(img), bits(theirs))
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls
