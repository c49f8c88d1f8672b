"""GPU parity tests proper: hit records from the CUDA traversal kernels, called
through the C ABI, must equal the compiled reference's on the same rays -- winner
(shape, face, fan triangle) identical and t, normal, colour modifier BIT-equal."""
import numpy as np
import pytest

from tests.raybatches import axis_parallel_rays, bits, random_rays

pytestmark = pytest.mark.gpu


def _compare_closest(dev, ref_scene, rays, label):
    mine = dev.trace_closest(rays, extended=True)
    theirs = ref_scene.trace_closest(rays)
    bad = np.flatnonzero((mine["shape"] != theirs["shape"]) | (mine["face"] != theirs["face"]) |
                         (mine["tri"] != theirs["tri"]) | (bits(mine["t"]) != bits(theirs["t"])))
    assert bad.size == 0, "%s: %d/%d hit records differ, first %d: mine %s ref %s" % (
        label, bad.size, len(rays), bad[0], mine[bad[0]], theirs[bad[0]])
    hit = theirs["shape"] >= 0
    assert np.array_equal(bits(mine["normal"][hit]), bits(theirs["normal"][hit])), label + ": normals differ"
    assert np.array_equal(bits(mine["color_modifier"][hit]), bits(theirs["color_modifier"][hit, 0])), label
    plain = dev.trace_closest(rays, extended=False)
    assert np.array_equal(plain["shape"], mine["shape"]) and np.array_equal(bits(plain["t"]), bits(mine["t"]))
    return mine


def _compare_any(dev, ref_scene, rays, label):
    mine = dev.trace_any(rays)
    theirs = ref_scene.trace_any(rays)
    bad = np.flatnonzero(mine != theirs)
    assert bad.size == 0, "%s: %d/%d any-hit results differ, first %d" % (label, bad.size, len(rays), bad[0])
    return mine


@pytest.fixture(scope="module")
def dev1(capi, scene1_host):
    d = capi.DeviceScene(scene1_host.desc)
    yield d
    d.close()


@pytest.fixture(scope="module")
def dev2(capi, scene2_host):
    d = capi.DeviceScene(scene2_host.desc)
    yield d
    d.close()


def test_scene1_random_rays(dev1, scene1_ref):
    rays = random_rays(1 << 18, seed=11, center=(0, -0.5, 0), radius=12.0, target_radius=4.0)
    hits = _compare_closest(dev1, scene1_ref, rays, "scene1 random")
    # the batch must actually exercise the scene: every shape type gets hit
    assert set(np.unique(hits["shape"])) >= set(range(-1, 9))
    assert (hits["face"] >= 0).sum() > 10000


def test_scene1_random_rays_any(dev1, scene1_ref):
    rays = random_rays(1 << 18, seed=12, center=(0, -0.5, 0), radius=12.0, target_radius=4.0, shadow_fraction=0.7)
    got = _compare_any(dev1, scene1_ref, rays, "scene1 any")
    assert 0.05 < got.mean() < 0.999


def test_scene1_mesh_focus(dev1, scene1_ref):
    # Aim at bumpy.obj (around the origin, radius ~1.3) so most rays traverse the face BVH
    rays = random_rays(1 << 18, seed=13, center=(0.1, 0.0, 0.0), radius=6.0, target_radius=1.4)
    hits = _compare_closest(dev1, scene1_ref, rays, "scene1 mesh focus")
    assert (hits["shape"] == 5).mean() > 0.3
    _compare_any(dev1, scene1_ref, rays, "scene1 mesh focus any")


def test_scene1_axis_parallel_and_negative_zero(dev1, scene1_ref):
    rays = axis_parallel_rays(1 << 16, seed=14)
    _compare_closest(dev1, scene1_ref, rays, "scene1 axis-parallel")
    _compare_any(dev1, scene1_ref, rays, "scene1 axis-parallel any")


def test_scene1_recorded_path_rays(dev1, scene1_ref, scene1_host, capi):
    """Every ray the reference casts while rendering a small frame (camera rays,
    BSDF-MIS probes, bounce rays; shadow rays with finite tMax and their times)."""
    spec = scene1_host.default_camera_spec()
    _img, stats = scene1_ref.render(spec, 128, 72, 2, ls=1, depth=3, record_rays=True)
    closest = scene1_ref.recorded_rays(0, capi.RAY_DTYPE)
    shadow = scene1_ref.recorded_rays(1, capi.RAY_DTYPE)
    assert len(closest) == stats.closest_calls and len(shadow) == stats.any_calls
    assert len(closest) > 100000 and len(shadow) > 30000
    _compare_closest(dev1, scene1_ref, closest, "scene1 recorded closest")
    _compare_any(dev1, scene1_ref, shadow, "scene1 recorded any")


def test_scene2_tumbling_cubes(dev2, scene2_ref):
    rays = random_rays(1 << 18, seed=21, center=(0, 4.0, 1.0), radius=30.0, target_radius=11.0, shadow_fraction=0.0)
    hits = _compare_closest(dev2, scene2_ref, rays, "scene2 random")
    assert len(np.unique(hits["shape"])) > 15
    rays = random_rays(1 << 17, seed=22, center=(0, 4.0, 1.0), radius=30.0, target_radius=11.0, shadow_fraction=0.8)
    _compare_any(dev2, scene2_ref, rays, "scene2 any")


def test_synthetic_mesh_scene(capi, scene5_host, scene5_ref):
    """The procedural displaced-sphere mesh of BASELINE.json configs[4] (small grid):
    two finite shapes besides the mesh -> top-level BVH of 4 shapes, deep face BVH."""
    dev = capi.DeviceScene(scene5_host.desc)
    rays = random_rays(1 << 17, seed=51, center=(0, 0.2, 0), radius=9.0, target_radius=2.6, shadow_fraction=0.3)
    hits = _compare_closest(dev, scene5_ref, rays, "synthetic")
    assert (hits["face"] >= 0).mean() > 0.4
    _compare_any(dev, scene5_ref, rays, "synthetic any")
    dev.close()


def test_stage6_scene(capi, scene6_host, scene6_ref):
    """Stage 6 rules (config C3): no transforms at all (-0.0 is not canonicalised), a face
    claims the hit at its first fan triangle, flat normals stay un-normalised."""
    assert scene6_host.desc.contents.semantics == 6
    dev = capi.DeviceScene(scene6_host.desc)
    rays = random_rays(1 << 18, seed=61, center=(0, -0.5, 0), radius=12.0, target_radius=4.0)
    hits = _compare_closest(dev, scene6_ref, rays, "stage6 random")
    assert set(np.unique(hits["shape"])) >= set(range(-1, 9))
    rays = random_rays(1 << 18, seed=62, center=(0.0, 0.0, 0.0), radius=6.0, target_radius=1.4)
    hits = _compare_closest(dev, scene6_ref, rays, "stage6 mesh focus")
    assert (hits["shape"] == 5).mean() > 0.3
    _compare_any(dev, scene6_ref, rays, "stage6 mesh focus any")
    # the box (shape 4): flat shading keeps the raw cross product as the normal
    rays = random_rays(1 << 16, seed=63, center=(0.5, -1.5, -1.5), radius=5.0, target_radius=0.8)
    hits = _compare_closest(dev, scene6_ref, rays, "stage6 box")
    assert (hits["shape"] == 4).mean() > 0.3
    rays = axis_parallel_rays(1 << 16, seed=64)
    _compare_closest(dev, scene6_ref, rays, "stage6 axis-parallel")
    _compare_any(dev, scene6_ref, rays, "stage6 axis-parallel any")
    rays = random_rays(1 << 17, seed=65, center=(0, -0.5, 0), radius=12.0, target_radius=4.0, shadow_fraction=0.7)
    _compare_any(dev, scene6_ref, rays, "stage6 any")
    dev.close()


def test_stage6_recorded_path_rays(capi, scene6_host, scene6_ref):
    """Every ray the Stage 6 reference casts while rendering a small frame."""
    dev = capi.DeviceScene(scene6_host.desc)
    spec = scene6_host.default_camera_spec()
    _img, stats = scene6_ref.render(spec, 96, 54, 2, ls=1, depth=3, record_rays=True)
    closest = scene6_ref.recorded_rays(0, capi.RAY_DTYPE)
    shadow = scene6_ref.recorded_rays(1, capi.RAY_DTYPE)
    assert len(closest) == stats.closest_calls and len(shadow) == stats.any_calls
    assert len(closest) > 50000 and len(shadow) > 20000
    _compare_closest(dev, scene6_ref, closest, "stage6 recorded closest")
    _compare_any(dev, scene6_ref, shadow, "stage6 recorded any")
    dev.close()


def test_edge_scenes(capi, ref, scene7_host, scene7_ref, scene8_host, scene8_ref):
    """Linear shape list instead of a top-level BVH, n-gon faces with per-vertex normals, scale
    keys (the general transform path: divide by the scale, rotate), a tilted rectangle light, a
    three-shape BVH, no lights -- and the empty set, where every ray misses."""
    dev = capi.DeviceScene(scene7_host.desc)
    assert scene7_host.desc.contents.num_top_nodes == 0
    rays = random_rays(1 << 17, seed=71, center=(0, -0.3, 0.2), radius=7.0, target_radius=2.2, shadow_fraction=0.3)
    hits = _compare_closest(dev, scene7_ref, rays, "edge scene (linear list)")
    assert (hits["face"] >= 0).mean() > 0.05 and set(np.unique(hits["tri"])) >= {0, 1, 2, 3}
    _compare_any(dev, scene7_ref, rays, "edge scene (linear list) any")
    _compare_closest(dev, scene7_ref, axis_parallel_rays(1 << 15, seed=72), "edge scene axis-parallel")
    dev.close()

    dev = capi.DeviceScene(scene8_host.desc)
    assert scene8_host.desc.contents.num_top_nodes == 5
    rays = random_rays(1 << 17, seed=81, center=(0.3, -0.5, 0), radius=9.0, target_radius=3.0, shadow_fraction=0.3)
    hits = _compare_closest(dev, scene8_ref, rays, "edge scene (scaled spheres)")
    assert set(np.unique(hits["shape"])) >= {-1, 0, 1, 2, 3}
    _compare_any(dev, scene8_ref, rays, "edge scene (scaled spheres) any")
    _compare_closest(dev, scene8_ref, axis_parallel_rays(1 << 15, seed=82), "edge scene (scaled spheres) axis-parallel")
    dev.close()

    empty_host = capi.HostScene(capi.RECIPE_EDGE_EMPTY)
    dev = capi.DeviceScene(empty_host.desc)
    rays = random_rays(4096, seed=91)
    hits = _compare_closest(dev, ref.RefScene(9), rays, "empty set")
    assert (hits["shape"] == -1).all()
    assert not dev.trace_any(rays).any()
    dev.close()


def test_empty_and_tiny_batches(dev1, capi):
    empty = np.zeros(0, capi.RAY_DTYPE)
    assert len(dev1.trace_closest(empty)) == 0
    assert len(dev1.trace_any(empty)) == 0
    one = random_rays(1, seed=3)
    assert len(dev1.trace_closest(one)) == 1
    for n in (31, 33, 127, 129):
        assert len(dev1.trace_any(random_rays(n, seed=n))) == n
