"""CPU tests that pin the oracle: the plain-C restatement (oracle/port.c), fed with
the PRODUCT's flattened scene, must give bit-identical hit records, sampler
permutations and camera rays to the compiled reference (oracle/_ref).  This checks
the restatement and, with no GPU involved, the host-side flattening."""
import numpy as np
import pytest

from tests.raybatches import axis_parallel_rays, bits, deep_scene_rays, random_rays


@pytest.fixture(scope="module")
def port():
    from oracle import portapi
    if not portapi.available():
        pytest.skip("oracle/_build/libport.so not built")
    return portapi


def _check(port, capi, host_scene, ref_scene, rays, label):
    mine = port.trace_closest(host_scene.desc, rays, capi.HITEX_DTYPE)
    theirs = ref_scene.trace_closest(rays)
    for f in ("shape", "face", "tri"):
        assert np.array_equal(mine[f], theirs[f]), "%s: %s differs" % (label, f)
    assert np.array_equal(bits(mine["t"]), bits(theirs["t"])), label
    hit = theirs["shape"] >= 0
    assert np.array_equal(bits(mine["normal"][hit]), bits(theirs["normal"][hit])), label
    assert np.array_equal(bits(mine["color_modifier"][hit]), bits(theirs["color_modifier"][hit, 0])), label
    assert np.array_equal(port.trace_any(host_scene.desc, rays), ref_scene.trace_any(rays)), label
    return mine


def test_port_hits_scene1(port, capi, scene1_host, scene1_ref):
    rays = random_rays(40000, seed=31, center=(0, -0.5, 0), radius=12.0, target_radius=4.0, shadow_fraction=0.3)
    hits = _check(port, capi, scene1_host, scene1_ref, rays, "scene1")
    assert set(np.unique(hits["shape"])) >= set(range(-1, 9))
    rays = random_rays(20000, seed=32, center=(0.1, 0, 0), radius=6.0, target_radius=1.4)
    hits = _check(port, capi, scene1_host, scene1_ref, rays, "scene1 mesh")
    assert (hits["face"] >= 0).mean() > 0.3
    _check(port, capi, scene1_host, scene1_ref, axis_parallel_rays(8000, seed=33), "scene1 axis-parallel")


def test_port_hits_scene2(port, capi, scene2_host, scene2_ref):
    rays = random_rays(40000, seed=41, center=(0, 4.0, 1.0), radius=30.0, target_radius=11.0, shadow_fraction=0.3)
    hits = _check(port, capi, scene2_host, scene2_ref, rays, "scene2")
    assert len(np.unique(hits["shape"])) > 15


def test_port_sample_stream_and_camera(port, capi, ref, scene1_host, scene1_ref):
    # literal Rng walk of the port == jump-ahead of the product (both against the same stream)
    for (w, h, depth, x, y) in [(64, 48, 3, 0, 0), (64, 48, 3, 17, 5), (64, 48, 3, 63, 47), (50, 31, 2, 49, 30), (9, 9, 1, 8, 8)]:
        assert np.array_equal(port.pixel_permutations(w, h, depth, x, y), capi.sample_permutations(w, h, depth, x, y))
    assert port.pixel_permutations(3, 2, 2, 2, 0) is None
    # camera rays: port vs the reference's recorded primary rays
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps, depth = 24, 16, 2, 2
    scene1_ref.render(spec, W, H, ps, ls=1, depth=depth, record_rays=True)
    rec = scene1_ref.recorded_rays(0, capi.RAY_DTYPE)
    origin = np.array(spec[1:4], np.float32)
    primary = rec[(rec["origin"] == origin).all(axis=1) & (rec["tmax"] == np.float32(1e30))]
    mine = np.array([port.camera_ray(cam, W, H, ps, depth, x, y, psi, capi.RAY_DTYPE)
                     for y in range(H) for x in range(W) for psi in range(ps * ps)])
    a = np.ascontiguousarray(mine).view(np.uint32).reshape(-1, 8)
    b = np.ascontiguousarray(primary).view(np.uint32).reshape(-1, 8)
    assert a.shape == b.shape
    assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])


def test_port_hits_deep_scenes(port, capi, deep_host, deep_ref, deepboth_host, deepboth_ref):
    """The wedge scenes (face BVH 42 / 46 deep, top level 18 deep): restatement == reference, and
    the batches really reach stack depths beyond 32 (what the deep-stack GPU kernels are for)."""
    for host, refscene, label in ((deep_host, deep_ref, "deep"), (deepboth_host, deepboth_ref, "deep both")):
        rays = deep_scene_rays(30000, seed=51)
        port.work_reset()
        hits = _check(port, capi, host, refscene, rays, label)
        work = port.work_counters()
        assert (hits["face"] >= 0).sum() > 1000
        assert work["max_stack_mesh"] > 32, work
        ap = axis_parallel_rays(6000, seed=52)
        ap["origin"] *= np.float32(0.25)
        _check(port, capi, host, refscene, ap, label + " axis-parallel")
    assert work["max_stack_top"] > 8


def test_port_work_counters(port, capi, scene1_host):
    """Counter definitions (oracle/port.c): one ray against scene 1 pops at least the top-level
    root, evaluates the set's keyless transform and the plane's single-key one; counts add up over
    a batch and reset."""
    rays = random_rays(2000, seed=61, center=(0, -0.5, 0), radius=12.0, target_radius=4.0)
    port.work_reset()
    port.trace_closest(scene1_host.desc, rays[:1], capi.HITEX_DTYPE)
    one = port.work_counters()
    assert one["node_pops"] >= 1 and one["shape_tests"] >= 1            # the plane is always tested
    assert one["xform_evals"] >= 2 and one["xform_keyed"] == one["xform_evals"] - 1   # set transform: keyless
    port.work_reset()
    port.trace_closest(scene1_host.desc, rays, capi.HITEX_DTYPE)
    a = port.work_counters()
    port.trace_closest(scene1_host.desc, rays, capi.HITEX_DTYPE)
    b = port.work_counters()
    for k in ("node_pops", "tri_tests", "shape_tests", "xform_evals", "xform_keyed", "xform_pairs"):
        assert b[k] == 2 * a[k] and a[k] > 0
    assert a["xform_pairs"] < a["xform_keyed"] < a["xform_evals"]
    assert 4.0 < a["node_pops"] / len(rays) < 40.0


def test_grazing_camera_reaches_deep_stacks(port, capi, deep_host, deep_ref):
    """The camera the GPU deep-tree render test uses (tests/test_gpu_deep.py GRAZE_CAMERA): the
    rays of that frame hold more than 32 live stack entries in the reference's traversal."""
    spec = np.array([1e-8, -0.8, -0.5, 0.0, 1.5, -0.5, 0.0, 0, 1, 0, 16, 0, 0, 0], np.float32)
    deep_ref.render(spec, 64, 36, 2, ls=1, depth=3, record_rays=True)
    closest = deep_ref.recorded_rays(0, capi.RAY_DTYPE)
    port.work_reset()
    port.trace_closest(deep_host.desc, closest, capi.HITEX_DTYPE)
    assert port.work_counters()["max_stack_mesh"] > 32
