"""The roofline's work counters are pinned to the oracle.

bench.py's `roofline.achieved` is ALGORITHMIC bytes over time, and the bytes are made of
counts -- BVH nodes popped, triangles tested, analytic shapes tested, keyed transforms
evaluated per ray.  The CUDA kernels count their own work
(RT_RENDER_COUNT_WORK, rt_trace_*_counted); here those counts must EQUAL what the oracle's
restatement of the reference's traversal (oracle/port.c, itself pinned hit for hit to the
compiled reference) does on the very same rays:

  * seeded ray batches through rt_trace_closest_counted / rt_trace_any_counted (unified kernel);
  * every ray the reference casts while rendering a frame (recorded by oracle/_ref), against the
    counters of the render of that frame by the default split passes (tabulated top-level walk,
    face-BVH pass, resume pass), the per-lane top-level pass and the unified kernel."""
import numpy as np
import pytest

from tests.raybatches import axis_parallel_rays, deep_scene_rays, random_rays

pytestmark = pytest.mark.gpu

FIELDS = ("node_pops", "tri_tests", "shape_tests", "xform_evals", "xform_keyed", "xform_pairs")


@pytest.fixture(scope="module")
def port():
    from oracle import portapi
    if not portapi.available():
        pytest.skip("oracle/_build/libport.so not built")
    return portapi


def _oracle_work(port, capi, host, closest=None, shadow=None):
    port.work_reset()
    if closest is not None and len(closest):
        port.trace_closest(host.desc, closest, capi.HITEX_DTYPE)
    if shadow is not None and len(shadow):
        port.trace_any(host.desc, shadow)
    return port.work_counters()


@pytest.mark.parametrize("which,make_rays", [
    ("scene1", lambda: random_rays(1 << 16, seed=201, center=(0, -0.5, 0), radius=12.0, target_radius=4.0, shadow_fraction=0.4)),
    ("scene1", lambda: axis_parallel_rays(1 << 14, seed=202)),
    ("scene2", lambda: random_rays(1 << 16, seed=203, center=(0, 4.0, 1.0), radius=30.0, target_radius=11.0, shadow_fraction=0.4)),
    ("scene5", lambda: random_rays(1 << 16, seed=204, center=(0, 0.2, 0), radius=9.0, target_radius=2.6, shadow_fraction=0.4)),
    ("scene7", lambda: random_rays(1 << 15, seed=205, center=(0, -0.3, 0.2), radius=7.0, target_radius=2.2, shadow_fraction=0.4)),
    ("deep", lambda: deep_scene_rays(1 << 16, seed=206)),
    ("deepboth", lambda: deep_scene_rays(1 << 16, seed=207)),
])
def test_batch_counters_equal_oracle(which, make_rays, request, capi, port):
    host = request.getfixturevalue(which + "_host")
    rays = make_rays()
    dev = capi.DeviceScene(host.desc)
    _h, mine_c = dev.trace_counted(rays, any_hit=False)
    _s, mine_a = dev.trace_counted(rays, any_hit=True)
    dev.close()
    want_c = _oracle_work(port, capi, host, closest=rays)
    want_a = _oracle_work(port, capi, host, shadow=rays)
    for f in FIELDS:
        assert mine_c[f] == want_c[f], "closest %s: kernel %d, oracle %d" % (f, mine_c[f], want_c[f])
        assert mine_a[f] == want_a[f], "any-hit %s: kernel %d, oracle %d" % (f, mine_a[f], want_a[f])
    assert want_c["node_pops"] > len(rays) and want_c["xform_evals"] >= len(rays)
    if which.startswith("deep"):
        # the batch really needs the deep stacks (reference stack entries, RAccel.h:379)
        assert want_c["max_stack_mesh"] > 32
    if which == "deepboth":
        assert want_c["max_stack_top"] > 8


@pytest.mark.parametrize("which,W,H,ps,ls,depth", [
    ("scene1", 128, 72, 2, 1, 3),
    ("scene1", 64, 36, 2, 2, 2),
    ("scene2", 96, 54, 2, 1, 3),
    ("scene5", 80, 45, 2, 1, 3),
    ("scene7", 72, 40, 2, 1, 3),
    ("deep", 64, 36, 2, 1, 3),
])
def test_render_counters_equal_oracle_on_recorded_rays(which, W, H, ps, ls, depth, request, capi, port):
    """bench.py's roofline reads exactly these counters (one RT_RENDER_COUNT_WORK step)."""
    host = request.getfixturevalue(which + "_host")
    refscene = request.getfixturevalue(which + "_ref")
    spec = host.default_camera_spec()
    _img, rstats = refscene.render(spec, W, H, ps, ls=ls, depth=depth, record_rays=True)
    closest = refscene.recorded_rays(0, capi.RAY_DTYPE)
    shadow = refscene.recorded_rays(1, capi.RAY_DTYPE)
    assert len(closest) == rstats.closest_calls and len(shadow) == rstats.any_calls
    want = _oracle_work(port, capi, host, closest=closest, shadow=shadow)
    dev = capi.DeviceScene(host.desc)
    cam = capi.camera_from_spec(spec)
    modes = [("split, tabulated top level (default)", {}), ("split, per-lane top level", {"dynamic_top": True}),
             ("unified kernel", {"unified": True})]
    for label, flags in modes:
        _mine, stats = dev.render(cam, W, H, ps, ls=ls, depth=depth, count_work=True, **flags)
        assert stats.closest_rays == len(closest) and stats.any_rays == len(shadow), label
        for f in FIELDS:
            assert getattr(stats, f) == want[f], "%s, %s: kernel %d, oracle %d" % (label, f, getattr(stats, f), want[f])
    dev.close()
    assert want["xform_keyed"] < want["xform_evals"]       # the set's own transform is keyless
