"""GPU parity tests of the deep-stack traversal kernels.

The face-BVH pass is instantiated for trees that need at most 32 stack entries and for deeper
ones (k_split_mesh<64>: 14 shared-memory stack slots per lane, the rest in local memory), and
the unified kernel behind rt_trace_closest / rt_trace_any exists for 32, 64 and 104 combined
entries.  The GUI scenes are 20 deep; config C5 (10 M triangles) is 37 deep.  These tests run
the deep instantiations against the compiled reference: the wedge scenes are 42-46 deep over a
few hundred faces (and, with the sphere chain, 18 deep at the top level too: 66 combined
entries), the displaced sphere at 1 M quads is the C5 mesh family at depth 32, and -- marked
slow but still part of `-m gpu` -- C5's own 2236 x 2236 grid (depth 37)."""
import numpy as np
import pytest

from tests.conftest import BIG_GRID, DEEP_GRID, DEEPER_GRID
from tests.raybatches import axis_parallel_rays, bits, deep_scene_rays, random_rays
from tests.test_gpu_trace import _compare_any, _compare_closest

pytestmark = pytest.mark.gpu

TIP_CAMERA = np.array([12, 0.3, 0.6, 2.2, -0.45, -0.5, 0.0, 0, 1, 0, 16, 0, 0, 1], np.float32)
# Behind the wedge's tip, looking along its axis through a field of view of 1e-8 degrees with the
# shutter closed at time 0: every camera ray grazes the wedge and holds up to 42 live stack
# entries in the reference's traversal (counted by the oracle, tests/test_oracle_port.py) -- far
# beyond the 14 shared-memory slots of k_split_mesh<64>, so its local-memory spill runs too.
GRAZE_CAMERA = np.array([1e-8, -0.8, -0.5, 0.0, 1.5, -0.5, 0.0, 0, 1, 0, 16, 0, 0, 0], np.float32)


def _images_equal(capi, host, refscene, spec, W, H, ps, ls, depth, label, **flags):
    dev = capi.DeviceScene(host.desc)
    cam = capi.camera_from_spec(spec)
    theirs, rstats = refscene.render(spec, W, H, ps, ls=ls, depth=depth)
    mine, stats = dev.render(cam, W, H, ps, ls=ls, depth=depth, **flags)
    dev.close()
    assert np.array_equal(bits(mine), bits(theirs)), "%s: %.4f of the pixels bit-identical" % (
        label, (bits(mine) == bits(theirs)).all(axis=-1).mean())
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls, label
    return stats


def test_deep_mesh_split_kernels(capi, deep_host, deep_ref):
    """Face BVH 42 deep, top level 3 deep: the split path with k_split_mesh<64> (render) and
    trace_wave<64> (ray batches)."""
    assert deep_host.depth(0) >= 40 and deep_host.depth(-1) <= 7
    dev = capi.DeviceScene(deep_host.desc)
    rays = deep_scene_rays(1 << 18, seed=101)
    hits = _compare_closest(dev, deep_ref, rays, "deep mesh")
    assert (hits["face"] >= 0).sum() > 10000
    # hits spread over the whole depth of the wedge, down to its last rows
    assert len(np.unique(hits["face"][hits["face"] >= 0] // DEEP_GRID[1])) >= DEEP_GRID[0] - 4
    _compare_any(dev, deep_ref, rays, "deep mesh any")
    ap = axis_parallel_rays(1 << 15, seed=102)
    ap["origin"] *= np.float32(0.25)        # bring the grid of exact coordinates down to the wedge's scale
    _compare_closest(dev, deep_ref, ap, "deep mesh axis-parallel")
    _compare_any(dev, deep_ref, ap, "deep mesh axis-parallel any")
    dev.close()


@pytest.mark.parametrize("camera", ["default", "tip", "graze"])
def test_deep_mesh_images(capi, deep_host, deep_ref, camera):
    spec = {"default": deep_host.default_camera_spec(), "tip": TIP_CAMERA, "graze": GRAZE_CAMERA}[camera]
    _images_equal(capi, deep_host, deep_ref, spec, 96, 54, 3, 1, 3, "deep mesh, %s camera" % camera)
    _images_equal(capi, deep_host, deep_ref, spec, 48, 27, 2, 2, 2, "deep mesh, %s camera, unified" % camera, unified=True)


def test_deep_mesh_and_deep_top_level(capi, deepboth_host, deepboth_ref):
    """Face BVH 46 deep AND top level 18 deep: 66 combined stack entries, so ray batches and
    renders go through the unified kernel's 104-entry instantiation."""
    assert deepboth_host.depth(0) >= 44 and deepboth_host.depth(-1) >= 17
    assert deepboth_host.depth(0) + deepboth_host.depth(-1) + 2 > 64
    dev = capi.DeviceScene(deepboth_host.desc)
    rays = deep_scene_rays(1 << 18, seed=111)
    hits = _compare_closest(dev, deepboth_ref, rays, "deep both")
    assert len(np.unique(hits["shape"])) >= 15
    _compare_any(dev, deepboth_ref, rays, "deep both any")
    ap = axis_parallel_rays(1 << 15, seed=112)
    ap["origin"] *= np.float32(0.25)
    _compare_closest(dev, deepboth_ref, ap, "deep both axis-parallel")
    dev.close()
    _images_equal(capi, deepboth_host, deepboth_ref, deepboth_host.default_camera_spec(), 96, 54, 3, 1, 3, "deep both")
    _images_equal(capi, deepboth_host, deepboth_ref, TIP_CAMERA, 64, 36, 2, 1, 4, "deep both, tip camera")
    _images_equal(capi, deepboth_host, deepboth_ref, GRAZE_CAMERA, 64, 36, 2, 1, 3, "deep both, grazing camera")


@pytest.fixture(scope="module")
def big_host(capi):
    return capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, BIG_GRID)


@pytest.fixture(scope="module")
def big_ref(ref):
    return ref.RefScene(5, None, BIG_GRID)


def test_displaced_sphere_1m_quads(capi, big_host, big_ref):
    """The C5 mesh family at 1 M quads (2 M triangles): depth 32 -> 33 stack entries, the first
    grid size that selects the deep face-BVH pass.  2^20 seeded rays, closest and any hit, plus
    the axis-parallel batch, and one small render."""
    assert big_host.depth(0) >= 32
    dev = capi.DeviceScene(big_host.desc)
    rays = random_rays(1 << 20, seed=121, center=(0, 0.2, 0), radius=9.0, target_radius=2.6, shadow_fraction=0.3)
    hits = _compare_closest(dev, big_ref, rays, "1 M quads")
    assert (hits["face"] >= 0).mean() > 0.4
    _compare_any(dev, big_ref, rays, "1 M quads any")
    ap = axis_parallel_rays(1 << 16, seed=122)
    _compare_closest(dev, big_ref, ap, "1 M quads axis-parallel")
    _compare_any(dev, big_ref, ap, "1 M quads axis-parallel any")
    dev.close()
    _images_equal(capi, big_host, big_ref, big_host.default_camera_spec(), 96, 54, 2, 1, 3, "1 M quads")


def test_config_c5_mesh_hit_ids(capi, ref):
    """BASELINE.json configs[4] by name: the 2236 x 2236 grid = 9 999 392 triangles, face BVH 37
    deep.  Hit ids (shape, face, fan triangle) and t bit-exact on 2^19 seeded rays (closest and any
    hit) and on the axis-parallel batch, and a 64 x 36 x 4 spp image bit-identical.  The reference
    needs ~3 s to build this BVH and traces ~0.5 M rays/s."""
    grid = (2236, 2236)
    host = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, grid)
    assert host.depth(0) == 37 and host.desc.contents.num_faces == 4999696
    refscene = ref.RefScene(5, None, grid)
    dev = capi.DeviceScene(host.desc)
    rays = random_rays(1 << 19, seed=131, center=(0, 0.2, 0), radius=9.0, target_radius=2.6, shadow_fraction=0.3)
    hits = _compare_closest(dev, refscene, rays, "C5 mesh")
    assert (hits["face"] >= 0).mean() > 0.4
    _compare_any(dev, refscene, rays, "C5 mesh any")
    ap = axis_parallel_rays(1 << 15, seed=132)
    _compare_closest(dev, refscene, ap, "C5 mesh axis-parallel")
    dev.close()
    _images_equal(capi, host, refscene, host.default_camera_spec(), 64, 36, 2, 1, 3, "C5 mesh")
