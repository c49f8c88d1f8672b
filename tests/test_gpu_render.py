"""GPU parity tests of the wavefront renderer against the compiled reference's own
raytrace() on the same scene, resolution and sample counts.

Bars (BASELINE.json north_star): camera rays bit-exact (the counter-based sample
stream reproduces the reference's); Monte-Carlo images within a stated per-pixel
RMSE.  Since the device also reproduces the C library's sinf/cosf/powf bit for bit
(rt_libm.cuh) every sample takes the reference's path, so the float images come out
BIT-IDENTICAL and the ray counts equal; that is what is asserted here, the RMSE bar
being the contract's fallback."""
import numpy as np
import pytest

from tests.raybatches import bits

pytestmark = pytest.mark.gpu

# Stated tolerances for Monte-Carlo stages (linear float RGB, per pixel, relative to
# the image's mean luminance), and for the 8-bit display image (gamma 2.2).
RMSE_REL_TOL = 0.0
MIN_IDENTICAL_FRACTION = 1.0


def _rays_as_rows(rays):
    return np.ascontiguousarray(rays).view(np.uint32).reshape(len(rays), 8)


def _sorted_rows(rows):
    return rows[np.lexsort(rows.T[::-1])]


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


def _compare_images(mine, theirs, label):
    assert mine.shape == theirs.shape
    assert not np.isnan(mine).any(), label + ": NaN pixels"
    same = (bits(mine) == bits(theirs)).all(axis=-1)
    rmse = float(np.sqrt(np.mean((mine.astype(np.float64) - theirs.astype(np.float64)) ** 2)))
    rel = rmse / max(float(theirs.mean()), 1e-12)
    g_m = np.clip(np.power(np.clip(mine, 0, None), 1 / 2.2), 0, 1) * 255.0
    g_t = np.clip(np.power(np.clip(theirs, 0, None), 1 / 2.2), 0, 1) * 255.0
    diff8 = np.abs(g_m.astype(np.uint8).astype(int) - g_t.astype(np.uint8).astype(int))
    print("%s: identical pixels %.4f, rmse %.3e (rel %.3e), 8-bit max diff %d, 8-bit pixels differing %.4f" % (
        label, same.mean(), rmse, rel, diff8.max(), (diff8.max(axis=-1) > 0).mean()))
    return same.mean(), rel


def test_camera_rays_reproduce_reference_stream(dev1, scene1_ref, scene1_host, capi):
    """Every camera ray the reference casts for a frame, found among its recorded
    rays by origin, must be generated bit-for-bit by the counter-based stream (MWC
    jump-ahead + CMJ), for every pixel and every pixel-sample index."""
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 72, 44, 3      # 4x4 chunks of 18x11
    _img, _stats = scene1_ref.render(spec, W, H, ps, ls=1, depth=3, record_rays=True)
    rec = scene1_ref.recorded_rays(0, capi.RAY_DTYPE)
    origin = np.array(spec[1:4], np.float32)
    primary = rec[(rec["origin"] == origin).all(axis=1) & (rec["tmax"] == np.float32(1e30))]
    assert len(primary) == W * H * ps * ps
    mine = np.concatenate([dev1.camera_rays(cam, W, H, ps, psi) for psi in range(ps * ps)])
    assert np.array_equal(_sorted_rows(_rays_as_rows(mine)), _sorted_rows(_rays_as_rows(primary)))


def test_camera_rays_uneven_chunks(dev1, scene1_ref, scene1_host, capi):
    """Image sizes that do not divide by four make a fifth row/column of chunks."""
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 50, 31, 2
    scene1_ref.render(spec, W, H, ps, ls=1, depth=1, record_rays=True)
    rec = scene1_ref.recorded_rays(0, capi.RAY_DTYPE)
    origin = np.array(spec[1:4], np.float32)
    primary = rec[(rec["origin"] == origin).all(axis=1)]
    mine = np.concatenate([dev1.camera_rays(cam, W, H, ps, psi, depth=1) for psi in range(ps * ps)])
    assert np.array_equal(_sorted_rows(_rays_as_rows(mine)), _sorted_rows(_rays_as_rows(primary)))


@pytest.mark.parametrize("W,H,ps,ls,depth", [(128, 72, 4, 1, 3), (64, 36, 2, 2, 4), (40, 24, 1, 1, 1), (3, 2, 2, 1, 2)])
def test_scene1_image_matches_reference(dev1, scene1_ref, scene1_host, capi, W, H, ps, ls, depth):
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    theirs, rstats = scene1_ref.render(spec, W, H, ps, ls=ls, depth=depth)
    mine, stats = dev1.render(cam, W, H, ps, ls=ls, depth=depth)
    same, rel = _compare_images(mine, theirs, "scene1 %dx%d ps%d ls%d d%d" % (W, H, ps, ls, depth))
    if W >= 4 and H >= 4:
        assert stats.samples == W * H * ps * ps
    # ray counts: a "ray" is one scene.intersect / doesIntersect call of pathTrace
    ref_rays = rstats.closest_calls + rstats.any_calls
    my_rays = stats.closest_rays + stats.any_rays
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls, (my_rays, ref_rays)
    assert rel <= RMSE_REL_TOL
    assert same >= MIN_IDENTICAL_FRACTION


def test_scene2_image_matches_reference(dev2, scene2_ref, scene2_host, capi):
    spec = scene2_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 96, 64, 3
    theirs, rstats = scene2_ref.render(spec, W, H, ps, ls=1, depth=3)
    mine, stats = dev2.render(cam, W, H, ps, ls=1, depth=3)
    same, rel = _compare_images(mine, theirs, "scene2")
    assert rel <= RMSE_REL_TOL
    assert same >= MIN_IDENTICAL_FRACTION


def test_synthetic_mesh_image_matches_reference(capi, scene5_host, scene5_ref):
    dev = capi.DeviceScene(scene5_host.desc)
    spec = scene5_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 80, 45, 2
    theirs, rstats = scene5_ref.render(spec, W, H, ps, ls=1, depth=3)
    mine, stats = dev.render(cam, W, H, ps, ls=1, depth=3)
    dev.close()
    same, rel = _compare_images(mine, theirs, "synthetic mesh")
    assert rel <= RMSE_REL_TOL and same >= MIN_IDENTICAL_FRACTION
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls


def test_scene1_full_sample_count_matches_reference(dev1, scene1_ref, scene1_host, capi):
    """The bench configuration's sample count (256 spp, ls 1, depth 3) at a resolution the
    CPU reference finishes in seconds: 8.3 M samples, every pixel bit-identical."""
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 240, 135, 16
    theirs, rstats = scene1_ref.render(spec, W, H, ps, ls=1, depth=3)
    mine, stats = dev1.render(cam, W, H, ps, ls=1, depth=3)
    same, rel = _compare_images(mine, theirs, "scene1 240x135 256spp")
    assert same == 1.0 and rel == 0.0
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls


def test_depth_of_field_matches_reference(dev1, scene1_ref, scene1_host, capi):
    """Thin-lens camera (RaytraceMain.cpp:237-264): lens radius > 0, focal distance 16."""
    spec = scene1_host.default_camera_spec().copy()
    spec[11] = 0.35      # lens radius
    cam = capi.camera_from_spec(spec)
    W, H, ps = 96, 54, 3
    theirs, rstats = scene1_ref.render(spec, W, H, ps, ls=1, depth=2)
    mine, stats = dev1.render(cam, W, H, ps, ls=1, depth=2)
    same, rel = _compare_images(mine, theirs, "scene1 depth of field")
    assert same == 1.0 and rel == 0.0
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls


def test_mesh_light_matches_reference(capi, ref, obj_path):
    """bumpy.obj as an area light: face pick by area CDF, triangle sampling, Mesh::pdfSA
    (RMesh.h:133-196), through a ShapeLight."""
    host = capi.HostScene(capi.RECIPE_STAGE7_SCENE1_MESHLIGHT, obj_path)
    refscene = ref.RefScene(3, obj_path)
    dev = capi.DeviceScene(host.desc)
    spec = host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 80, 45, 2
    theirs, rstats = refscene.render(spec, W, H, ps, ls=2, depth=2)
    mine, stats = dev.render(cam, W, H, ps, ls=2, depth=2)
    dev.close()
    same, rel = _compare_images(mine, theirs, "scene1 mesh light")
    assert same == 1.0 and rel == 0.0
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls


# Stage 6 (config C3) draws its light and BRDF samples from one serial, data-dependent
# Rng per image chunk (S6 RaytraceMain.cpp:57, 276-277, 315), which no parallel renderer
# can replay; there the contract's Monte-Carlo bar applies: same estimator, images equal
# within noise.  Bars: channel means of the whole frame within 1 % (330 k samples: the
# frame mean's own noise is ~0.3 %), 6x6-pixel block means within 6 % RMS of the mean
# luminance at 64 spp, and ray counts per sample within 1 %.
S6_MEAN_TOL = 0.01
S6_BLOCK_RMS_TOL = 0.06
S6_RAYS_TOL = 0.01


def _block_means(img, b):
    h, w, _ = img.shape
    return img[:h - h % b, :w - w % b].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))


@pytest.mark.parametrize("ps,ls,depth", [(8, 1, 3), (6, 2, 2), (8, 1, 1)])
def test_stage6_image_matches_reference_within_noise(capi, scene6_host, scene6_ref, ps, ls, depth):
    dev = capi.DeviceScene(scene6_host.desc)
    spec = scene6_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H = 96, 54
    theirs, rstats = scene6_ref.render(spec, W, H, ps, ls=ls, depth=depth)
    mine, stats = dev.render(cam, W, H, ps, ls=ls, depth=depth)
    dev.close()
    assert not np.isnan(mine).any()
    mean_t, mean_m = theirs.mean(axis=(0, 1)), mine.mean(axis=(0, 1))
    rel_mean = np.abs(mean_m - mean_t) / mean_t
    bm, bt = _block_means(mine.astype(np.float64), 6), _block_means(theirs.astype(np.float64), 6)
    block_rms = float(np.sqrt(np.mean((bm - bt) ** 2)) / theirs.mean())
    ref_rays = rstats.closest_calls + rstats.any_calls
    my_rays = stats.closest_rays + stats.any_rays
    print("stage6 ps%d ls%d d%d: channel means ref %s mine %s (rel %s), block rms %.4f, rays ref %d mine %d" % (
        ps, ls, depth, mean_t, mean_m, rel_mean, block_rms, ref_rays, my_rays))
    assert stats.samples == W * H * ps * ps
    assert (rel_mean <= S6_MEAN_TOL).all()
    assert block_rms <= S6_BLOCK_RMS_TOL
    assert abs(my_rays - ref_rays) / ref_rays <= S6_RAYS_TOL
    # first-bounce camera rays hit the same things: the primary-visibility part of the
    # image (emitters seen directly) is noise-free up to pixel jitter
    if depth == 1:
        assert stats.closest_rays >= W * H * ps * ps


@pytest.mark.parametrize("recipe,ls,depth", [(7, 2, 3), (8, 1, 4), (9, 1, 2)])
def test_edge_scene_images_match_reference(capi, ref, recipe, ls, depth):
    """Edge-case scenes (fixtures/scene_recipes.h buildEdgeScene): linear shape list with n-gon
    faces, scale keys and a tilted light; no lights (black image, mirror bounces still traced);
    the empty set."""
    host = capi.HostScene(recipe)
    refscene = ref.RefScene(recipe)
    dev = capi.DeviceScene(host.desc)
    spec = host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 72, 40, 3
    theirs, rstats = refscene.render(spec, W, H, ps, ls=ls, depth=depth)
    mine, stats = dev.render(cam, W, H, ps, ls=ls, depth=depth)
    dev.close()
    same, rel = _compare_images(mine, theirs, "edge scene %d" % recipe)
    assert same == 1.0 and rel == 0.0
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls
    if recipe != 7:
        assert not mine.any()


def test_tile_sharding_is_exact(dev1, scene1_host, capi):
    """Any partition of the image into rank-owned tiles reproduces the single-GPU
    image bit for bit (the sample stream is position-addressable), and small
    batches give the same image as one big batch."""
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 100, 60, 2
    whole, _ = dev1.render(cam, W, H, ps)
    for world in (2, 3):
        merged = np.full((H, W, 3), -1.0, np.float32)
        total = 0
        for rank in range(world):
            _, st = dev1.render(cam, W, H, ps, rank=rank, world=world, tile_size=16, out=merged)
            total += st.samples
        assert total == W * H * ps * ps
        assert np.array_equal(bits(merged), bits(whole))
    small, _ = dev1.render(cam, W, H, ps, tile_size=8, max_batch_samples=1024)
    assert np.array_equal(bits(small), bits(whole))


def test_packed_tiles_gather_and_scatter(dev1, scene1_host, capi):
    """The multi-GPU tile assembly, rank by rank on one device: every rank's tiles rendered into
    its packed buffer (rt_render_tiles_packed), each packed buffer scattered into the frame
    (rt_unpack_tiles) -- what rt_render_multi does around its ncclSend / ncclRecv -- gives the
    single-GPU image bit for bit, including tiles that hang over the image edge and a rank that
    owns no tile at all; rt_render_multi itself on a one-rank communicator renders in place."""
    import ctypes as C
    import torch
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 100, 60, 2
    whole, _ = dev1.render(cam, W, H, ps)
    lib = capi.core()
    for world, tile in ((2, 16), (3, 32), (5, 64), (2, 0)):
        frame = torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda:0")
        total = 0
        for rank in range(world):
            n = capi.packed_floats(W, H, world, rank, tile)
            ts = tile if tile else 32
            owners, _ = capi.tile_owners(W, H, world, tile)
            assert n == int((owners == rank).sum()) * ts * ts * 3
            if n == 0:
                continue
            packed = torch.full((n,), -2.0, dtype=torch.float32, device="cuda:0")
            params = capi.RtRenderParams(W, H, ps, 1, 3, tile, rank, world, 1024 if rank == 0 else 0, 0)
            stats = capi.RtRenderStats()
            capi.check(lib.rt_render_tiles_packed(dev1.handle, C.byref(cam), C.byref(params), packed.data_ptr(), n,
                                                  C.byref(stats), None), "rt_render_tiles_packed")
            total += stats.samples
            capi.check(lib.rt_unpack_tiles(0, packed.data_ptr(), W, H, tile, world, rank, frame.data_ptr(), None),
                       "rt_unpack_tiles")
        torch.cuda.synchronize()
        assert total == W * H * ps * ps
        assert np.array_equal(bits(frame.cpu().numpy()), bits(whole)), (world, tile)
    comm = capi.Comm(bytes(capi.Comm.ID_BYTES), 0, 1, 0)
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
    params = capi.RtRenderParams(W, H, ps, 1, 3, 0, 0, 1, 0, 0)
    stats, _ms = comm.render_multi(dev1, cam, params, frame.data_ptr())
    comm.close()
    assert stats.samples == W * H * ps * ps
    assert np.array_equal(bits(frame.cpu().numpy()), bits(whole))


@pytest.mark.parametrize("which", ["scene1", "scene2", "scene7", "scene8"])
def test_split_and_unified_traversal_agree(which, request, capi):
    """The split top-level / mesh passes (default) and the single unified kernel must
    trace the very same rays through the very same nodes: identical images, ray
    counts and work counters."""
    host = request.getfixturevalue(which + "_host")
    dev = capi.DeviceScene(host.desc)
    cam = capi.camera_from_spec(host.default_camera_spec())
    W, H, ps = 96, 54, 3
    img_s, st_s = dev.render(cam, W, H, ps, ls=1, depth=3, count_work=True)
    img_d, st_d = dev.render(cam, W, H, ps, ls=1, depth=3, count_work=True, dynamic_top=True)
    img_u, st_u = dev.render(cam, W, H, ps, ls=1, depth=3, count_work=True, unified=True)
    dev.close()
    assert np.array_equal(bits(img_s), bits(img_u)) and np.array_equal(bits(img_d), bits(img_u))
    # the tabulated top-level walk (default), the per-lane top-level pass and the unified kernel
    # pop the same nodes and test the same shapes
    for field in ("closest_rays", "any_rays", "node_pops", "tri_tests", "shape_tests", "xform_evals"):
        assert getattr(st_s, field) == getattr(st_u, field), field
        assert getattr(st_d, field) == getattr(st_u, field), field
    if which != "scene8":        # no mesh in that scene: split mode does not apply
        assert st_s.kernel_launches > st_u.kernel_launches


def test_work_counters(dev1, scene1_host, capi):
    spec = scene1_host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    img_a, plain = dev1.render(cam, 64, 36, 2)
    img_b, counted = dev1.render(cam, 64, 36, 2, count_work=True)
    assert np.array_equal(bits(img_a), bits(img_b))
    assert counted.closest_rays == plain.closest_rays and counted.any_rays == plain.any_rays
    rays = counted.closest_rays + counted.any_rays
    assert 4.0 < counted.node_pops / rays < 40.0
    assert counted.tri_tests > 0 and counted.shape_tests > 0 and counted.xform_evals > rays


def test_tonemap_matches_display_image(capi):
    rng = np.random.RandomState(5)
    rgb = rng.uniform(0, 2.0, size=(257, 3)).astype(np.float32)
    rgb[3] = (-0.1, 0.5, 0.5)          # negative -> green
    rgb[4] = (np.nan, 0.5, 0.5)        # NaN -> blue
    out = capi.tonemap_bgra8(rgb, exposure_stops=0.0, gamma=2.2)
    assert tuple(out[3]) == (0, 255, 0, 255)
    assert tuple(out[4]) == (255, 0, 0, 255)
    ok = np.ones(len(rgb), bool)
    ok[3] = ok[4] = False
    expect = np.clip(np.power(rgb[ok].astype(np.float64), 1 / np.float32(2.2)), 0, 1)
    got = out[ok][:, [2, 1, 0]].astype(int)
    assert np.abs(got - np.floor(expect * 255.0)).max() <= 1


def test_raytrace_call_and_device_shards_match_reference(capi, scene1_ref, obj_path):
    """The reference-facing call itself: Rayito::raytrace() on an application-built scene
    (rth_app_raytrace: findLights + prepare + flatten + upload + render + download per call,
    RaytraceMain.cpp:485-579) returns the reference's image bit for bit, repeated calls on the
    same scene agree (prepare() is redone each time, as in the reference), and
    rayito_b200::raytraceToDevice() per rank leaves tiles in HBM whose sum over ranks -- the
    multi-GPU tile assembly -- is that same image."""
    import ctypes as C
    import torch
    lib = capi.host()
    app = lib.rth_app_create(capi.RECIPE_STAGE7_SCENE1, obj_path.encode(), 0, 0)
    assert app
    try:
        host_scene = capi.HostScene(capi.RECIPE_STAGE7_SCENE1, obj_path)
        spec = host_scene.default_camera_spec()
        W, H, ps, ls, depth = 96, 54, 3, 1, 3
        theirs, ref_stats = scene1_ref.render(spec, W, H, ps, ls=ls, depth=depth)
        stats = capi.RtRenderStats()
        for _ in range(2):
            img = np.zeros((H, W, 3), np.float32)
            rc = lib.rth_app_raytrace(app, spec.ctypes.data, W, H, ps, ls, depth, 0, 0, 1, 0, img.ctypes.data, 0, C.byref(stats))
            assert rc == 0, lib.rth_last_error_string()
            assert np.array_equal(bits(img), bits(theirs))
            assert stats.closest_rays + stats.any_rays == ref_stats.closest_calls + ref_stats.any_calls
        # the frame handed back in place (Image storage is recycled between calls; a different
        # size in between must not confuse the spare block)
        pixels = C.c_void_p()
        for w2, h2 in ((W, H), (W // 2, H // 2), (W, H)):
            rc = lib.rth_app_raytrace_image(app, spec.ctypes.data, w2, h2, ps, ls, depth, 0, 0, 1, 0,
                                            C.byref(pixels), C.byref(stats))
            assert rc == 0, lib.rth_last_error_string()
            frame = np.ctypeslib.as_array(C.cast(pixels, C.POINTER(C.c_float)), shape=(h2, w2, 3)).copy()
            if (w2, h2) == (W, H):
                assert np.array_equal(bits(frame), bits(theirs))
            else:
                small, _ = scene1_ref.render(spec, w2, h2, ps, ls=ls, depth=depth)
                assert np.array_equal(bits(frame), bits(small))
        world = 3
        total = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
        for rank in range(world):
            shard = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
            rc = lib.rth_app_raytrace(app, spec.ctypes.data, W, H, ps, ls, depth, 0, rank, world, 0,
                                      shard.data_ptr(), 1, C.byref(stats))
            assert rc == 0, lib.rth_last_error_string()
            total += shard          # what the NCCL sum-reduce of bench.py does across GPUs
        torch.cuda.synchronize()
        assert np.array_equal(bits(total.cpu().numpy()), bits(theirs))
    finally:
        lib.rth_app_destroy(app)


@pytest.mark.parametrize("align", ["1", "0"])
def test_large_mesh_threaded_prepare_image_matches_reference(capi, ref, align, monkeypatch):
    """A mesh big enough for the threaded host path (81 920 quads: subtree jobs, chunked
    face tables, in-place parallel staging of triangle / node records) renders the reference's
    image bit for bit, with the sibling-pair node alignment of the device layout on and off."""
    grid = (320, 256)
    monkeypatch.setenv("RAYITO_B200_NODE_ALIGN", align)
    host = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, grid)
    dev = capi.DeviceScene(host.desc)
    spec = host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    W, H, ps = 64, 36, 2
    theirs, rstats = ref.RefScene(5, None, grid).render(spec, W, H, ps, ls=1, depth=3)
    mine, stats = dev.render(cam, W, H, ps, ls=1, depth=3)
    dev.close()
    same, rel = _compare_images(mine, theirs, "large synthetic mesh, align " + align)
    assert rel <= RMSE_REL_TOL and same >= MIN_IDENTICAL_FRACTION
    assert stats.closest_rays == rstats.closest_calls and stats.any_rays == rstats.any_calls
