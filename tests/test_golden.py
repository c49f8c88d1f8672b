"""Parity against the committed golden vectors (tests/golden/, generated from the
unmodified reference by tests/golden/make_golden.py).  The CPU part pins the C
restatement and the host build of the sampler; the GPU part checks the CUDA path
through the C ABI, and works on boxes where oracle/_ref is absent."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.raybatches import bits

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _same_hits(mine, want, label):
    for f in ("shape", "face", "tri"):
        assert np.array_equal(mine[f], want[f]), "%s: %s differs" % (label, f)
    assert np.array_equal(bits(mine["t"]), bits(want["t"])), label
    hit = want["shape"] >= 0
    assert np.array_equal(bits(mine["normal"][hit]), bits(want["normal"][hit])), label
    assert np.array_equal(bits(mine["color_modifier"][hit]), bits(want["color_modifier"][hit, 0])), label


@pytest.mark.parametrize("which,fixture", [("scene1", "scene1_hits.npz"), ("scene2", "scene2_hits.npz")])
def test_port_matches_golden_hits(which, fixture, request, capi):
    from oracle import portapi
    if not portapi.available():
        pytest.skip("oracle/_build/libport.so not built")
    host = request.getfixturevalue(which + "_host")
    g = _load(fixture)
    rays = g["rays"].view(capi.RAY_DTYPE).reshape(-1) if g["rays"].dtype != capi.RAY_DTYPE else g["rays"]
    _same_hits(portapi.trace_closest(host.desc, rays, capi.HITEX_DTYPE), g["closest"], which)
    assert np.array_equal(portapi.trace_any(host.desc, rays), g["shadow"])


def test_sampler_matches_golden_stream(capi):
    from oracle import portapi
    g = _load("sample_stream.npz")
    lib = capi.core()
    checked = 0
    for key in g.files:
        parts = key.split("_")
        if parts[0] == "cmj1d":
            samples, perm = int(parts[1]), int(parts[2])
            want = g[key]
            mine = np.array([lib.rt_cmj_sample1d(i, samples, perm) for i in range(len(want))], np.float32)
            assert np.array_equal(bits(mine), bits(want)), key
            if portapi.available():
                port = np.array([portapi.lib().port_cmj_1d(i, samples, perm) for i in range(len(want))], np.float32)
                assert np.array_equal(bits(port), bits(want)), key
            checked += 1
        elif parts[0] == "cmj2d":
            xs, ys, perm = int(parts[1]), int(parts[2]), int(parts[3])
            want = g[key]
            u, v = C.c_float(), C.c_float()
            mine = np.zeros_like(want)
            for i in range(len(want)):
                lib.rt_cmj_sample2d(i, xs, ys, perm, C.byref(u), C.byref(v))
                mine[i] = (u.value, v.value)
            assert np.array_equal(bits(mine), bits(want)), key
            checked += 1
    assert checked >= 8
    # Rng: the first chunk of a 3840x2160 frame is seeded (960, 540); pixel k of that chunk
    # uses draws [18k, 18k+18) of the golden literal sequence (depth 3)
    seq = g["rng_960_540"]
    for k in (0, 1, 2, 5, 27):
        x, y = k % 960, k // 960
        mine = capi.sample_permutations(3840, 2160, 3, x, y)
        want = seq[18 * k:18 * k + 18].copy()
        if k > 0:
            want[15], want[16] = want[16], want[15]      # refill order: lens, time (RaytraceMain.cpp:167-168)
        assert np.array_equal(mine, want), k


@pytest.mark.gpu
@pytest.mark.parametrize("which,fixture", [("scene1", "scene1_hits.npz"), ("scene2", "scene2_hits.npz")])
def test_gpu_matches_golden_hits(which, fixture, request, capi):
    host = request.getfixturevalue(which + "_host")
    dev = capi.DeviceScene(host.desc)
    g = _load(fixture)
    rays = g["rays"]
    _same_hits(dev.trace_closest(rays, extended=True), g["closest"], which)
    assert np.array_equal(dev.trace_any(rays), g["shadow"])
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which,fixture,W,H,ps,ls,depth", [
    ("scene1", "scene1_render_64x36_ps2_ls1_d3.npz", 64, 36, 2, 1, 3),
    ("scene2", "scene2_render_48x32_ps2_ls2_d2.npz", 48, 32, 2, 2, 2)])
def test_gpu_matches_golden_render(which, fixture, W, H, ps, ls, depth, request, capi):
    host = request.getfixturevalue(which + "_host")
    dev = capi.DeviceScene(host.desc)
    g = _load(fixture)
    image, stats = dev.render(capi.camera_from_spec(g["camera"]), W, H, ps, ls=ls, depth=depth)
    dev.close()
    want = g["image"]
    identical = (bits(image) == bits(want)).all(axis=-1).mean()
    rmse = float(np.sqrt(np.mean((image.astype(np.float64) - want) ** 2)))
    assert identical == 1.0 and rmse == 0.0, (identical, rmse)
    assert int(stats.closest_rays) == int(g["closest_calls"]) and int(stats.any_rays) == int(g["any_calls"])
