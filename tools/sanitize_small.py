"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                   # noqa: E402
from rayito_b200 import build, capi                  # noqa: E402
from tests.raybatches import random_rays             # noqa: E402

obj = build.model_path("bumpy.obj")
for recipe, path in ((capi.RECIPE_STAGE7_SCENE1, obj), (capi.RECIPE_STAGE6_SCENE, obj), (capi.RECIPE_EDGE_LINEAR_LIST, None)):
    h = capi.HostScene(recipe, path)
    d = capi.DeviceScene(h.desc)
    cam = capi.camera_from_spec(h.default_camera_spec())
    rays = random_rays(4096, seed=1, center=(0, -0.5, 0), radius=12.0, target_radius=4.0, shadow_fraction=0.3)
    d.trace_closest(rays, extended=True)
    d.trace_any(rays)
    for kw in (dict(), dict(unified=True), dict(dynamic_top=True), dict(count_work=True), dict(rank=1, world=2, tile_size=16)):
        img, st = d.render(cam, 48, 27, 2, ls=2, depth=3, **kw)
        assert not np.isnan(img).any()
    d.close()
capi.stage23_render(3, 48, 40, 2, 2)
capi.stage23_render(2, 48, 40, 5, 1)
capi.stage1_render(32, 32)
capi.tonemap_bgra8(np.ones((7, 3), np.float32))
print("sanitize_small: done")
