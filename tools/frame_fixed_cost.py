#!/usr/bin/env python
"""What does one rt_render call cost when there is almost nothing to render?  Device time (CUDA events inside
rt_render) and host wall time of Stage 7 scene 1 frames from 16x9 to 960x540 at 256 spp: the intercept of time against
samples is the per-frame fixed cost that limits the 8-GPU efficiency on short frames (profiles/README.md, round 2)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from rayito_b200 import build, capi
    host = capi.HostScene(capi.RECIPE_STAGE7_SCENE1, build.model_path("bumpy.obj"))
    dev = capi.DeviceScene(host.desc)
    cam = capi.camera_from_spec(host.default_camera_spec())
    for (w, h) in ((16, 9), (64, 36), (240, 135), (480, 270), (960, 540)):
        ms, wall = [], []
        for k in range(6):
            t0 = time.perf_counter()
            _img, st = dev.render(cam, w, h, 16, ls=1, depth=3)
            wall.append(1e3 * (time.perf_counter() - t0))
            ms.append(st.render_ms)
        print("%4dx%-4d %9d samples  %6d launches  device %8.3f ms (min %8.3f)  wall %8.3f ms" % (
            w, h, st.samples, st.kernel_launches, float(np.median(ms[1:])), min(ms[1:]), float(np.median(wall[1:]))))
    dev.close()


if __name__ == "__main__":
    main()
