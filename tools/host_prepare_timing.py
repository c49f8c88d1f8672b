#!/usr/bin/env python
"""Host-side prepare() timing of the 10 M-triangle scene (no GPU work): A/B of
RAYITO_B200_WIDE_SPLITS and the worker count on the machine it runs on."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "rayito_b200", "host", "librayito_host.so"))
lib.rth_scene_create.restype = ctypes.c_void_p
lib.rth_scene_create.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint, ctypes.c_uint]
lib.rth_scene_prepare_seconds.restype = ctypes.c_double
lib.rth_scene_prepare_seconds.argtypes = [ctypes.c_void_p]
lib.rth_scene_destroy.argtypes = [ctypes.c_void_p]
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
print("cores", os.cpu_count())
for rep in range(2):
    for wide in ("0", "1"):
        os.environ["RAYITO_B200_WIDE_SPLITS"] = wide
        t0 = time.time()
        s = lib.rth_scene_create(5, None, grid, grid)
        print("wide=%s prepare %.3f s (create %.2f s)" % (wide, lib.rth_scene_prepare_seconds(s), time.time() - t0), flush=True)
        lib.rth_scene_destroy(s)
