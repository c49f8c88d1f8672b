#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref,
compiled from /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures let the parity tests run where the reference library is absent and pin
the C restatement (oracle/port.c) to outputs of the reference itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refapi                      # noqa: E402
from rayito_b200 import build, capi            # noqa: E402
from tests.raybatches import axis_parallel_rays, random_rays   # noqa: E402


def hits_fixture(name, ref_scene, batches):
    rays = np.concatenate(batches)
    closest = ref_scene.trace_closest(rays)
    shadow = ref_scene.trace_any(rays)
    np.savez_compressed(os.path.join(HERE, name), rays=rays, closest=closest, shadow=shadow)
    print(name, len(rays), "rays,", int((closest["shape"] >= 0).sum()), "hits,", int(shadow.sum()), "occluded")


def stage23_fixture():
    """Stage 2: the reference's golden image Rayito_Stage2/out_ref.ppm (reproducible here).
    Stage 3: Rayito_Stage3/out_ref.ppm is NOT reproducible by the reference's own code
    (SURVEY.md section 4), so the vector is the rebuilt reference binary's out.ppm.  Kept
    as digests plus a 32x32 decimation (every 16th pixel) for readable failures."""
    import hashlib
    import json
    import subprocess
    import tempfile
    out = {}
    for stage in (2, 3):
        with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "oracle", "_build")) as d:
            subprocess.run([refapi.stage_binary(stage)], cwd=d, check=True)
            data = open(os.path.join(d, "out.ppm"), "rb").read()
        cut = data.index(b"255\n") + 4
        px = np.frombuffer(data[cut:], np.uint8).reshape(512, 512, 3)
        golden = open("/root/reference/Rayito_Stage%d/out_ref.ppm" % stage, "rb").read()
        gpx = np.frombuffer(golden[golden.index(b"255\n") + 4:], np.uint8).reshape(512, 512, 3)
        out["stage%d" % stage] = {
            "source": "oracle/_ref/stage%d (Rayito_Stage%d/main.cpp, g++ -O3) -> out.ppm" % (stage, stage),
            "md5": hashlib.md5(data).hexdigest(), "payload_md5": hashlib.md5(data[cut:]).hexdigest(),
            "bytes": len(data), "header": data[:cut].decode(), "width": 512, "height": 512,
            "decimated_16": px[::16, ::16].tolist(),
            "out_ref_ppm_md5": hashlib.md5(golden).hexdigest(),
            "pixels_differing_from_out_ref_ppm": int((px != gpx).any(axis=-1).sum()),
        }
        print("stage", stage, out["stage%d" % stage]["md5"], "differs from out_ref.ppm in",
              out["stage%d" % stage]["pixels_differing_from_out_ref_ppm"], "pixels")
    json.dump(out, open(os.path.join(HERE, "stage23_out.json"), "w"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "stage23":
        stage23_fixture()
        return
    obj = build.model_path("bumpy.obj")
    s1 = refapi.RefScene(1, obj)
    hits_fixture("scene1_hits.npz", s1, [
        random_rays(3072, seed=101, center=(0, -0.5, 0), radius=12.0, target_radius=4.0, shadow_fraction=0.3),
        random_rays(2048, seed=102, center=(0.1, 0, 0), radius=6.0, target_radius=1.4),
        axis_parallel_rays(1024, seed=103)])
    s2 = refapi.RefScene(2)
    hits_fixture("scene2_hits.npz", s2, [
        random_rays(4096, seed=201, center=(0, 4.0, 1.0), radius=30.0, target_radius=11.0, shadow_fraction=0.3)])

    spec1 = np.array([30, -4, 5, 15, 0, 0, 0, 0, 1, 0, 16, 0, 0, 1], np.float32)
    img, st = s1.render(spec1, 64, 36, 2, ls=1, depth=3)
    np.savez_compressed(os.path.join(HERE, "scene1_render_64x36_ps2_ls1_d3.npz"), image=img, camera=spec1,
                        closest_calls=st.closest_calls, any_calls=st.any_calls)
    spec2 = np.array([30, -4, 10, 30, 0, 5, 0, 0, 1, 0, 16, 0, 0, 1], np.float32)
    img, st = s2.render(spec2, 48, 32, 2, ls=2, depth=2)
    np.savez_compressed(os.path.join(HERE, "scene2_render_48x32_ps2_ls2_d2.npz"), image=img, camera=spec2,
                        closest_calls=st.closest_calls, any_calls=st.any_calls)

    # sample stream: literal Rng sequences and CMJ tables
    seeds = [(362436069, 521288629), (960, 540), ((960 << 16 | 1920) ^ 960, (540 << 16 | 1080) ^ 540), (1, 1), (0xffffffff, 7)]
    rng = {("rng_%d_%d" % s): refapi.rng_sequence(s[0], s[1], 512) for s in seeds}
    cmj = {}
    for samples, perm in [(16, 12345), (256, 0xdeadbeef), (65536, 99), (3, 7)]:
        cmj["cmj1d_%d_%d" % (samples, perm)] = refapi.cmj_1d(samples, perm, min(samples, 256))
    for xs, ys, perm in [(4, 4, 4242), (16, 16, 0xcafef00d), (3, 5, 17), (32, 32, 1)]:
        cmj["cmj2d_%d_%d_%d" % (xs, ys, perm)] = refapi.cmj_2d(xs, ys, perm, min(xs * ys, 256))
    np.savez_compressed(os.path.join(HERE, "sample_stream.npz"), **rng, **cmj)
    # Stage 1's golden image (Rayito_Stage1/out_ref.ppm) is two flat colours: keep its digest and
    # its structure instead of a 786 KB copy
    import hashlib
    import json
    data = open("/root/reference/Rayito_Stage1/out_ref.ppm", "rb").read()
    cut = data.index(b"255\n") + 4
    px = np.frombuffer(data[cut:], np.uint8).reshape(512, 512, 3)
    lit = np.flatnonzero(px.any(axis=(1, 2)))
    desc = {"source": "Rayito_Stage1/out_ref.ppm", "md5": hashlib.md5(data).hexdigest(), "bytes": len(data),
            "header": data[:cut].decode(), "width": 512, "height": 512,
            "first_lit_row": int(lit[0]), "last_lit_row": int(lit[-1]),
            "lit_colour": [int(v) for v in px[lit[0], 0]], "payload_md5": hashlib.md5(data[cut:]).hexdigest()}
    assert (px[lit[0]:] == px[lit[0], 0]).all() and not px[:lit[0]].any()
    json.dump(desc, open(os.path.join(HERE, "stage1_out_ref.json"), "w"), indent=1)
    stage23_fixture()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
