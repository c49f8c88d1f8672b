"""Stage 2 and Stage 3 (config C2: the Stage 3 pixel-sample sweep): the serial-Rng programs.

Golden vectors (tests/golden/stage23_out.json, made by tests/golden/make_golden.py):
  * Stage 2: the reference's own Rayito_Stage2/out_ref.ppm -- the rebuilt program
    reproduces it byte for byte, so it is a true known-answer test;
  * Stage 3: Rayito_Stage3/out_ref.ppm is NOT reproducible by the reference's own code
    (47 477 pixels differ), so the vector is the rebuilt program's
    out.ppm, as the contract prescribes.
The oracle libraries (oracle/_ref/libref_s2.so, libref_s3.so: the unmodified code with a
variable sample count) are pinned to those vectors here; the GPU path through the C ABI
must match both BIT FOR BIT (0 differing bytes, identical float images, equal ray counts)."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from tests.raybatches import bits

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "stage23_out.json")))


@pytest.fixture(scope="module")
def refapi():
    from oracle import refapi as r
    if not (r.stage_available(2) and r.stage_available(3)):
        pytest.skip("oracle/_ref/libref_s2.so / libref_s3.so not built")
    return r


def _payload_md5(rgb8):
    return hashlib.md5(np.ascontiguousarray(rgb8).tobytes()).hexdigest()


@pytest.mark.parametrize("stage", [2, 3])
def test_reference_binary_reproduces_golden(stage, refapi, tmp_path):
    binary = refapi.stage_binary(stage)
    if not os.path.exists(binary):
        pytest.skip("oracle/_ref/stage%d not built" % stage)
    subprocess.run([binary], cwd=str(tmp_path), check=True, timeout=120)
    data = open(tmp_path / "out.ppm", "rb").read()
    g = GOLD["stage%d" % stage]
    assert len(data) == g["bytes"] and hashlib.md5(data).hexdigest() == g["md5"]
    if stage == 2:
        assert g["md5"] == g["out_ref_ppm_md5"]        # the reference's own golden image


def test_oracle_stage2_library_matches_golden(refapi):
    g = GOLD["stage2"]
    rgb, rgb8, flags, rays = refapi.stage_render(2, 512, 512, 64, want_flags=True)
    assert _payload_md5(rgb8) == g["payload_md5"]
    assert np.array_equal(rgb8[::16, ::16], np.array(g["decimated_16"], np.uint8))
    assert rays == 512 * 512 * 64 + 2 * int(flags.sum())


def test_oracle_stage3_library_matches_golden(refapi):
    g = GOLD["stage3"]
    rgb, rgb8, flags, rays = refapi.stage_render(3, 512, 512, 4, 4, want_flags=True)
    assert _payload_md5(rgb8) == g["payload_md5"]
    assert np.array_equal(rgb8[::16, ::16], np.array(g["decimated_16"], np.uint8))
    assert rays == 512 * 512 * 16 + 32 * int(flags.sum())


# ---- GPU ---------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_stage2_matches_reference_golden_image(capi):
    """Known-answer test against the reference's own out_ref.ppm: 0 differing bytes."""
    g = GOLD["stage2"]
    rgb, rgb8, stats = capi.stage23_render(2, 512, 512, 64)
    assert np.array_equal(rgb8[::16, ::16], np.array(g["decimated_16"], np.uint8))
    assert _payload_md5(rgb8) == g["payload_md5"]
    ppm = g["header"].encode() + rgb8.tobytes()
    assert hashlib.md5(ppm).hexdigest() == g["out_ref_ppm_md5"]
    assert stats.samples == 512 * 512 * 64


@pytest.mark.gpu
def test_gpu_stage3_matches_rebuilt_reference(capi):
    g = GOLD["stage3"]
    rgb, rgb8, stats = capi.stage23_render(3, 512, 512, 4, 4)
    want = np.array(g["decimated_16"], np.uint8)
    assert np.array_equal(rgb8[::16, ::16], want), "decimated image differs in %d pixels" % int(
        (rgb8[::16, ::16] != want).any(axis=-1).sum())
    assert _payload_md5(rgb8) == g["payload_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("stage,w,h,nu,nv", [
    (3, 128, 128, 1, 1), (3, 128, 128, 2, 2), (3, 160, 96, 4, 4), (3, 96, 96, 8, 8), (3, 64, 64, 16, 16), (3, 50, 37, 3, 2), (3, 33, 31, 1, 1), (3, 3, 3, 1, 1),
    (2, 128, 128, 64, 1), (2, 100, 60, 7, 1), (2, 256, 256, 1, 1), (2, 17, 9, 3, 1)])
def test_gpu_float_image_and_rays_equal_oracle(capi, refapi, stage, w, h, nu, nv):
    """The spp sweep of config C2 (and Stage 2 at other sample counts): float images before
    clamp bit-identical, every ray accounted for."""
    ref_rgb, ref_rgb8, flags, ref_rays = refapi.stage_render(stage, w, h, nu, nv, want_flags=True)
    rgb, rgb8, stats = capi.stage23_render(stage, w, h, nu, nv)
    same = (bits(rgb) == bits(ref_rgb)).all(axis=-1)
    assert same.all(), "%d/%d pixels differ, first at %s" % ((~same).sum(), same.size, np.argwhere(~same)[0])
    assert np.array_equal(rgb8, ref_rgb8)
    assert stats.closest_rays == ref_rays
    assert stats.samples == flags.size


@pytest.mark.gpu
def test_gpu_stage23_rejects_bad_arguments(capi):
    with pytest.raises(capi.RtError):
        capi.stage23_render(4, 64, 64, 1, 1)
    with pytest.raises(capi.RtError):
        capi.stage23_render(3, 64, 64, 0, 1)
    with pytest.raises(capi.RtError):
        capi.stage23_render(3, 1, 64, 1, 1)
