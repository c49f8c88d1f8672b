"""The command-line stages (BASELINE.json configs[0]: "via its Makefile CLI"): cli/rayito_stage{1,2,3}
keep the reference's contract -- `make && ./rayito` writes out.ppm into the working directory -- with the
render on the B200.  Stage 1 and Stage 2 must reproduce the reference's own out_ref.ppm (md5 in
tests/golden), Stage 3 the rebuilt reference's out.ppm; --reference-pfm the bytes of the reference
programs built with WRITE_PFM; --pfm a standard PFM of the unclamped float image."""
import hashlib
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD1 = json.load(open(os.path.join(HERE, "golden", "stage1_out_ref.json")))
GOLD23 = json.load(open(os.path.join(HERE, "golden", "stage23_out.json")))


def _build(stage, tmp_path):
    """Copy the program's directory layout next to the libraries and run its Makefile there."""
    src = os.path.join(ROOT, "cli", "rayito_stage%d" % stage)
    subprocess.run(["make", "-C", src], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
    return os.path.join(src, "rayito")


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_cli_builds_and_fails_loudly_without_a_device(stage, tmp_path, capi):
    capi.host()
    exe = _build(stage, tmp_path)
    if capi.core().rt_device_count() > 0:
        pytest.skip("a CUDA device is present: covered by the gpu tests")
    proc = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert proc.returncode == 1 and "no CUDA device" in proc.stderr
    assert not os.path.exists(tmp_path / "out.ppm")


@pytest.mark.gpu
@pytest.mark.parametrize("stage", [1, 2, 3])
def test_make_and_run_writes_the_reference_image(stage, tmp_path, capi):
    capi.host()
    exe = _build(stage, tmp_path)
    subprocess.run([exe], cwd=str(tmp_path), check=True, timeout=300)
    want = GOLD1["md5"] if stage == 1 else GOLD23["stage%d" % stage]["md5"]
    assert _md5(tmp_path / "out.ppm") == want
    if stage == 2:
        assert want == GOLD23["stage2"]["out_ref_ppm_md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("stage", [1, 3])
def test_pfm_outputs(stage, tmp_path, capi):
    capi.host()
    exe = _build(stage, tmp_path)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "stage%d_pfm" % stage)
    if not os.path.exists(ref_bin):
        pytest.skip("oracle/_ref/stage%d_pfm not built" % stage)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    subprocess.run([ref_bin], cwd=str(ref_dir), check=True, timeout=300)
    subprocess.run([exe, "--reference-pfm"], cwd=str(tmp_path), check=True, timeout=300)
    assert _md5(tmp_path / "out.pfm") == _md5(ref_dir / "out.pfm")
    # the standard PFM: little-endian floats, rows bottom-up, quantising to the PPM's bytes
    subprocess.run([exe, "--pfm", "-o", "std.ppm"], cwd=str(tmp_path), check=True, timeout=300)
    raw = open(tmp_path / "std.pfm", "rb").read()
    header = b"PF\n512 512\n-1.0\n"
    assert raw.startswith(header) and len(raw) == len(header) + 512 * 512 * 12
    img = np.frombuffer(raw[len(header):], "<f4").reshape(512, 512, 3)[::-1]
    ppm = open(tmp_path / "std.ppm", "rb").read()
    payload = np.frombuffer(ppm[len(b"P6\n512 512\n255\n"):], np.uint8).reshape(512, 512, 3)
    q = (np.clip(img, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(q, payload)
